"""A minimal in-memory AnnData look-alike.

``anndata`` is not installed in this image; the reference's boundary only uses
``X`` (scipy CSR), ``var[...]``, ``shape`` / ``n_obs`` / ``n_vars``, ``obsm``,
``uns`` and ``isbacked`` (tools/_embedding.py:223-293), plus ``obsp`` for the
neighbour graph (preprocessing/_knn.py:85), so this is all the mirror needs.  A real ``anndata.AnnData`` works with ``tl.spectral`` as well.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp


class MiniAnnData:
    def __init__(self, X, obs=None, var=None):
        if not sp.issparse(X):
            X = sp.csr_matrix(np.asarray(X))
        self.X = X
        n, m = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(m).astype(str))
        self.obsm: dict = {}
        self.obsp: dict = {}
        self.uns: dict = {}
        self.isbacked = False

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    def __repr__(self):
        return f"MiniAnnData(n_obs={self.n_obs}, n_vars={self.n_vars}, obsm={list(self.obsm)}, uns={list(self.uns)})"
