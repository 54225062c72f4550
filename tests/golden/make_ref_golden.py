"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN CODE (not the oracle).

Run from the repo root, in the container that has ``/root/reference``:

    python tests/golden/make_ref_golden.py

What is executed, unmodified, from ``/root/reference``:

* ``snapatac2-python/python/snapatac2/tools/_embedding.py`` is loaded with
  ``importlib`` under stub modules for ``snapatac2._snapatac2`` (the PyO3
  extension, not buildable here: no cargo) and ``snapatac2._utils``.  From it run
    - ``SpectralMatrixFree(out_dim=k, feature_weights=w).fit(X).transform()``
      (:434-481, with ``_eigen`` :474-481) -- the reference's pure-Python
      statement of the matrix-free cosine embedding: eigenvalues, eigenvectors
      and (captured from the ``_eigen`` call) the degree vector;
    - ``orthogonalize`` (:397-413) on the Nystrom fixture;
    - the wrapper ``spectral`` (:129-295) itself, with ``internal.spectral_embedding``
      answered from the recorded vectors, for the post-processing semantics
      (n_comps clamp, ``weighted_by_sd``, inplace keys).
* the Python snippet embedded in ``snapatac2-python/src/embedding.rs:456-460``
  (``frobenius_norm``) is cut out of the Rust source text and ``exec``-ed on both
  scipy containers (``csr_matrix``: ``np.power`` resolves to ``__pow__`` = matrix
  power; ``csr_array``: element-wise).

IDF weights are computed by Rust only (embedding.rs:269-286); the fixtures
therefore feed the *stored* IDF weights of the companion oracle fixture to
``feature_weights`` (the reference's Python class takes them as an argument)
and the test asserts separately that those weights follow the closed form.

Outputs: ``<name>_ref.npz`` next to the input fixtures (inputs are not
duplicated: the ``_ref`` file refers to ``<name>.npz`` for the CSR arrays).
"""

import importlib.util
import re
import sys
import types
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

REF = Path("/root/reference/snapatac2-python")
OUT = Path(__file__).resolve().parent


def load_reference_embedding_module():
    """The reference's tools/_embedding.py, unmodified, with the two modules it imports at the top
    (:10-11) replaced by empty stubs."""
    pkg = types.ModuleType("snapatac2")
    pkg.__path__ = []
    utils = types.ModuleType("snapatac2._utils")
    utils.get_igraph_from_adjacency = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    utils.is_anndata = lambda x: True
    internal = types.ModuleType("snapatac2._snapatac2")
    internal.AnnData = object
    internal.AnnDataSet = object
    sys.modules["snapatac2"] = pkg
    sys.modules["snapatac2._utils"] = utils
    sys.modules["snapatac2._snapatac2"] = internal
    pkg._utils, pkg._snapatac2 = utils, internal
    path = REF / "python" / "snapatac2" / "tools" / "_embedding.py"
    spec = importlib.util.spec_from_file_location("snapatac2_reference_embedding", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, internal


def frobenius_snippet():
    """The literal Python source inside ``frobenius_norm`` (embedding.rs:454-471)."""
    src = (REF / "src" / "embedding.rs").read_text()
    m = re.search(r'fn frobenius_norm.*?from_code_bound\(\s*py,\s*"(.*?)",', src, re.S)
    assert m, "frobenius_norm snippet not found"
    code = m.group(1)
    lines = code.split("\n")
    # the Rust string literal keeps the source indentation of the continuation lines: dedent them
    body = [ln for ln in lines[1:]]
    ind = min(len(ln) - len(ln.lstrip()) for ln in body if ln.strip())
    code = lines[0] + "\n" + "\n".join("    " + ln[ind:] for ln in body) + "\n"
    ns = {}
    exec(code, ns)
    return code, ns["f"]


def run_twin(mod, X, k, w):
    """SpectralMatrixFree.fit/transform with the degree vector captured from the _eigen call."""
    captured = {}
    orig = mod._eigen

    def spy(Xn, D, k):
        captured["dinv"] = np.array(D, dtype=np.float64)
        return orig(Xn, D, k)

    mod._eigen = spy
    try:
        np.random.seed(0)
        model = mod.SpectralMatrixFree(out_dim=k, feature_weights=w)
        evals, evecs = model.fit(sp.csr_matrix(X, dtype=np.float64)).transform()
    finally:
        mod._eigen = orig
    return np.asarray(evals), np.asarray(evecs), 1.0 / captured["dinv"]


def main():
    from conftest import load_golden
    mod, internal = load_reference_embedding_module()
    code, frob = frobenius_snippet()
    print("frobenius_norm snippet (embedding.rs):\n" + code)

    for name in ["tile_600x4000", "counts_300x1000", "dense_50x100", "masked_400x3000"]:
        X, z = load_golden(name)
        k = int(z["k"])
        if "features" in z:
            X = sp.csr_matrix(X[:, np.flatnonzero(z["features"])])
        w = z["feature_weights"] if "feature_weights" in z else z["idf"]
        evals, evecs, degree = run_twin(mod, X, k, w)
        # the wrapper's post-processing, executed from the reference: spectral() with the native call
        # answered by the vectors above (the stub stands where the PyO3 function is)
        internal.spectral_embedding = lambda adata, features, n_comps, rs, fw, _e=evals, _v=evecs: (_e[:n_comps].copy(), _v[:, :n_comps].copy())
        from snapatac2_b200 import MiniAnnData
        ad = MiniAnnData(X)
        wrapped = mod.spectral(ad, n_comps=k, features=None, inplace=False)
        np.savez_compressed(OUT / f"{name}_ref.npz", evals=evals, evecs=evecs, degree=degree, weights=np.asarray(w),
                            wrapped_evals=wrapped[0], wrapped_evecs=wrapped[1])
        print(f"{name}_ref: k={k} evals[:4]={evals[:4]}")

    # orthogonalize (:397-413) on the Nystrom fixture's extension
    _, z = load_golden("nystrom_500x3000")
    ev_o, q_o = mod.orthogonalize(np.array(z["evals"]), np.array(z["q"]))
    np.savez_compressed(OUT / "nystrom_500x3000_ref.npz", evals=np.asarray(ev_o), evecs=np.asarray(q_o))
    print("nystrom_500x3000_ref: orthogonalize evals", np.real(ev_o))

    # frobenius_norm snippet on unit-norm rows of two views, both scipy containers
    import oracle
    from snapatac2_b200 import synth
    views = {}
    for tag, (n, m, nnz, seed) in {"a": (150, 900, 60, 1), "b": (150, 300, 30, 2)}.items():
        Xv = synth.generate_csr(synth.make_spec(n, m, nnz, n_clusters=4, seed=seed), dtype=np.float64)
        xhat = oracle.normalize(Xv, oracle.idf(Xv))
        views[f"{tag}_indptr"] = xhat.indptr.astype(np.int64)
        views[f"{tag}_indices"] = xhat.indices.astype(np.int32)
        views[f"{tag}_data"] = xhat.data
        views[f"{tag}_shape"] = np.asarray(xhat.shape, dtype=np.int64)
        views[f"{tag}_sum_csr_matrix"] = np.float64(frob(sp.csr_matrix(xhat)))
        views[f"{tag}_sum_csr_array"] = np.float64(frob(sp.csr_array(xhat)))
        print(f"view {tag}: snippet on csr_matrix = {views[f'{tag}_sum_csr_matrix']:.6f}, on csr_array = {views[f'{tag}_sum_csr_array']:.6f}")
    np.savez_compressed(OUT / "frobenius_snippet_ref.npz", snippet=np.array(code), **views)


if __name__ == "__main__":
    main()
