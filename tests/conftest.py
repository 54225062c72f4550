import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    import scipy.sparse as sp
    z = np.load(GOLDEN / f"{name}.npz")
    X = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    return X, z


def eigvec_agreement(evals_ref, U_ref, U, rel_cluster=2e-3):
    """Per-component |cos| between two eigenvector sets, with components whose
    reference eigenvalues lie within ``rel_cluster`` of a neighbour compared as
    subspaces (smallest principal-angle cosine of the cluster) -- north_star:
    "up to sign/rotation within degenerate clusters"."""
    k = len(evals_ref)
    Un = U / np.linalg.norm(U, axis=0, keepdims=True)
    Rn = U_ref / np.linalg.norm(U_ref, axis=0, keepdims=True)
    out = np.zeros(k)
    i = 0
    while i < k:
        j = i + 1
        while j < k and abs(evals_ref[j] - evals_ref[j - 1]) <= rel_cluster * max(abs(evals_ref[j - 1]), 1e-300):
            j += 1
        if j - i == 1:
            out[i] = abs(float(Un[:, i] @ Rn[:, i]))
        else:
            qa, _ = np.linalg.qr(Un[:, i:j])
            qb, _ = np.linalg.qr(Rn[:, i:j])
            s = np.linalg.svd(qa.T @ qb, compute_uv=False)
            out[i:j] = s.min()
        i = j
    return out


@pytest.fixture(scope="session")
def engine():
    from snapatac2_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
