"""CPU: host-side logic -- library loads and exports the whole C ABI, the host
Rayleigh-Ritz eigensolver, the synthetic generator's shard invariance, the
wrapper's argument handling, the sharding helpers (incl. a world_size-2 gloo
run)."""

import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

from snapatac2_b200 import synth, dist, tl, MiniAnnData, _lib
from snapatac2_b200.engine import sym_eig

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = (ROOT / "include" / "snapb200.h").read_text()
    declared = set(re.findall(r"\b(snapb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.snapb200_version() >= 100


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from snapatac2_b200 import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 129])
def test_sym_eig_matches_lapack(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a = a + a.T
    w, v = sym_eig(a)
    np.testing.assert_allclose(w, np.linalg.eigvalsh(a), rtol=0, atol=1e-12 * max(1, n))
    np.testing.assert_allclose(a @ v, v * w, atol=1e-11 * max(1, n))
    np.testing.assert_allclose(v.T @ v, np.eye(n), atol=1e-12 * max(1, n))


def test_sym_eig_degenerate_and_banded():
    a = np.diag([2.0, 2.0, 2.0, -1.0, 0.0])
    w, v = sym_eig(a)
    np.testing.assert_allclose(w, [-1, 0, 2, 2, 2], atol=1e-14)
    n = 40
    t = np.diag(np.linspace(1, 2, n)) + np.diag(np.full(n - 8, 0.1), 8) + np.diag(np.full(n - 8, 0.1), -8)
    w, v = sym_eig(t)
    np.testing.assert_allclose(w, np.linalg.eigvalsh(t), atol=1e-13)


def test_synth_is_shard_invariant_and_sorted():
    spec = synth.make_spec(300, 5000, 200, n_clusters=10, seed=4)
    full = synth.generate_csr(spec)
    parts = [synth.generate_csr(spec, a, b) for a, b in ((0, 77), (77, 200), (200, 300))]
    stacked = sp.vstack(parts, format="csr")
    assert (full != stacked).nnz == 0
    assert full.has_sorted_indices
    rows = np.diff(full.indptr)
    assert rows.max() <= 200 and rows.min() > 150
    for i in range(0, 300, 37):
        r = full.indices[full.indptr[i]:full.indptr[i + 1]]
        assert np.all(np.diff(r) > 0)           # sorted, unique
    z = synth.cluster_of_rows(spec, np.arange(300))
    assert z.min() >= 0 and z.max() < 10 and len(np.unique(z)) >= 8


def test_wrapper_argument_errors_before_any_gpu_work():
    spec = synth.make_spec(50, 300, 20, n_clusters=3, seed=1)
    ad = MiniAnnData(synth.generate_csr(spec))
    with pytest.raises(NameError):
        tl.spectral(ad)                                      # _embedding.py:229
    with pytest.raises(ValueError):
        tl.spectral(ad, features=None, sample_size=1)        # :238
    with pytest.raises(ValueError):
        tl.spectral(ad, features=None, sample_size=1.5)      # :243
    with pytest.raises(NotImplementedError):
        tl.spectral(ad, features=None, distance_metric="jaccard")
    with pytest.raises(NotImplementedError):
        tl.spectral(ad, features=None, sample_size=10, distance_metric="jaccard")


def test_orthogonalize_matches_the_reference_statement():
    import oracle
    rng = np.random.default_rng(0)
    u = rng.standard_normal((400, 7))
    v = np.sort(rng.uniform(0.01, 1.0, 7))[::-1]
    ev, evec = tl.orthogonalize(v, u)
    ev_o, evec_o = oracle.orthogonalize(v, u)
    np.testing.assert_allclose(ev, ev_o, rtol=1e-12)
    np.testing.assert_allclose(np.abs(evec), np.abs(evec_o), rtol=1e-9, atol=1e-12)
    # columns are orthonormal and diagonalise the projected operator
    np.testing.assert_allclose(evec.T @ evec, np.eye(7), atol=1e-9)


def test_feature_mask_and_weight_permutation():
    mask, fw = tl._feature_mask(np.array([5, 1, 3]), 8, [50.0, 10.0, 30.0])
    assert mask.tolist() == [False, True, False, True, False, True, False, False]
    assert fw.tolist() == [10.0, 30.0, 50.0]
    with pytest.raises(ValueError):
        tl._feature_mask(np.array([1, 1]), 4, None)
    m, _ = tl._feature_mask(np.array([True, False, True]), 3, None)
    assert m.tolist() == [True, False, True]


def test_row_split_helpers():
    indptr = np.concatenate([[0], np.cumsum(np.r_[np.full(10, 100), np.full(90, 1)])])
    b = dist.balanced_row_splits(indptr, 4)
    assert b[0] == 0 and b[-1] == 100 and np.all(np.diff(b) >= 0)
    nnz = np.diff(indptr[b])
    assert nnz.max() <= 400                       # heavy rows spread, light rows lumped
    assert dist.equal_row_splits(10, 3).tolist() == [0, 3, 6, 10]
    assert dist.shard_offsets([3, 4, 5]) == [0, 3, 7]


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as td
from snapatac2_b200 import dist, synth
td.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, ws = dist.world()
assert (rank, ws) == (int(sys.argv[1]), 2) and dist.is_distributed()
# the unique-id exchange path (payload is opaque bytes)
uid = bytes(range(128)) if rank == 0 else None
got = dist.broadcast_bytes(uid, src=0)
assert got == bytes(range(128))
# shard bookkeeping: every rank generates its own row block, offsets come from an all-gather
spec = synth.make_spec(120, 2000, 60, n_clusters=5, seed=8)
bounds = dist.equal_row_splits(spec.n, ws)
mine = synth.generate_csr(spec, int(bounds[rank]), int(bounds[rank + 1]))
n_locals = dist.allgather_ints(mine.shape[0])
assert sum(n_locals) == spec.n and dist.shard_offsets(n_locals)[rank] == int(bounds[rank])
# global document frequencies = sum of shard counts (what the library all-reduces)
import torch
df = torch.from_numpy(np.bincount(mine.indices, minlength=spec.m).astype(np.int64))
td.all_reduce(df)
full = synth.generate_csr(spec)
assert np.array_equal(df.numpy(), np.bincount(full.indices, minlength=spec.m))
# multi_spectral's exchange of the sampled rows: shard pieces stacked in rank order = the global rows
import scipy.sparse as sp
rows = np.sort(np.random.RandomState(3).choice(spec.n, 40, replace=False))
r0 = int(bounds[rank])
local = rows[(rows >= r0) & (rows < r0 + mine.shape[0])] - r0
stacked = sp.vstack(dist.allgather_objects(sp.csr_matrix(mine[local])), format="csr")
assert (stacked != full[rows]).nnz == 0
# Nystrom per-chunk normalisation on shards (tl._nystrom_normalise): global chunk_size blocks that span
# the two shards take their column sums and smallest positive degree from both
from snapatac2_b200 import tl
rng = np.random.RandomState(5)
q_full = rng.standard_normal((spec.n, 4))
q_full[7] *= -30.0                                     # some non-positive degrees, to exercise the clamp
v = np.array([1.0, 0.5, 0.3, 0.2])
chunk = 50                                             # shards are 60 rows: chunk 1 spans both
want = []
for i in range(0, spec.n, chunk):
    qc = q_full[i:i + chunk].copy()
    t = qc.sum(axis=0) * v
    dd = qc @ t
    dd[dd <= 0] = np.min(dd[dd > 0])
    want.append(qc / np.sqrt(dd)[:, None])
want = np.vstack(want)
got = tl._nystrom_normalise(q_full[r0:r0 + mine.shape[0]].copy(), v, chunk, spec.n, r0)
assert np.allclose(got, want[r0:r0 + mine.shape[0]], rtol=1e-12, atol=0), np.abs(got - want[r0:r0 + mine.shape[0]]).max()
assert [c for c, _, _ in dist.chunk_overlaps(r0, mine.shape[0], chunk)] == ([0, 1] if rank == 0 else [1, 2])
assert np.array_equal(dist.allreduce_array(np.array([rank + 1.0, 5.0]), "sum"), [3.0, 10.0])
assert np.array_equal(dist.allreduce_array(np.array([rank + 1.0]), "min"), [1.0])
# pp.knn on row shards: the points are all-gathered on the host, the rank's query range starts at its first row
from snapatac2_b200 import pp
pts_full = rng.standard_normal((spec.n, 6))
pts_all, q0 = pp._gather_points(pts_full[r0:r0 + mine.shape[0]])
assert q0 == r0 and np.array_equal(pts_all, pts_full)
td.barrier()
td.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=str(ROOT), port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"ok {r}" in o


def test_row_blocks_sources():
    """tl.RowBlocks: a backed element with chunked(), a row-sliceable lazy array and a plain iterable of CSR
    blocks all yield the matrix's rows in order (what Engine.load_blocks assembles on the device)."""
    X = sp.random(53, 40, density=0.2, format="csr", random_state=1)

    class Backed:
        shape = X.shape

        def chunked(self, chunk_size):
            for i in range(0, X.shape[0], chunk_size):
                yield X[i:i + chunk_size], i, min(i + chunk_size, X.shape[0])

    class Sliceable:
        shape = X.shape

        def __getitem__(self, key):
            return X[key]

    for src in (Backed(), Sliceable(), (X[i:i + 7] for i in range(0, 53, 7)), [X[:20], X[20:]]):
        blocks = list(tl.RowBlocks(src, 40).blocks(16))
        got = sp.vstack(blocks, format="csr")
        assert (got != X).nnz == 0 and got.shape == X.shape

    # the on-disk CSR group of an .h5ad (h5py is absent here: numpy arrays slice the same way)
    class Group(dict):
        attrs = {"encoding-type": "csr_matrix", "shape": np.array([53, 40])}

    g = Group(indptr=X.indptr, indices=X.indices, data=X.data)
    rb = tl.RowBlocks.from_csr_group(g)
    assert rb.n_vars == 40
    for chunk in (1, 16, 53, 1000):
        got = sp.vstack(list(rb.blocks(chunk)), format="csr")
        assert (got != X).nnz == 0 and got.shape == X.shape
    Group.attrs = {"encoding-type": "csc_matrix", "shape": np.array([53, 40])}
    with pytest.raises(ValueError, match="CSR layout"):
        tl.RowBlocks.from_csr_group(g)
    Group.attrs = {"encoding-type": "csr_matrix", "shape": np.array([53, 40])}

    class FakeAnn:
        def __init__(self, x):
            self.X, self.n_vars, self.n_obs = x, 40, 53

    assert isinstance(tl._get_csr(FakeAnn(Backed())), tl.RowBlocks)
    assert sp.issparse(tl._get_csr(FakeAnn(X)))
    with pytest.raises(ValueError):
        tl._get_csr(FakeAnn(X.tocsc()))


def _delta_selftest(idx):
    import ctypes as C
    from snapatac2_b200 import _lib
    lib = _lib.load()
    n_side = C.c_int64(-1)
    idx = np.ascontiguousarray(idx)
    rc = lib.snapb200_delta_selftest_host(_lib.ptr(idx), 8 * idx.dtype.itemsize, idx.size, C.byref(n_side))
    return rc, n_side.value


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_delta_encoding_of_the_index_transfer_round_trips(dtype):
    """csrc/ingest.cu ships column indices as 16-bit differences + a side list of absolute values
    (markers at tile starts, row starts and gaps >= 0xFFFF); the host-only replay must give back every
    index, whatever the gaps."""
    from snapatac2_b200 import synth
    rng = np.random.default_rng(5)
    # (1) a realistic CSR: sorted rows, mixed lengths
    X = synth.generate_csr(synth.make_spec(700, 500_000, 900, n_clusters=5, seed=3))
    rc, n_side = _delta_selftest(X.indices.astype(dtype))
    assert rc == 0
    # markers: one per 2048-entry tile + about one per row (+ rare wide gaps)
    assert n_side <= X.nnz // 2048 + 1 + X.shape[0] + X.nnz // 50
    # (2) adversarial gaps around the 16-bit limit, equal neighbours, descending runs, tiny and ragged sizes
    base = np.cumsum(rng.choice([0, 1, 7, 65534, 65535, 65536, 70000], size=40_000)).astype(np.int64) % (2**31 - 1)
    for arr in (base, base[::-1].copy(), base[:1], base[:7], base[:2047], base[:2048], base[:2049], base[:4099],
                np.zeros(5000, np.int64), np.full(3000, 2**31 - 1, np.int64)):
        rc, n_side = _delta_selftest(arr.astype(dtype))
        assert rc == 0, arr[:8]
        assert n_side >= (arr.size + 2047) // 2048
    # (3) more than one 3 Mi-entry chunk
    big = np.sort(rng.integers(0, 2**31 - 1, size=(3 << 20) + 12345)).astype(dtype)
    assert _delta_selftest(big)[0] == 0
    # (4) nothing but wide gaps: the side list cannot hold a chunk -> the caller is told to ship plain int32
    wide = (np.arange(3 << 20, dtype=np.int64) % 2) * 100_000
    assert _delta_selftest(wide.astype(dtype))[0] == 1


def test_delta_encoding_property_random_streams():
    """Random index streams (hypothesis): whatever the mix of gaps, repeats, descents and run lengths, the chunk
    format either replays to the input or reports that the side list would overflow (never a mismatch)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 70_000),
           wide=st.floats(0.0, 1.0), dtype=st.sampled_from([np.int32, np.int64]))
    def run(seed, n, wide, dtype):
        rng = np.random.default_rng(seed)
        steps = np.where(rng.random(n) < wide, rng.integers(-(2**31 - 1), 2**31 - 1, size=n), rng.integers(0, 70_000, size=n))
        idx = np.abs(np.cumsum(steps)) % (2**31 - 1)
        rc, n_side = _delta_selftest(idx.astype(dtype))
        assert rc in (0, 1)
        if rc == 0:
            assert n_side >= (n + 2047) // 2048

    run()
