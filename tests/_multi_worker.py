"""Worker for tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU).

Shard invariance: the N-rank row-sharded solve must reproduce the 1-GPU solve
(idf / degree to 1e-9, eigenvalues to 1e-6 relative, eigenvectors |cos| >= 0.99999)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as td

from snapatac2_b200 import Engine, MiniAnnData, dist, synth, tl

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
td.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
mode = sys.argv[1] if len(sys.argv) > 1 else "auto"

# SNAPB200_MULTI_CONFIG=c3: the same check on BASELINE.json configs[2] (1M x 500k, ~5k nnz/cell, k=30);
# rank 0 then holds its shard and the whole matrix (fits one B200)
BIG = os.environ.get("SNAPB200_MULTI_CONFIG", "") == "c3"
if BIG:
    spec = synth.make_spec(1_000_000, 500_000, 5_000, n_clusters=48, seed=0)
    k = 30
else:
    spec = synth.make_spec(9000, 60000, 1200, n_clusters=40, seed=3)
    k = 20
bounds = dist.equal_row_splits(spec.n, world)
r0, r1 = int(bounds[rank]), int(bounds[rank + 1])

eng = Engine(local)
eng.set_spmm_mode(mode)
dist.attach_engine_comm(eng)
eng.generate(spec, row0=r0, n_local=r1 - r0)
idf, deg = eng.prepare()
evals, evecs = eng.eigsh(k, seed=0)
stats = eng.stats()

# gather the sharded result on rank 0
parts = [None] * world
td.gather_object((deg, evecs), parts if rank == 0 else None, dst=0)

# the public wrapper in distributed mode: every rank passes its own row block
if not BIG:
    X_local = eng.export_csr()
    ad = MiniAnnData(X_local)
    tl._engine = eng
    ev_w, emb_w = tl.spectral(ad, n_comps=k, features=None, inplace=False, weighted_by_sd=False)
    assert np.allclose(ev_w, evals, rtol=1e-9), (ev_w, evals)
    assert emb_w.shape == (r1 - r0, k)

if rank == 0:
    deg_all = np.concatenate([p[0] for p in parts])
    evec_all = np.concatenate([p[1] for p in parts], axis=0)
    solo = Engine(local)
    solo.set_spmm_mode(mode)
    solo.generate(spec)
    idf1, deg1 = solo.prepare()
    ev1, evec1 = solo.eigsh(k, seed=0)
    np.testing.assert_allclose(idf, idf1, rtol=1e-12)
    np.testing.assert_allclose(deg_all, deg1, rtol=1e-9)
    np.testing.assert_allclose(evals, ev1, rtol=1e-6)
    cos = np.abs(np.sum(evec_all * evec1, axis=0)) / (np.linalg.norm(evec_all, axis=0) * np.linalg.norm(evec1, axis=0))
    assert cos.min() > 0.99999, cos
    assert abs(np.linalg.norm(evec_all[:, 3]) - 1.0) < 1e-5
    print(f"MULTI_OK world={world} mode={mode} n={spec.n} k={k} n_ops={stats['n_ops']} ms_comm={stats['ms_comm']:.3f} "
          f"max_rel_eval_diff={np.max(np.abs(evals - ev1) / np.abs(ev1)):.3e} min_cos={cos.min():.8f}")
    solo.close()
if BIG:
    td.barrier()
    eng.close()
    td.destroy_process_group()
    sys.exit(0)

# ---- multi-view (tl.multi_spectral, embedding.rs:388-452) on row shards: every rank passes its
#      block of both views; the result must match the CPU oracle on the full views
spec_b = synth.make_spec(spec.n, 4000, 150, n_clusters=40, seed=4)
eng.set_feature_weights(None)
eng.generate(spec_b, row0=r0, n_local=r1 - r0)
XB_local = eng.export_csr().astype(np.float64)
XB_local.data = 1.0 + (XB_local.indices % 3)          # counts: a function of the column, shard independent
rows = np.sort(np.random.RandomState(7).choice(spec.n, 2000, replace=False))
km = 12
ev_m, emb_m = tl.multi_spectral([MiniAnnData(X_local), MiniAnnData(XB_local)], n_comps=km, features=None,
                                weights=[1.0, 0.5], weighted_by_sd=False, engine=eng, sample_rows=rows)
parts_m = [None] * world
td.gather_object(emb_m, parts_m if rank == 0 else None, dst=0)
if rank == 0:
    import oracle
    from conftest import eigvec_agreement
    XA = synth.generate_csr(spec, dtype=np.float64)
    XB = synth.generate_csr(spec_b, dtype=np.float64)
    XB.data = 1.0 + (XB.indices % 3)
    ev_o, evec_o = oracle.multi_spectral_embedding([XA, XB], [None, None], [1.0, 0.5], km, 0, sample_rows=rows)
    emb_all = np.concatenate(parts_m, axis=0)
    np.testing.assert_allclose(ev_m, ev_o, rtol=1e-4)
    assert eigvec_agreement(ev_o, evec_o, emb_all).min() >= 0.999
    print(f"MULTIVIEW_OK world={world}")
# ---- Nystrom path on row shards (embedding.rs:61-129): every rank passes its block, the landmark matrix
#      is itself sharded (each rank contributes the landmarks inside its block); against the CPU oracle
lm = np.sort(np.random.RandomState(11).choice(spec.n, 2500, replace=False))
kn, chunk = 10, 2000
v_n, q_n = tl.spectral_embedding_nystrom(eng, X_local, None, kn, lm.size, False, chunk, landmarks=lm)
parts_n = [None] * world
td.gather_object(q_n, parts_n if rank == 0 else None, dst=0)
if rank == 0:
    import oracle
    from conftest import eigvec_agreement
    XA = synth.generate_csr(spec, dtype=np.float64)
    ev_o, q_o = oracle.spectral_embedding_nystrom(XA, None, kn, lm, chunk)
    q_all = np.concatenate(parts_n, axis=0)
    np.testing.assert_allclose(v_n, ev_o, rtol=1e-4)
    assert eigvec_agreement(ev_o, q_o, q_all).min() >= 0.999
    np.testing.assert_allclose(np.linalg.norm(q_all, axis=0), np.linalg.norm(q_o, axis=0), rtol=1e-3)
    print(f"NYSTROM_OK world={world}")

# ---- pp.knn on row shards: every rank passes its rows of the embedding, gets its rows of the graph (global columns);
#      against the CPU oracle on the whole point set
from snapatac2_b200 import pp
rng_k = np.random.default_rng(5)
centres = rng_k.normal(scale=3.0, size=(6, 30))
lab = rng_k.integers(0, 6, size=5003)
P_all = centres[lab] + rng_k.normal(size=(5003, 30)) * 0.3
kb = dist.equal_row_splits(P_all.shape[0], world)
ad_k = MiniAnnData(np.ones((int(kb[rank + 1] - kb[rank]), 2)))
ad_k.obsm["X_spectral"] = P_all[int(kb[rank]):int(kb[rank + 1])]
pp.knn(ad_k, n_neighbors=30, engine=eng)
g_local = ad_k.obsp["distances"]
assert g_local.shape == (int(kb[rank + 1] - kb[rank]), P_all.shape[0])
parts_k = [None] * world
td.gather_object((g_local.indices, g_local.data), parts_k if rank == 0 else None, dst=0)
if rank == 0:
    import oracle
    want = oracle.knn.nearest_neighbour_graph_kdtree(P_all, 30)
    np.testing.assert_array_equal(np.concatenate([p[0] for p in parts_k]), want.indices)
    np.testing.assert_array_equal(np.concatenate([p[1] for p in parts_k]), want.data)
    print(f"KNN_OK world={world}")
td.barrier()
eng.close()
td.destroy_process_group()
