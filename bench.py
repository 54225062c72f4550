#!/usr/bin/env python
"""Benchmark of the hot path: ``snap.tl.spectral`` cells/s on a synthetic
binarised tile matrix (BASELINE.json), plus the SpMM roofline line and the CPU
baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full spectral embedding of the resident matrix: prepare
(feature-major copy, IDF, row norms, column sums, degrees) + the block-Lanczos
eigensolve + eigenvectors copied to the host.  `value` times K steps with the
CSR already in HBM; `e2e` times the same call with the CSR in pinned host
memory (host->device copy inside the timed region).  The workload is C3 of
BASELINE.json (1M x 500k, ~5k nnz/cell, n_comps=30) at every N -- total work
fixed, rows sharded across ranks ("scaling": "strong").
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

CONFIGS = {
    # name: (n cells, m bins, nominal nnz/cell, clusters, n_comps)
    "c1": (5_000, 100_000, 3_000, 48, 30),
    "c2": (100_000, 500_000, 5_000, 48, 30),
    "c3": (1_000_000, 500_000, 5_000, 48, 30),
    "c4": (10_000_000, 500_000, 3_000, 64, 50),
    "tiny": (2_000, 20_000, 500, 40, 30),
    "c3s": (125_000, 500_000, 5_000, 48, 30),   # the size of one rank's shard of c3 at 8 GPUs (tuning aid)
    # c5 = multi_spectral: the ATAC view below + an RNA count view of (n, 30_000, ~1500 nnz/cell), see run_multiview
    "c5": (1_000_000, 500_000, 5_000, 48, 30),
    "c5s": (125_000, 500_000, 5_000, 48, 30),
}
C5_RNA = (30_000, 1_500)     # bins, nominal nnz per cell of the second view
CONFIG_TEXT = {
    "c1": "synthetic 5k-cell x 100k-bin binarised tile matrix (~3k nnz/cell), n_comps=30",
    "c2": "synthetic 100k cells x 500k bins (~5k nnz/cell), n_comps=30",
    "c3": "synthetic 1M cells x 500k bins (~5k nnz/cell), n_comps=30, row-sharded",
    "c4": "synthetic 10M cells x 500k bins (~3k nnz/cell), n_comps=50, row-sharded",
    "tiny": "synthetic 2k x 20k (~500 nnz/cell), n_comps=30 (harness test only)",
    "c3s": "synthetic 125k cells x 500k bins (~5k nnz/cell), n_comps=30 (one rank's share of c3 at 8 GPUs; tuning aid)",
    "c5": "snap.tl.multi_spectral: synthetic ATAC tile (1M x 500k, ~5k nnz/cell, binarised) + RNA count (1M x 30k, ~1.5k nnz/cell, counts 1-3), n_comps=30, row-sharded",
    "c5s": "snap.tl.multi_spectral: ATAC 125k x 500k + RNA 125k x 30k (one rank's share of c5 at 8 GPUs; tuning aid)",
}
METRIC = "snap.tl.spectral cells/s"
CPU_SAMPLE_ROWS = 4000


def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while running."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t_end = time.time()
        t_begin = getattr(self, "t_begin", 0.0)
        inside = [r for (ts, r) in self.rows if t_begin <= ts <= t_end + 0.2]
        region = "timed region"
        if not inside:
            inside = [r for (_, r) in self.rows]
            region = "warm-up + timed region (timed region shorter than the sampling period)"
        sm, smax, reasons = [], [], set()
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "sampled_over": region, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# CPU legs (the only place bench.py runs oracle/)
# --------------------------------------------------------------------------
def cpu_oracle_step(cfg_name, rows=CPU_SAMPLE_ROWS):
    """One run of the reference's CPU path (oracle: same scipy ARPACK call) on a
    bounded sample: the first `rows` cells of the workload, all bins."""
    import oracle
    from snapatac2_b200 import synth
    n, m, nnz_row, K, k = CONFIGS[cfg_name]
    rows = min(rows, n)
    spec = synth.make_spec(n, m, nnz_row, K, seed=0)
    X = synth.generate_csr(spec, 0, rows, dtype=np.float64)
    counter = [0]
    t0 = time.perf_counter()
    oracle.spectral_embedding(X, None, min(k, rows - 1), 0, counter=counter)
    dt = time.perf_counter() - t0
    return rows / dt, dt, counter[0], rows, X.nnz


def cpu_slab_estimate(cfg_name, matvecs, device=0, rows=100_000, timed_matvecs=3):
    """the oracle's preparation (IDF, normalisation, degrees) and a few
    applications of its operator f(v) timed on a 100k-row slab of the workload (generated on the device and
    exported -- the host generator needs minutes for it), scaled by n / rows and by the ARPACK mat-vec count of
    the small full run.  Labelled extrapolated.  Needs a GPU only to synthesise the slab."""
    import oracle
    from snapatac2_b200 import Engine, synth
    n, m, nnz_row, K, k = CONFIGS[cfg_name]
    rows = min(rows, n)
    spec = synth.make_spec(n, m, nnz_row, K, seed=0)
    with Engine(device) as gen:
        gen.generate(spec, row0=0, n_local=rows)
        X = gen.export_csr().astype(np.float64)
    t0 = time.perf_counter()
    w = oracle.idf(X)
    xhat = oracle.normalize(X, w)
    xt, dinv, _, _ = oracle.operator_pieces(xhat)
    t_prep = time.perf_counter() - t0
    v = np.random.RandomState(0).rand(rows)
    t0 = time.perf_counter()
    for _ in range(timed_matvecs):
        v = xt @ (v.T @ xt).T - dinv * v
        v /= np.linalg.norm(v)
    t_mv = (time.perf_counter() - t0) / timed_matvecs
    total = t_prep + matvecs * t_mv
    return {"rows": rows, "nnz": int(X.nnz), "prepare_s": t_prep, "matvec_s": t_mv, "matvecs_assumed": int(matvecs),
            "cells_per_s": rows / total,
            "note": "extrapolated: (oracle prepare + ARPACK mat-vec count of the small full run x measured f(v)) on a "
                    f"{rows}-row slab; cost is linear in cells"}


def thread_info():
    info = {"cpu_count": os.cpu_count(), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}
    try:
        from threadpoolctl import threadpool_info
        info["blas_threads"] = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        pass
    return info


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (the
    oracle restatement -- the Rust extension cannot be built here) on this box's
    host cores.  scipy's sparse kernels are single-threaded, exactly as in the
    reference, whose mat-vec is the same scipy call (embedding.rs:162-163)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = max(0, min(args.warmup, 1))
    for _ in range(warm):
        cpu_oracle_step(args.config)
    vals, times, mv = [], [], 0
    for _ in range(steps):
        v, dt, mv, rows, nnz = cpu_oracle_step(args.config)
        vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    ti = thread_info()
    slab = None
    if CONFIGS[args.config][0] >= 100_000 and not args.no_slab:
        try:
            slab = cpu_slab_estimate(args.config, mv, device=int(os.environ.get("LOCAL_RANK", "0")))
        except Exception as e:
            slab = {"error": str(e)[:200]}
    sample = (f"first {rows} cells of the workload (all {CONFIGS[args.config][1]} bins, nnz={nnz}); full oracle run incl. "
              f"ARPACK ({mv} mat-vecs); cells/s = sample cells / wall time (cost is linear in cells)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIG_TEXT[args.config], "config": args.config},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": 1, "kind": "port", "sample": sample,
                         "threads": ti, "slab_100k": slab},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as td

    from snapatac2_b200 import Engine, dist, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world,
                              device_id=torch.device("cuda", local_rank))

    n, m, nnz_row, K, k = CONFIGS[args.config]
    spec = synth.make_spec(n, m, nnz_row, K, seed=0)
    bounds = dist.equal_row_splits(n, world)
    row0, row1 = int(bounds[rank]), int(bounds[rank + 1])
    n_local = row1 - row0

    eng = Engine(local_rank)
    eng.set_spmm_mode(args.spmm)
    if args.block:
        eng.set_block(args.block)
    dist.attach_engine_comm(eng)
    ext_stream = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local_rank))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    t_gen = time.perf_counter()
    eng.generate(spec, row0=row0, n_local=n_local)
    _, _, nnz_local = eng.shape()
    t_gen = time.perf_counter() - t_gen

    # the step's result lands in caller-provided pinned host memory (Engine.eigsh(out_evecs=...)): one DMA per
    # step; a pageable destination (what tl.spectral returns, timed in `e2e`) goes through the staging team instead
    try:
        evecs_t = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
    except Exception:
        evecs_t = torch.empty((n_local, k), dtype=torch.float64)
    evecs = evecs_t.numpy()

    call_ms = {"prepare": 0.0, "eigsh": 0.0}

    def step():
        t_a = time.perf_counter()
        eng.prepare(want_outputs=False)
        t_b = time.perf_counter()
        out = eng.eigsh(k, seed=0, tol=args.tol, block=args.block, out_evecs=evecs)
        t_c = time.perf_counter()
        call_ms["prepare"], call_ms["eigsh"] = 1e3 * (t_b - t_a), 1e3 * (t_c - t_b)
        return out

    # the clock sampler starts before the warm-up (nvidia-smi needs ~1 s to come up); only the samples
    # taken inside the timed region are reported (all samples under load if the region is too short)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()

    # ---- timed region: K steps, device events on the library's stream, max over ranks
    st0 = eng.stats()
    launches0 = st0["kernel_launches"]
    sync_all()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext_stream)
    t0 = time.perf_counter()
    step_wall = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        evals, _ = step()
        step_wall.append(round(1e3 * (time.perf_counter() - t_s), 2))
    ev1.record(ext_stream)
    ev1.synchronize()
    wall = time.perf_counter() - t0
    sync_all()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.stats()["kernel_launches"] - launches0
    stats = eng.stats()

    tmax = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
    total_ms = float(tmax[0].item())
    wall_ms = float(tmax[1].item())
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step / 1e3)

    # ---- roofline of the SpMM pair (the dominant kernels), timed alone with L2 flushed
    b = stats["block"] or 8
    p1, cm, p2 = eng.operator_time(b=b, iters=args.op_iters, flush_l2=True)
    rt = torch.tensor([p1, cm, p2], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(rt, op=td.ReduceOp.MAX)
    p1, cm, p2 = [float(x) for x in rt.tolist()]
    bytes_p1 = 4 * nnz_local + 8 * (m + 1) + 4 * b * n_local + 4 * b * m + 4 * m          # idx, ptr, read rV, write W, w^2
    bytes_p2 = 4 * nnz_local + 8 * (n_local + 1) + 4 * b * m + 3 * 4 * b * n_local + 8 * n_local  # idx, ptr, read W, V/Y, r/dinv
    peak, peak_src = peak_hbm()
    achieved = (bytes_p1 + bytes_p2) / ((p1 + p2) * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "spmm_traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            traffic = tj.get("dram_bytes_per_application")
            if traffic is not None and tj.get("nnz_of_capture"):
                # the ncu capture was taken on one GPU at C3; DRAM traffic is proportional to the stored entries
                traffic = float(traffic) * nnz_local / float(tj["nnz_of_capture"])
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "kernel": (f"sell_spmm_kernel<{b}> (shared-memory tiled)" if stats.get("spmm_tiled") else f"gather_rows_kernel<{b}> (CSR, L2 gather)")
        + " pass1+pass2 = one operator application",
        "algorithmic_bytes": bytes_p1 + bytes_p2, "ms_pass1": p1, "ms_pass2": p2, "ms_allreduce": cm,
        "frac_pass1": bytes_p1 / (p1 * 1e-3) / 1e9 / peak, "frac_pass2": bytes_p2 / (p2 * 1e-3) / 1e9 / peak,
        "frac_nominal_8TBs": achieved / 8000.0, "per_gpu": True,
        # the tiled format stores a 16-bit tile-local column per entry: the kernel moves about half the
        # algorithmic (CSR int32) bytes, so `frac` can exceed what a 4 B/entry stream could reach
        "stored_index_bytes_per_entry": 2 if stats.get("spmm_tiled") else 4,
    }

    # ---- the reference's own answer to this size: the Nystrom path (sample_size), timed beside the full operator
    nystrom = None
    if args.nystrom > 0:
        from snapatac2_b200 import tl
        full_evecs = evecs.copy()
        t0 = time.perf_counter()
        v_n, u_n = tl.spectral_embedding_nystrom(eng, None, None, k, args.nystrom, False, 20000, tol=args.tol, block=args.block)
        ev_n, q_n = tl.orthogonalize(v_n, u_n)
        torch.cuda.synchronize()
        dt_n = time.perf_counter() - t0
        tn = torch.tensor([dt_n], dtype=torch.float64, device="cuda")
        if world > 1:
            td.all_reduce(tn, op=td.ReduceOp.MAX)
        # principal cosines between the two k-dimensional embeddings (both have orthonormal columns over all cells)
        q_n = np.real(q_n)
        q_n = q_n / np.sqrt(dist.allreduce_array(np.sum(q_n * q_n, axis=0), "sum"))[None, :]
        cross = dist.allreduce_array(full_evecs.T @ q_n, "sum")
        pc = np.linalg.svd(cross, compute_uv=False)
        nystrom = {"sample_size": int(args.nystrom), "chunk_size": 20000, "seconds": float(tn.item()),
                   "cells_per_s": n / float(tn.item()), "evals_head": [float(x) for x in np.real(ev_n[:4])],
                   "principal_cosines_vs_full": {"max": float(pc.max()), "median": float(np.median(pc)), "min": float(pc.min()),
                                                 "n_above_0.9": int((pc > 0.9).sum()), "k": int(k)},
                   "note": "tl.spectral_embedding_nystrom + orthogonalize on the resident shards (landmark draw: numpy default_rng(2023))"}
        eng.prepare(want_outputs=False)      # the stats below describe the full path again
        eng.eigsh(k, seed=0, tol=args.tol, block=args.block, out_evecs=evecs)
        stats = eng.stats()

    # ---- e2e: the reference-facing plugin call, tl.spectral(adata, n_comps, features=None), on an in-memory
    #      AnnData whose X is an ordinary scipy CSR in pageable host memory (int64 indices as soon as the
    #      shard holds more than 2^31 entries, as scipy stores them; float32 values, all ones).  Everything
    #      is inside the timed region: host-side scan of the values, narrowing of the indices into pinned
    #      staging, the H2D copies, prepare, eigsh, the eigenvectors back in a fresh numpy array.
    e2e = None
    e2e_skipped = None
    knn = None
    if not args.no_e2e:
        import psutil
        need = 16 * nnz_local * world          # int64 indices + float32 values + the int32 export, all ranks of this box
        if need > 0.7 * psutil.virtual_memory().total:
            e2e_skipped = f"host arrays of all {world} ranks ({need / 1e9:.0f} GB) exceed 70% of host memory"
    if not args.no_e2e and e2e_skipped is None:
        import scipy.sparse as sp
        from concurrent.futures import ThreadPoolExecutor
        from snapatac2_b200 import MiniAnnData, tl

        t_build = time.perf_counter()
        np_ptr = np.empty(n_local + 1, dtype=np.int64)
        np_idx32 = np.empty(max(1, nnz_local), dtype=np.int32)
        eng.export_arrays(indptr=np_ptr, indices=np_idx32[:nnz_local])
        idx_dtype = np.int64 if nnz_local > np.iinfo(np.int32).max else np.int32   # scipy's own rule
        np_idx = np.empty(nnz_local, dtype=idx_dtype)
        np_val = np.empty(nnz_local, dtype=np.float32)
        cuts = np.linspace(0, nnz_local, 65).astype(np.int64)

        def fill(i):
            lo, hi = int(cuts[i]), int(cuts[i + 1])
            np.copyto(np_idx[lo:hi], np_idx32[lo:hi])
            np_val[lo:hi] = 1.0

        with ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 8) // max(1, world)))) as ex:
            list(ex.map(fill, range(64)))
        del np_idx32
        X = sp.csr_matrix((1, 1), dtype=np.float32)
        # assemble without scipy's validating constructor (it would copy the 40 GB index array once more)
        X.data, X.indices, X.indptr = np_val, np_idx, np_ptr.astype(idx_dtype)
        X._shape = (n_local, m)
        adata = MiniAnnData(X)
        t_build = time.perf_counter() - t_build

        def e2e_step():
            return tl.spectral(adata, n_comps=k, features=None, random_state=0, inplace=False, engine=eng,
                               tol=args.tol, block=args.block)

        e2e_step()   # warm-up (allocates the pinned staging ring)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ev_e2e, emb_e2e = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st_e2e = eng.stats()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            td.all_reduce(tt, op=td.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": n * args.e2e_steps / dt, "unit": "cells/s", "ms_per_step": 1e3 * dt / args.e2e_steps,
               "h2d_bytes_per_step": int(st_e2e["bytes_h2d"]), "d2h_bytes_per_step": int(emb_e2e.nbytes + ev_e2e.nbytes),
               "steps": args.e2e_steps, "pinned_host": False, "per_rank_bytes": True,
               "api": "tl.spectral(MiniAnnData(scipy.sparse.csr_matrix), n_comps=30, features=None, inplace=False)",
               "host_arrays": {"indices": str(np_idx.dtype), "indptr": str(X.indptr.dtype), "data": "float32 (all ones)",
                               "bytes": int(np_idx.nbytes + np_val.nbytes + X.indptr.nbytes), "memory": "pageable (numpy)"},
               "host_threads": int(st_e2e["host_threads"]), "ms_load": st_e2e["ms_load"],
               "ms_prepare": st_e2e["ms_prepare_wall"], "ms_eigsh": st_e2e["ms_eigsh"],
               "note": "values are scanned on the host (all ones -> not shipped), on background threads while the GPU already "
                       "prepares and solves; h2d bytes = indptr + the indices as 16-bit differences with a side list "
                       "(csrc/delta_encode.h; 4 bytes per index with SNAPB200_NO_DELTA=1)",
               "setup_s": t_build}
        del adata, X, np_idx, np_val
        # ---- the consumer next to the path, on what the call above returned (outside every timed region of the
        #      headline metric; one GPU only): pp.knn = exact 50-nearest-neighbour graph of the embedding
        if world == 1 and not args.no_knn and n <= 2_000_000:
            try:
                from snapatac2_b200 import pp
                t0 = time.perf_counter()
                adj = pp.knn(emb_e2e, n_neighbors=50, engine=eng)
                dt_knn = time.perf_counter() - t0
                knn = {"api": "pp.knn(X_spectral of the e2e call, n_neighbors=50)", "points": int(emb_e2e.shape[0]),
                       "dims": int(emb_e2e.shape[1]), "ms": 1e3 * dt_knn, "ms_device": eng.stats()["ms_knn"],
                       "cells_per_s": emb_e2e.shape[0] / dt_knn, "nnz": int(adj.nnz)}
                del adj
            except Exception as exc:          # never let the side leg take the headline line down
                knn = {"error": str(exc)[:300]}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, mv, rows, nnz_s = cpu_oracle_step(args.config)
        cpu = {"value": v, "unit": "cells/s", "cores": 1, "kind": "port",
               "sample": f"first {rows} cells of the workload (nnz={nnz_s}), full oracle run incl. ARPACK ({mv} mat-vecs) in {dt:.1f} s; "
                         f"scipy sparse kernels are single-threaded as in the reference",
               "threads": thread_info()}
        if n >= 100_000 and not args.no_slab:
            try:
                cpu["slab_100k"] = cpu_slab_estimate(args.config, mv, device=local_rank)
            except Exception as e:      # the baseline is a report, never a reason to lose the line
                cpu["slab_100k"] = {"error": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_TEXT[args.config], "config": args.config, "n_cells": n, "n_bins": m,
                       "nnz_per_gpu": nnz_local, "n_comps": k, "block": b, "tol": args.tol or 1e-5,
                       "parallelism": f"rows/{world}", "l2": "inputs larger than L2 (index stream >> 126 MB)"
                       if nnz_local * 4 > (256 << 20) else "inputs fit L2; operator_time flushes L2 between iterations"},
            "clocks": clocks, "e2e": e2e if e2e is not None else ({"skipped": e2e_skipped} if e2e_skipped else None),
            "nystrom": nystrom, "knn": knn, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "timing": {"device_ms_total": total_ms, "wall_ms_total": wall_ms, "generate_s": t_gen,
                       "last_step_call_ms": call_ms, "step_wall_ms": step_wall,
                       "pool_mallocs_in_timed_region": stats["pool_mallocs"] - st0["pool_mallocs"],
                       "pool_ms_in_timed_region": stats["ms_pool"] - st0["ms_pool"]},
            "solver": {kk: stats[kk] for kk in ("n_ops", "n_restarts", "basis_cols", "max_residual", "ms_transpose",
                                                "ms_prepare", "ms_format", "ms_eigsh", "ms_spmm", "ms_ortho", "ms_comm", "ms_host",
                                                "spmm_tiled", "ms_prepare_wall", "ms_pool", "pool_mallocs", "converged",
                                                "n_spec_ops", "ms_d2h", "fused_allreduce")},
            "evals_head": [float(x) for x in evals[:4]],
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()


def run_multiview(args):
    """configs[4]: snap.tl.multi_spectral co-embedding of an ATAC tile view and an RNA count view of the same
    cells.  A step = everything multi_spectral does after the views are on the device: per-view prepare (IDF,
    norms, both tiled copies, degrees), the frobenius_norm normaliser on 2000 sampled rows, the combination
    and the block-Lanczos solve on the virtual column concatenation.  `e2e` = tl.multi_spectral on host CSRs."""
    import torch
    import torch.distributed as td
    import scipy.sparse as sp

    from snapatac2_b200 import Engine, MiniAnnData, dist, synth, tl

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    n, m, nnz_row, K, k = CONFIGS[args.config]
    m2, nnz_row2 = C5_RNA
    spec_a = synth.make_spec(n, m, nnz_row, K, seed=0)
    spec_r = synth.make_spec(n, m2, nnz_row2, K, seed=0)          # same planted labels (keyed by seed, row)
    bounds = dist.equal_row_splits(n, world)
    row0, row1 = int(bounds[rank]), int(bounds[rank + 1])
    n_local = row1 - row0

    eng = Engine(local_rank)
    if args.block:
        eng.set_block(args.block)
    dist.attach_engine_comm(eng)
    views = tl._view_engines(eng, 2)
    ext_stream = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local_rank))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    # the RNA view's counts (1 + column % 3: a function of the column, shard independent) are attached on the host
    views[1].generate(spec_r, row0=row0, n_local=n_local)
    R = views[1].export_csr()
    R.data = (1.0 + (R.indices % 3)).astype(np.float32)
    views[0].generate(spec_a, row0=row0, n_local=n_local)
    views[1].load_csr(R, n_global=n, row0=row0)
    nnz_a, nnz_r = views[0].shape()[2], views[1].shape()[2]
    rows = np.sort(np.random.RandomState(2023).choice(n, min(2000, n), replace=False))
    mine = rows[(rows >= row0) & (rows < row0 + n_local)] - row0
    try:
        evecs_t = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
    except Exception:
        evecs_t = torch.empty((n_local, k), dtype=torch.float64)
    evecs = evecs_t.numpy()
    parts = {}

    def step():
        t_a = time.perf_counter()
        norms = []
        for v in views:
            v.set_feature_weights(None)
            v.prepare(want_outputs=False)
            norms.append(float(np.sqrt(v.view_frobenius(mine) - len(rows))))
        ws = [1.0 / nrm for nrm in norms]
        scales = [float(np.sqrt(w / sum(ws))) for w in ws]
        eng.combine_views(views, scales)
        t_b = time.perf_counter()
        out = eng.eigsh(k, seed=0, tol=args.tol, block=args.block, out_evecs=evecs)
        parts["prepare"], parts["eigsh"], parts["norms"] = 1e3 * (t_b - t_a), 1e3 * (time.perf_counter() - t_b), norms
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    launches0 = sum(v.stats()["kernel_launches"] for v in views)
    sync_all()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext_stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        evals, _ = step()
    ev1.record(ext_stream)
    ev1.synchronize()
    wall = time.perf_counter() - t0
    sync_all()
    clocks = sampler.stop()
    launches = sum(v.stats()["kernel_launches"] for v in views) - launches0
    stats = eng.stats()
    tmax = torch.tensor([ev0.elapsed_time(ev1), wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
    ms_per_step = float(tmax[0].item()) / args.steps

    # ---- e2e: tl.multi_spectral on host CSRs (ATAC: float32 ones; RNA: float32 counts)
    e2e = None
    if not args.no_e2e:
        A = views[0].export_csr()
        ad_a, ad_r = MiniAnnData(A), MiniAnnData(R)

        def e2e_step():
            return tl.multi_spectral([ad_a, ad_r], n_comps=k, features=None, engine=eng)

        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ev_e, emb_e = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            td.all_reduce(tt, op=td.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": n * args.e2e_steps / dt, "unit": "cells/s", "ms_per_step": 1e3 * dt / args.e2e_steps,
               "h2d_bytes_per_step": int(sum(v.stats()["bytes_h2d"] for v in views)),
               "d2h_bytes_per_step": int(emb_e.nbytes + ev_e.nbytes), "steps": args.e2e_steps, "per_rank_bytes": True,
               "api": "tl.multi_spectral([MiniAnnData(atac_csr), MiniAnnData(rna_csr)], n_comps=30, features=None)"}

    if rank == 0:
        line = {
            "metric": "snap.tl.multi_spectral cells/s", "value": n / (ms_per_step / 1e3), "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_TEXT[args.config], "config": args.config, "n_cells": n, "views": [[n, m], [n, m2]],
                       "nnz_per_gpu": [int(nnz_a), int(nnz_r)], "n_comps": k, "block": stats["block"],
                       "parallelism": f"rows/{world}", "l2": "inputs larger than L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None,
            "timing": {"wall_ms_total": float(tmax[1].item()), "last_step_call_ms": {kk: parts[kk] for kk in ("prepare", "eigsh")}},
            "view_norms": parts["norms"],
            "solver": {kk: stats[kk] for kk in ("n_ops", "n_restarts", "basis_cols", "max_residual", "ms_eigsh", "ms_spmm",
                                                "ms_ortho", "ms_comm", "ms_host", "converged")},
            "evals_head": [float(x) for x in evals[:4]],
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default=os.environ.get("SNAPB200_BENCH_CONFIG", "c3"))
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--spmm", choices=["auto", "csr", "tiled"], default="auto")
    ap.add_argument("--tol", type=float, default=0.0)
    ap.add_argument("--op-iters", type=int, default=5)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="skip the 100k-row slab leg of the CPU baseline")
    ap.add_argument("--no-knn", action="store_true", help="skip the pp.knn side leg on the embedding of the e2e call")
    ap.add_argument("--nystrom", type=int, default=0,
                    help="also time the Nystrom path (reference: sample_size) with this many landmarks on the same resident data")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config.startswith("c5"):
        run_multiview(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
