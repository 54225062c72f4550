"""pp.knn on a synthetic embedding: the exact neighbour graph on the GPU vs an exact kd-tree on the host.

    python scripts/bench_knn.py [--n 1000000] [--dims 30] [--k 50] [--steps 3] [--cpu-queries 4000]

Prints one JSON line: cells/s through the public call (host points in, host CSR arrays out: `e2e`) and on
the device alone (`value`, CUDA events around centring + scan), the fp32-FMA roofline of the filter kernel
(n^2 x DP fused multiply-adds; peak = SMs x 128 lanes x 2 x SM clock), and the CPU baseline: the oracle's
scipy cKDTree search (all host cores) timed on a bounded sample of query rows, tree construction included
pro rata.  The sampled rows double as a full-size parity check: indices and distances must be identical.
"""

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def blobs(n, d, seed=0, n_centres=24):
    rng = np.random.default_rng(seed)
    centres = rng.normal(scale=0.02, size=(n_centres, d))
    lab = rng.integers(0, n_centres, size=n)
    return centres[lab] + rng.normal(size=(n, d)) * rng.uniform(0.002, 0.01, size=n_centres)[lab][:, None]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dims", type=int, default=30)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-queries", type=int, default=4000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import torch
    from snapatac2_b200 import Engine, pp
    eng = Engine(0)
    P = blobs(args.n, args.dims)
    n, d, k = args.n, args.dims, args.k
    DP = 8 if d <= 8 else 16 if d <= 16 else 32 if d <= 32 else 64

    dev_ms, wall_ms = [], []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        adj = pp.knn(P, n_neighbors=k, engine=eng)
        t1 = time.perf_counter()
        if s >= args.warmup:
            wall_ms.append((t1 - t0) * 1e3)
            dev_ms.append(eng.stats()["ms_knn"])
    dev = float(np.mean(dev_ms))
    wall = float(np.mean(wall_ms))
    props = torch.cuda.get_device_properties(0)
    try:
        sm_mhz = torch.cuda.clock_rate()
    except Exception:
        sm_mhz = None
    peak = props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12          # TFLOP/s at the B200's 1965 MHz max SM clock
    flops = 2.0 * n * (-(-n // 128) * 128) * DP
    line = {
        "metric": "knn_cells_per_s", "value": n / (dev * 1e-3), "unit": "cells/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev, "higher_is_better": True, "dtype": "f32 filter + f64 exact distances",
        "data": "synthetic", "vs_baseline": None,
        "config": {"workload": f"exact kNN graph, {n} points x {d} dims (24 blobs), k={k}", "padded_dims": DP},
        "e2e": {"value": n / (wall * 1e-3), "unit": "cells/s", "ms_per_step": wall,
                "h2d_bytes_per_step": int(P.nbytes), "d2h_bytes_per_step": int(adj.indices.nbytes // 2 + adj.data.nbytes),
                "api": "pp.knn(ndarray, n_neighbors=k)  (includes building the scipy CSR)"},
        "roofline": {"bound": "fp32 FMA pipe", "achieved": flops / (dev * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": flops / (dev * 1e-3) / 1e12 / peak, "traffic": None,
                     "note": "filter flops only (n x n_padded x DP FMAs); peak = SMs x 128 x 2 x 1.965 GHz"},
        "gpu_launches": 3 * args.steps, "sm_clock_mhz_now": sm_mhz,
    }
    if not args.no_cpu:
        import oracle
        rng = np.random.default_rng(1)
        rows = np.sort(rng.choice(n, size=min(args.cpu_queries, n), replace=False))
        from scipy.spatial import cKDTree
        t0 = time.perf_counter()
        cKDTree(P)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = oracle.knn.nearest_neighbour_graph_kdtree(P, k, rows=rows, workers=-1)
        t_query = time.perf_counter() - t0 - t_build        # (the helper builds its own tree)
        est = t_build + max(t_query, 1e-9) * n / rows.size
        line["cpu_baseline"] = {"value": n / est, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"cKDTree over all {n} points ({t_build:.1f} s) + queries for {rows.size} rows "
                                          f"({t_query:.2f} s, workers=-1), query time scaled to {n} rows"}
        got = adj[rows]
        same = bool(np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data))
        line["parity_on_sample"] = {"rows": int(rows.size), "identical": same}
        if not same:
            bad = np.flatnonzero(np.any(got.indices.reshape(rows.size, -1) != want.indices.reshape(rows.size, -1), axis=1))
            line["parity_on_sample"]["rows_differing"] = int(bad.size)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
