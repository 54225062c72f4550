"""Build libsnapb200.so in-tree with nvcc for sm_100a (B200).

``python -m snapatac2_b200.build`` or ``build()`` from ``__graft_entry__``.
Each ``csrc/*.cu`` is compiled to an object (in parallel) and linked into
``snapatac2_b200/libsnapb200.so``.  nvcc cross-compiles without a GPU.
"""

from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "_build"
LIB = HERE / "libsnapb200.so"

SOURCES = ["api.cu", "ingest.cu", "pool.cu", "comm.cu", "util.cu", "synth.cu", "prep.cu", "transpose_tiled.cu", "spmm.cu", "sell_build.cu", "spmm_tiled.cu", "dense.cu", "lanczos.cu", "knn.cu"]


def _nccl_paths():
    import importlib.util
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec is None or not spec.submodule_search_locations:
        raise RuntimeError("nvidia.nccl (NCCL headers/lib) not found in this environment")
    root = Path(list(spec.submodule_search_locations)[0])
    return root / "include", root / "lib"


def _flags():
    inc, _ = _nccl_paths()
    return [
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-Xcompiler", "-mavx2", "-Xcompiler", "-pthread", "-I", str(inc), "-I", str(HERE.parent / "include"),
    ]


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(_flags()).encode())
    for p in [src, *sorted(CSRC.glob("*.cuh")), HERE.parent / "include" / "snapb200.h"]:
        h.update(p.read_bytes())
    return h.hexdigest()


def _compile(name: str, verbose: bool) -> Path:
    src = CSRC / name
    obj = BUILD / (name + ".o")
    stamp_file = BUILD / (name + ".stamp")
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    cmd = ["nvcc", *_flags(), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    stamp_file.write_text(stamp)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    if force:
        for p in BUILD.glob("*.stamp"):
            p.unlink()
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    _, nccl_lib = _nccl_paths()
    newest = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime >= newest and not force:
        return LIB
    cmd = ["nvcc", "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-L", str(nccl_lib), "-l:libnccl.so.2", "-Xlinker", f"-rpath={nccl_lib}", "-Xlinker", "--no-as-needed"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
