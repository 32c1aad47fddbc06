#!/usr/bin/env python
"""Builds pi-quant_b200/piquant/libpiquant.so with nvcc for sm_100a (cross-compiles without a GPU).

    python pi-quant_b200/build.py [--force] [--verbose]

The library is built IN-TREE, next to the Python package that dlopens it (the reference ships
libpiquant.so inside its package directory too, reference python/src/piquant/_bootstrap.py:91-93),
with a statically linked CUDA runtime so that it loads in any process, with or without PyTorch.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "piquant" / "libpiquant.so"
OBJ = HERE / "build"
SOURCES = ["context.cu", "quantize.cu", "quantize_tma.cu", "dequantize.cu", "dequantize_tma.cu", "requantize.cu", "minmax.cu", "reduce_sum.cu"]
HEADERS = [CSRC / "pq_device.cuh", CSRC / "pq_kernels.h", CSRC / "quantize_common.cuh", CSRC / "dequantize_common.cuh", CSRC / "pq_tma.cuh", CSRC / "pq_reduce.cuh",
           HERE.parent / "include" / "piquant.h", HERE.parent / "include" / "piquant_cuda.h"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",                 # every fma in the kernels is explicit; nothing may be contracted behind our back
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
    "-cudart", "static",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; libpiquant.so cannot be built")
    return exe


def stale() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: Path = OUT) -> Path:
    """`defines` / `out` build an experimental variant next to the product library (A/B runs on the GPU box)."""
    if not force and not defines and not stale():
        return OUT
    obj_dir = OBJ if not defines else HERE / ("build_" + "_".join(d.split("=")[0] for d in defines))
    obj_dir.mkdir(exist_ok=True)
    cc = nvcc()

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src + ".o")
        cmd = [cc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = out.with_suffix(".so.tmp")
    link = [cc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-o", str(tmp), *map(str, objs), "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    tmp.replace(out)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--define", action="append", default=[], help="extra -D for an experimental variant, e.g. PQ_BF16_FHADD")
    ap.add_argument("--out", default=str(OUT), help="output path of the variant (default: the product library)")
    a = ap.parse_args()
    print(build(a.force, a.verbose, tuple(a.define), Path(a.out)))
