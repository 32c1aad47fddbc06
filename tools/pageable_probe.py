#!/usr/bin/env python
"""Throughput of piquant_quantize on PAGEABLE host tensors (what a caller of the reference passes) against the number of
host copy workers (PIQUANT_COPY_THREADS caps them; the context's num_threads asks for them).  Development tool."""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

os.environ["PIQUANT_COPY_THREADS"] = "64"
import piquant  # noqa: E402
from piquant import DataType as D, RoundMode  # noqa: E402


def main() -> None:
    n = 1 << 28
    x = torch.empty(n, dtype=torch.float32).uniform_(-1, 1)        # pageable
    q = torch.empty(n, dtype=torch.uint8)
    xp, qp = x.pin_memory(), q.pin_memory()
    print(f"host cores: {os.cpu_count()}, numel = {n}")
    for threads in (1, 2, 4, 6, 8, 12, 16, 24, 32):
        if threads > (os.cpu_count() or 1):
            continue
        ctx = piquant.Context(threads)
        for src, dst, name in ((x, q, "pageable"), (xp, qp, "pinned  ")):
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter()
                ctx.quantize_ptr(src.data_ptr(), D.F32, dst.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
                best = min(best, time.perf_counter() - t0)
            print(f"  num_threads {threads:3d}  {name}  {n / best / 1e9:7.2f} Gelem/s   H2D {4 * n / best / 1e9:6.1f} GB/s", flush=True)
        del ctx


if __name__ == "__main__":
    main()
