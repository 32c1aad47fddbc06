#!/usr/bin/env python
"""Size sweep with ROTATING buffers (the working set always exceeds the 126 MB L2), direct vs TMA kernels.
Development tool: decides the size threshold of the per-cell variant choice.  python tools/sizebench.py"""
from __future__ import annotations

import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
from piquant import DataType as D, ReduceOp, RoundMode  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6533.8
SIZES = [1 << 20, 1 << 22, 1 << 24, 27_264_000, 1 << 26, 100_000_000, 1 << 28, 1_000_000_000]


def bench(fn_of_i, k: int, reps: int) -> float:
    for i in range(min(k, 3) * 2):
        fn_of_i(i % k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn_of_i(i % k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main() -> None:
    global SIZES
    if len(sys.argv) > 2 and sys.argv[1] == "--sizes":
        SIZES = [int(v) for v in sys.argv[2].split(",")]
    torch.cuda.set_device(0)
    ctx = piquant.Context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    print(f"{'cell':28s} {'numel':>11s} {'direct us':>10s} {'tma us':>10s} {'direct %':>9s} {'tma %':>7s}  (of measured {PEAK:.0f} GB/s)")
    for n in SIZES:
        cells = [
            ("quant f32->u8", torch.float32, D.F32, D.UINT8, 5.0, n, "q"),
            ("quant bf16->u4", torch.bfloat16, D.BF16, D.UINT4, 2.5, (n + 1) // 2, "q"),
            ("dequant u4->bf16 set", torch.bfloat16, D.UINT4, D.BF16, 2.5, (n + 1) // 2, "d"),
            ("dequant u8->f32 add", torch.float32, D.UINT8, D.F32, 9.0, n, "a"),
        ]
        for name, fdt, din, dout, bpe, qbytes, kind in cells:
            k = max(3, int(600e6 // (bpe * n)) + 1)
            k = min(k, 64)
            fl = [torch.empty(n, dtype=fdt, device="cuda").uniform_(-1, 1) for _ in range(k)]
            qs = [torch.randint(0, 255, (qbytes,), dtype=torch.uint8, device="cuda") for _ in range(k)]
            reps = max(10, min(400, int(0.2 / (bpe * n / 6e12))))
            row = []
            for variant in (1, 2):
                ctx.set_kernel_variant(variant)
                if kind == "q":
                    t = bench(lambda i: ctx.quantize_ptr(fl[i].data_ptr(), din, qs[i].data_ptr(), dout, n, 2 / 255, 8, RoundMode.NEAREST), k, reps)
                else:
                    op = ReduceOp.ADD if kind == "a" else ReduceOp.SET
                    t = bench(lambda i: ctx.dequantize_ptr(qs[i].data_ptr(), din, fl[i].data_ptr(), dout, n, 2 / 255, 8, op), k, reps)
                row.append(t)
            d, m = row
            print(f"{name:28s} {n:11d} {d*1e6:10.2f} {m*1e6:10.2f} {bpe*n/d/1e9/PEAK*100:8.1f}% {bpe*n/m/1e9/PEAK*100:6.1f}%", flush=True)
            del fl, qs
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
