#!/usr/bin/env python
"""Times every cell of the kernel matrix through the C ABI with CUDA events (development tool; the
judged numbers come from bench.py).  Usage: python tools/cellbench.py [--numel N] [--variants 1,2] [--reps R]"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
from piquant import DataType as D, ReduceOp, RoundMode  # noqa: E402

PEAK = 6533.8
try:
    PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass

FL = {D.F32: torch.float32, D.BF16: torch.bfloat16}


def time_fn(fn, reps: int, warm: int = 3) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=1_000_000_000)
    ap.add_argument("--variants", default="1,2")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cells", default="all")
    a = ap.parse_args()
    n = a.numel
    torch.cuda.set_device(0)
    ctx = piquant.Context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_stochastic_threshold(0.37)
    g = torch.Generator(device="cuda").manual_seed(0)
    xf = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-1, 1, generator=g)
    xb = xf.to(torch.bfloat16)
    q = torch.empty(n, dtype=torch.uint8, device="cuda")
    acc = torch.zeros(n, dtype=torch.float32, device="cuda")
    accb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    rows = []
    # bring the GPU to its steady (power-capped) clocks first: the first ~100 ms after idle measure ~5 % low
    import time
    t_end = time.perf_counter() + 1.5
    while time.perf_counter() < t_end:
        for _ in range(20):
            ctx.quantize_ptr(xf.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
        torch.cuda.synchronize()

    def report(name, bytes_per_elem, t, variant):
        gbs = bytes_per_elem * n / t / 1e9
        rows.append((name, variant, n / t / 1e9, gbs, gbs / PEAK))
        print(f"{name:44s} v{variant}  {t*1e3:9.3f} ms  {n/t/1e9:9.1f} Gelem/s  {gbs:8.1f} GB/s  {gbs/PEAK*100:6.1f}% of measured {PEAK:.0f}", flush=True)

    for variant in [int(v) for v in a.variants.split(",")]:
        ctx.set_kernel_variant(variant)
        for din, x in ((D.F32, xf), (D.BF16, xb)):
            isz = 4 if din == D.F32 else 2
            for dq in (D.UINT8, D.UINT4, D.UINT2):
                for mode in (RoundMode.NEAREST, RoundMode.STOCHASTIC, RoundMode.STOCHASTIC_PER_ELEMENT):
                    name = f"quant {din.name}->{dq.name} {mode.name.lower()}"
                    if a.cells != "all" and a.cells not in name:
                        continue
                    scale, zp = (2.0 / ((1 << dq.bit_size) - 1), (1 << dq.bit_size) // 2)
                    t = time_fn(lambda: ctx.quantize_ptr(x.data_ptr(), din, q.data_ptr(), dq, n, scale, zp, mode), a.reps)
                    report(name, isz + dq.bit_size / 8, t, variant)
    for variant in [int(v) for v in a.variants.split(",")]:
        ctx.set_kernel_variant(variant)
        for dq in (D.UINT8, D.UINT4, D.UINT2):
            for dout, o in ((D.F32, acc), (D.BF16, accb)):
                osz = 4 if dout == D.F32 else 2
                for op in (ReduceOp.SET, ReduceOp.ADD):
                    name = f"dequant {dq.name}->{dout.name} {op.name.lower()}"
                    if a.cells != "all" and a.cells not in name:
                        continue
                    scale, zp = (2.0 / ((1 << dq.bit_size) - 1), (1 << dq.bit_size) // 2)
                    t = time_fn(lambda: ctx.dequantize_ptr(q.data_ptr(), dq, o.data_ptr(), dout, n, scale, zp, op), a.reps)
                    report(name, dq.bit_size / 8 + osz * (2 if op == ReduceOp.ADD else 1), t, variant)
    ctx.set_kernel_variant(0)
    for din, x in ((D.F32, xf), (D.BF16, xb)):
        name = f"minmax->params {din.name}"
        if a.cells != "all" and a.cells not in name:
            continue
        fn = ctx.compute_quant_params_ptr_float32 if din == D.F32 else ctx.compute_quant_params_ptr_bfloat16
        t = time_fn(lambda: fn(x.data_ptr(), D.UINT8, n), a.reps)
        report(name, 4 if din == D.F32 else 2, t, 0)
    for din, x, o in ((D.F32, xf, acc), (D.BF16, xb, accb)):
        name = f"requant {din.name} via UINT8 set"
        if a.cells != "all" and a.cells not in name:
            continue
        t = time_fn(lambda: ctx.requantize_ptr(x.data_ptr(), din, o.data_ptr(), D.UINT8, n, 2 / 255, 128), a.reps)
        report(name, 2 * (4 if din == D.F32 else 2), t, 0)
        for mode, op, label in ((RoundMode.STOCHASTIC, ReduceOp.SET, "stochastic set"), (RoundMode.NEAREST, ReduceOp.ADD, "nearest add")):
            t = time_fn(lambda: ctx.requantize_ptr(x.data_ptr(), din, o.data_ptr(), D.UINT8, n, 2 / 255, 128, mode, op), a.reps)
            report(f"requant {din.name} via UINT8 {label}", (3 if op == ReduceOp.ADD else 2) * (4 if din == D.F32 else 2), t, 0)
    # torch copy for calibration of the peak on this very box
    t = time_fn(lambda: acc.copy_(xf), a.reps)
    report("torch copy_ f32 (calibration)", 8, t, 0)
    t = time_fn(lambda: acc.copy_(q), a.reps)
    report("torch u8->f32 copy_ (1R:4W calib.)", 5, t, 0)
    t = time_fn(lambda: accb.copy_(q), a.reps)
    report("torch u8->bf16 copy_ (1R:2W calib.)", 3, t, 0)
    t = time_fn(lambda: q.copy_(xf), a.reps)
    report("torch f32->u8 copy_ (4R:1W calib.)", 5, t, 0)
    t = time_fn(lambda: acc.fill_(1.0), a.reps)
    report("torch fill_ f32 (write-only calib.)", 4, t, 0)
    t = time_fn(lambda: torch.cuda.memset if False else acc.zero_(), a.reps)
    report("torch zero_ f32 (memset calib.)", 4, t, 0)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"cellbench_{n}.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
