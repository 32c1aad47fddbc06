#!/usr/bin/env python
"""Device time of the round-2 fused kernels against their algorithmic bytes (CUDA events, rotating accumulators so that L2
cannot serve them): min/max + parameters, dequantize-ADD + min/max, the multi-source reduce with 1..8 sources.  Development tool."""
from __future__ import annotations

import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
from piquant import DataType as D  # noqa: E402


def main() -> None:
    torch.cuda.set_device(0)
    ctx = piquant.Context()
    dev, st = 0, torch.cuda.current_stream().cuda_stream
    try:
        peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:      # noqa: BLE001
        peak = 6550.7
    for n in (1 << 25, 125_000_000):
        accs = [torch.zeros(n, dtype=torch.float32, device="cuda") for _ in range(4)]
        accb = [torch.zeros(n, dtype=torch.bfloat16, device="cuda") for _ in range(4)]
        srcs = [torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda") for _ in range(8)]
        x = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-1, 1)
        meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
        nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
        ctx.compute_meta_on_stream(x.data_ptr(), D.F32, 1 << 20, D.UINT8, meta.data_ptr(), piquant.Context.FLAG_LOCAL, dev, st)

        def ev(fn, reps=20):
            for k in range(3):
                fn(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps):
                fn(k)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3

        rows = [("min/max + parameters (f32)", 4, ev(lambda k: ctx.compute_meta_on_stream(x.data_ptr(), D.F32, n, D.UINT8, nxt.data_ptr(), piquant.Context.FLAG_LOCAL, dev, st)))]
        rows.append(("dequantize-ADD + min/max (u8 -> f32)", 9, ev(lambda k: ctx.dequantize_add_minmax_on_stream(
            srcs[0].data_ptr(), D.UINT8, accs[k % 4].data_ptr(), D.F32, n, meta.data_ptr(), D.UINT8, nxt.data_ptr(), 0, dev, st))))
        for k_src in (1, 2, 4, 7, 8):
            rows.append((f"multi-source reduce, {k_src} x u8 -> f32", 8 + k_src, ev(lambda k: ctx.dequantize_sum_minmax_on_stream(
                [t.data_ptr() for t in srcs[:k_src]], D.UINT8, accs[k % 4].data_ptr(), D.F32, n, [meta.data_ptr()] * k_src, D.UINT8, nxt.data_ptr(), 0, dev, st))))
        for k_src in (1, 7):
            rows.append((f"multi-source reduce, {k_src} x u4 -> bf16", 4 + 0.5 * k_src, ev(lambda k: ctx.dequantize_sum_minmax_on_stream(
                [t.data_ptr() for t in srcs[:k_src]], D.UINT4, accb[k % 4].data_ptr(), D.BF16, n, [meta.data_ptr()] * k_src, D.UINT8, nxt.data_ptr(), 0, dev, st))))
        rows.append(("7 separate dequantize-ADD launches (u8 -> f32)", 63, ev(lambda k: [ctx.dequantize_meta_on_stream(
            srcs[i].data_ptr(), D.UINT8, accs[k % 4].data_ptr(), D.F32, n, piquant.ReduceOp.ADD, meta.data_ptr(), dev, st) for i in range(7)], reps=6)))
        print(f"numel = {n}")
        for name, bpe, us in rows:
            gbs = bpe * n / us / 1e3
            print(f"  {name:50s} {us:9.2f} us  {gbs:8.1f} GB/s  {100 * gbs / peak:6.1f} % of the measured copy peak ({bpe} B/element)")


if __name__ == "__main__":
    main()
