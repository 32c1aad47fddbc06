#!/usr/bin/env python
"""Launches a few chosen cells of the kernel matrix a fixed number of times so that `ncu` can capture them
(development tool).  Usage, on the GPU box:

    ncu --set full --clock-control none --import-source on -k regex:'quant_stream|requant' \
        -o gpurun_out/cells python tools/ncu_cells.py --numel 1000000000 --launches 2
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
from piquant import DataType as D, RoundMode  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=1_000_000_000)
    ap.add_argument("--launches", type=int, default=2)
    ap.add_argument("--cells", default="bf16u2n,bf16u2s,bf16u4s,bf16u4n,requant_bf16")
    ap.add_argument("--variant", type=int, default=1, help="1 = direct LDG.256 kernels, 2 = TMA ring kernels")
    a = ap.parse_args()
    n = a.numel
    torch.cuda.set_device(0)
    ctx = piquant.Context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_stochastic_threshold(0.37)
    ctx.set_kernel_variant(a.variant)
    g = torch.Generator(device="cuda").manual_seed(0)
    xf = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-1, 1, generator=g)
    xb = xf.to(torch.bfloat16)
    q = torch.empty(n, dtype=torch.uint8, device="cuda")
    ob = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    table = {
        "f32u8n": lambda: ctx.quantize_ptr(xf.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST),
        "bf16u8n": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST),
        "bf16u8s": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.STOCHASTIC),
        "bf16u4n": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT4, n, 2 / 15, 8, RoundMode.NEAREST),
        "bf16u4s": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT4, n, 2 / 15, 8, RoundMode.STOCHASTIC),
        "bf16u2n": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT2, n, 2 / 3, 2, RoundMode.NEAREST),
        "bf16u2s": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT2, n, 2 / 3, 2, RoundMode.STOCHASTIC),
        "f32u8pe": lambda: ctx.quantize_ptr(xf.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.STOCHASTIC_PER_ELEMENT),
        "bf16u8pe": lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.STOCHASTIC_PER_ELEMENT),
        "f32i8n": lambda: ctx.quantize_ptr(xf.data_ptr(), D.F32, q.data_ptr(), D.INT8, n, 2 / 255, 0, RoundMode.NEAREST),
        "requant_bf16": lambda: ctx.requantize_ptr(xb.data_ptr(), D.BF16, ob.data_ptr(), D.UINT8, n, 2 / 255, 128),
    }
    # round-2 kernels: parameters on the device, fused accumulate, multi-source reduce (chunk of an 8-GPU all-reduce by default)
    dev_i, st = torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream
    meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
    nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
    srcs = []

    def sources(k):
        while len(srcs) < k:
            srcs.append(torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda"))
        return srcs[:k]

    ctx.compute_meta_on_stream(xf.data_ptr(), D.F32, min(n, 1 << 20), D.UINT8, meta.data_ptr(), piquant.Context.FLAG_LOCAL, dev_i, st)
    acc = torch.zeros(n, dtype=torch.float32, device="cuda")
    table.update({
        "meta_f32": lambda: ctx.compute_meta_on_stream(xf.data_ptr(), D.F32, n, D.UINT8, nxt.data_ptr(), piquant.Context.FLAG_LOCAL, dev_i, st),
        "addmm_u8_f32": lambda: ctx.dequantize_add_minmax_on_stream(q.data_ptr(), D.UINT8, acc.data_ptr(), D.F32, n, meta.data_ptr(), D.UINT8,
                                                                    nxt.data_ptr(), 0, dev_i, st),
        "sum7_u8_f32": lambda: ctx.dequantize_sum_minmax_on_stream([t.data_ptr() for t in sources(7)], D.UINT8, acc.data_ptr(), D.F32, n,
                                                                   [meta.data_ptr()] * 7, D.UINT8, nxt.data_ptr(), 0, dev_i, st),
        "sum1_u8_f32": lambda: ctx.dequantize_sum_minmax_on_stream([t.data_ptr() for t in sources(1)], D.UINT8, acc.data_ptr(), D.F32, n,
                                                                   [meta.data_ptr()], D.UINT8, nxt.data_ptr(), 0, dev_i, st),
        "sum7_u4_bf16": lambda: ctx.dequantize_sum_minmax_on_stream([t.data_ptr() for t in sources(7)], D.UINT4, ob.data_ptr(), D.BF16, n,
                                                                    [meta.data_ptr()] * 7, D.UINT8, nxt.data_ptr(), 0, dev_i, st),
    })
    for name in a.cells.split(","):
        for _ in range(a.launches):
            table[name]()
        torch.cuda.synchronize()
        print("launched", name, flush=True)


if __name__ == "__main__":
    main()
