#!/usr/bin/env python
"""Device time of piquant_cuda_quantize_batch on 1000 distinct 1e6-element f32 tensors (preallocated outputs), next to ONE launch
over the same 1e9 elements.  Development tool."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
import piquant.torch as pt  # noqa: E402


def main() -> None:
    torch.cuda.set_device(0)
    ctx = piquant.Context()
    runs, numel = 1000, 1_000_000
    many = torch.rand(runs, numel, dtype=torch.float32, device="cuda")
    ins = [many[i] for i in range(runs)]
    for tdt, bpe in ((torch.quint8, 5.0), (torch.quint4x2, 4.5), (torch.quint2x4, 4.25)):
        outs = [torch.empty(numel, dtype=tdt, device="cuda") for _ in range(runs)]
        args = dict(scales=[2 / 255] * runs, zero_points=[128] * runs, dtype=tdt, ctx=ctx, outs=outs)
        whole = torch.empty(runs * numel, dtype=tdt, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res = []
        prepared = pt.QuantizeBatch(ins, **args)
        for fn in (lambda: pt.quantize_batch(ins, **args), prepared.run, lambda: pt.quantize(many.view(-1), scale=2 / 255, zero_point=128, dtype=tdt, ctx=ctx, out=whole)):
            fn()
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) * 1e-3)
        print(f"{str(tdt):18s} batch of {runs} x {numel}: {res[0] / runs * 1e6:6.3f} us per tensor = {bpe * runs * numel / res[0] / 1e9:7.1f} GB/s;   "
              f"prepared batch: {res[1] / runs * 1e6:6.3f} us per tensor = {bpe * runs * numel / res[1] / 1e9:7.1f} GB/s;   "
              f"one launch over {runs * numel}: {bpe * runs * numel / res[2] / 1e9:7.1f} GB/s")
        del outs, whole


if __name__ == "__main__":
    main()
