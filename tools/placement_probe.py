#!/usr/bin/env python
"""Does throughput depend on WHICH buffers are used (physical placement) or on rotating between them?
Times f32->u8 quantize at numel=1e9 on 3 separately allocated buffer pairs, each pair alone (fixed) and
all three in rotation, for both kernel variants.  Development tool."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
from piquant import DataType as D, RoundMode  # noqa: E402

n = 1_000_000_000
torch.cuda.set_device(0)
ctx = piquant.Context()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
xs = [torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-1, 1) for _ in range(3)]
qs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(3)]
print("x ptrs", [hex(x.data_ptr()) for x in xs], "q ptrs", [hex(q.data_ptr()) for q in qs])


def run(pairs, reps=12):
    for i in range(4):
        x, q = pairs[i % len(pairs)]
        ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        x, q = pairs[i % len(pairs)]
        ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for variant in (1, 2):
    ctx.set_kernel_variant(variant)
    for rnd in range(2):
        res = []
        for i in range(3):
            for j in range(3):
                res.append((f"x{i}q{j}", run([(xs[i], qs[j])])))
        res.append(("rot", run([(xs[i], qs[i]) for i in range(3)])))
        print(f"variant {variant} round {rnd}: " + "  ".join(f"{k}={5 * n / (v * 1e-3) / 1e9:6.0f}" for k, v in res) + "  GB/s", flush=True)
