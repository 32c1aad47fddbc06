#!/usr/bin/env python
"""One-off probes of the GPU box (development tool)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)
import torch
print("cpu_count", os.cpu_count(), "torch", torch.__version__, torch.cuda.get_device_name(0), flush=True)
os.system("grep -m1 'model name' /proc/cpuinfo; free -g | head -2; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,pcie.link.gen.current,pcie.link.width.current --format=csv")
for dt in (torch.quint8, torch.quint4x2, torch.quint2x4):
    try:
        t = torch.empty((3, 5), dtype=dt, device="cuda")
        print(dt, "cuda empty ok", t.shape, t.untyped_storage().nbytes(), t.data_ptr() % 256)
    except Exception as e:
        print(dt, "cuda empty FAILED:", str(e)[:200])
import piquant, piquant.torch as pt
x = torch.rand(1000, device="cuda") * 2 - 1
for dt in (torch.quint8, torch.quint4x2, torch.quint2x4, torch.uint8):
    try:
        s, z = pt.compute_quant_params(x, dtype=dt)
        q = pt.quantize(x, scale=s, zero_point=z, dtype=dt)
        y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32)
        print(dt, "torch surface ok", s, z, float((y - x).abs().max()))
    except Exception as e:
        print(dt, "torch surface FAILED:", repr(e)[:300])
# host paths
n = 1 << 28
ctx = piquant.Context()
from piquant import DataType as D, RoundMode
xh = torch.empty(n, dtype=torch.float32).uniform_(-1, 1)
xp = xh.pin_memory()
qp = torch.empty(n, dtype=torch.uint8).pin_memory()
qh = torch.empty(n, dtype=torch.uint8)
for name, a, b in (("pinned->pinned", xp, qp), ("pageable->pageable", xh, qh)):
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.quantize_ptr(a.data_ptr(), D.F32, b.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
        dt_ = time.perf_counter() - t0
    print(f"host staged {name}: {dt_*1e3:.1f} ms  {n/dt_/1e9:.2f} Gelem/s  H2D {4*n/dt_/1e9:.1f} GB/s", flush=True)
os.environ["PIQUANT_CUDA_HOST_MODE"] = "zerocopy"
ctx2 = piquant.Context()
for rep in range(3):
    t0 = time.perf_counter()
    ctx2.quantize_ptr(xp.data_ptr(), D.F32, qp.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
    dt_ = time.perf_counter() - t0
print(f"host zerocopy pinned: {dt_*1e3:.1f} ms  {n/dt_/1e9:.2f} Gelem/s  H2D {4*n/dt_/1e9:.1f} GB/s", flush=True)
qd = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx.quantize_ptr(xh.cuda().data_ptr(), D.F32, qd.data_ptr(), D.UINT8, n, 2 / 255, 128, RoundMode.NEAREST)
torch.cuda.synchronize()
print("host results equal device result:", bool((qd.cpu() == qp).all()), bool((qd.cpu() == qh).all()))
t0 = time.perf_counter(); s = ctx.compute_quant_params_ptr_float32(xp.data_ptr(), D.UINT8, n); dt_ = time.perf_counter() - t0
print("host params", s, f"{dt_*1e3:.1f} ms", "device params", ctx.compute_quant_params_ptr_float32(xh.cuda().data_ptr(), D.UINT8, n))
