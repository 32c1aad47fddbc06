#!/usr/bin/env python
"""Small, fast exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np

from gpu_util import Gpu
from oracle import port
from oracle.port import ADD, BF16, F32, INT4, INT8, SET, UINT2, UINT4, UINT8, f32_to_bf16_bits, packed_bytes

rng = np.random.default_rng(0)
for variant in (1, 2):
    g = Gpu(variant=variant)
    for n in (70_001, 5):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        xb = f32_to_bf16_bits(x)
        for xin in (x, xb):
            for dq in (UINT8, UINT4, UINT2):
                s, z = g.compute_quant_params(xin, dq)
                assert (s, z) == port.compute_quant_params(xin, dq)
                if variant == 1:      # per-element stochastic rounding (extension), direct kernels: vector and byte-granular paths
                    for out_off in (0, 3):
                        assert np.array_equal(g.quantize_sr(xin, dq, s, z, 77, out_off=out_off), port.quantize_sr(xin, dq, s, z, 77))
                for mode in (0, 1):
                    q = g.quantize(xin, dq, s, z, mode, xi=0.3)
                    assert np.array_equal(q, port.quantize(xin, dq, s, z, mode, xi=0.3))
                for op in (SET, ADD):
                    prev = rng.uniform(-1, 1, n).astype(np.float32)
                    prev = prev if xin.dtype == np.float32 else f32_to_bf16_bits(prev)
                    dt_out = F32 if xin.dtype == np.float32 else BF16
                    y = g.dequantize(q, dq, n, dt_out, s, z, op, prev=prev)
                    assert np.array_equal(y, port.dequantize(q, dq, n, dt_out, s, z, op, out=prev.copy()))
                    r = g.requantize(xin, dq, s, z, 0, None, op, prev=prev)
                    assert np.array_equal(r, port.requantize(xin, dq, s, z, 0, 0.0, op, out=prev.copy()))
    # signed extension dtypes
    for sdt in (INT8, INT4):
        x = rng.uniform(-1, 1, 70_001).astype(np.float32)
        s, z = g.compute_quant_params(x, sdt)
        q = g.quantize(x, sdt, s, z)
        assert np.array_equal(q, port.quantize(x, sdt, s, z))
        assert np.array_equal(g.dequantize(q, sdt, x.size, F32, s, z), port.dequantize(q, sdt, x.size, F32, s, z))
print("sanitize probe ok")
