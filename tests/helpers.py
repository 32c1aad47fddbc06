"""Shared helpers for the test-suite (numpy side)."""
from __future__ import annotations

import itertools

import numpy as np

from oracle.port import (ADD, BF16, BITS, F32, NEAREST, SET, STOCHASTIC, UINT2, UINT4, UINT8,
                         bf16_bits_to_f32, f32_to_bf16_bits, packed_bytes)

FLOAT_DTYPES = (F32, BF16)
QUANT_DTYPES = (UINT2, UINT4, UINT8)
DT_NAME = {F32: "f32", BF16: "bf16", UINT2: "u2", UINT4: "u4", UINT8: "u8"}
MODE_NAME = {NEAREST: "nearest", STOCHASTIC: "stochastic"}
OP_NAME = {SET: "set", ADD: "add"}

QUANT_CELLS = list(itertools.product(FLOAT_DTYPES, QUANT_DTYPES))
DEQUANT_CELLS = list(itertools.product(QUANT_DTYPES, FLOAT_DTYPES, (SET, ADD)))


def cell_id(cell) -> str:
    names = []
    for i, c in enumerate(cell):
        names.append(DT_NAME[c] if i < 2 else (OP_NAME[c] if len(cell) == 3 else str(c)))
    return "-".join(names)


def aligned(nbytes: int, dtype=np.uint8, off: int = 0, align: int = 64) -> np.ndarray:
    """Zeroed array whose address is `off` bytes past a multiple of `align` -- the reference's
    f32->u8 kernel picks scalar-head / SIMD-body per element from the OUTPUT ADDRESS
    (kernels_specialized.inl:52-56), so byte-for-byte comparisons must fix the alignment."""
    raw = np.zeros(nbytes + align + off, dtype=np.uint8)
    start = (-raw.ctypes.data) % align + off
    return raw[start:start + nbytes].view(dtype)


def make_input(rng: np.random.Generator, n: int, dt: int, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """U(lo,hi) like the reference's tests (test/quant.cpp:66, bench.cpp:24); bf16 = RNE of the f32 sample."""
    x = rng.uniform(lo, hi, n).astype(np.float32)
    return x if dt == F32 else f32_to_bf16_bits(x)


def as_f32(x: np.ndarray) -> np.ndarray:
    return x if x.dtype == np.float32 else bf16_bits_to_f32(x)


def special_values(scale: float) -> np.ndarray:
    """Inputs that land on the reference's corner cases once multiplied by 1/scale: ties, NaN, infinities,
    out-of-int32-range products, pred(0.5) (where body and tail formulas disagree), denormals."""
    base = [0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, np.nan, np.inf, -np.inf, 3e9, -3e9, 1e20, -1e20,
            2147483520.0, -2147483648.0, 2147483648.0, 0.49999997, -0.49999997, 8388609.0, -8388609.0,
            16777215.0, 1e-40, -1e-40, 254.5, 255.5, 14.5, 15.5, 3.5, 2.5, 127.49999, 0.99999994]
    with np.errstate(all="ignore"):
        return (np.array(base, dtype=np.float32) * np.float32(scale)).astype(np.float32)


def unpack(q: np.ndarray, dt: int, numel: int) -> np.ndarray:
    """Packed bytes -> one uint8 per element (low element in low bits, piquant.hpp:50-76 / quantize.inl:36-50)."""
    bits = BITS[dt]
    per = 8 // bits
    if per == 1:
        return q[:numel].copy()
    shifts = (np.arange(per, dtype=np.uint8) * bits)[None, :]
    return ((q[:, None] >> shifts) & ((1 << bits) - 1)).reshape(-1)[:numel].astype(np.uint8)


def infer_xi(x_f32: np.ndarray, scale: float, zero_point: int, qmax: int, q_elems: np.ndarray) -> float | None:
    """A per-call stochastic threshold consistent with the reference's output, or None (oracle.port.infer_stochastic_threshold)."""
    from oracle.port import infer_stochastic_threshold
    return infer_stochastic_threshold(x_f32, scale, zero_point, qmax, q_elems)
