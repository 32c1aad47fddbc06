"""ctypes binding of liboracle.so (piquant_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Arrays are numpy: f32 -> float32, bf16 -> uint16 (raw bits), quantized -> uint8 (packed bytes).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import PORT_LIB, build

F32, BF16, UINT2, UINT4, UINT8 = 0, 1, 2, 3, 4
INT2, INT4, INT8 = 5, 6, 7          # signed extension (piquant_oracle.h): offset-binary view of the unsigned types
NEAREST, STOCHASTIC = 0, 1
SET, ADD = 0, 1
SEM_BODY, SEM_REF = 0, 1
BITS = {F32: 32, BF16: 16, UINT2: 2, UINT4: 4, UINT8: 8, INT2: 2, INT4: 4, INT8: 8}
NP_DTYPE = {F32: np.float32, BF16: np.uint16, UINT2: np.uint8, UINT4: np.uint8, UINT8: np.uint8, INT2: np.uint8, INT4: np.uint8, INT8: np.uint8}
SIGNED = {INT2: UINT2, INT4: UINT4, INT8: UINT8}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build(ref=False)
        L = C.CDLL(str(PORT_LIB))
        vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
        L.orc_packed_bytes.argtypes = [i32, C.c_size_t]; L.orc_packed_bytes.restype = C.c_size_t
        L.orc_storage_bytes.argtypes = [i32, C.c_size_t]; L.orc_storage_bytes.restype = C.c_size_t
        L.orc_quantize.argtypes = [vp, i32, vp, i32, i64, f32, i64, i32, f32, i32, i32]; L.orc_quantize.restype = i32
        L.orc_quantize_sr.argtypes = [vp, i32, vp, i32, i64, f32, i64, C.c_uint64, i64]; L.orc_quantize_sr.restype = i32
        L.orc_requantize_sr.argtypes = [vp, i32, vp, i32, i64, f32, i64, C.c_uint64, i64, i32, i32]; L.orc_requantize_sr.restype = i32
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]; L.orc_philox4x32_10.restype = None
        L.orc_dequantize.argtypes = [vp, i32, vp, i32, i64, f32, i64, i32, i32, i32]; L.orc_dequantize.restype = i32
        L.orc_requantize.argtypes = [vp, i32, vp, i32, i64, f32, i64, i32, f32, i32, i32]; L.orc_requantize.restype = i32
        L.orc_minmax_f32.argtypes = [vp, i64, vp]; L.orc_minmax_f32.restype = None
        L.orc_minmax_bf16.argtypes = [vp, i64, vp]; L.orc_minmax_bf16.restype = None
        L.orc_params_from_minmax.argtypes = [f64, f64, i32, C.POINTER(f32), C.POINTER(i64)]; L.orc_params_from_minmax.restype = i32
        L.orc_compute_quant_params_f32.argtypes = [vp, i64, i32, C.POINTER(f32), C.POINTER(i64)]; L.orc_compute_quant_params_f32.restype = i32
        L.orc_compute_quant_params_bf16.argtypes = [vp, i64, i32, C.POINTER(f32), C.POINTER(i64)]; L.orc_compute_quant_params_bf16.restype = i32
        L.orc_set_fma_contract.argtypes = [i32]; L.orc_set_fma_contract.restype = None
        L.orc_f32_to_bf16.argtypes = [f32]; L.orc_f32_to_bf16.restype = C.c_uint16
        L.orc_bf16_to_f32.argtypes = [C.c_uint16]; L.orc_bf16_to_f32.restype = f32
        L.orc_quant_step_body.argtypes = [f32, f32, C.c_int32, C.c_int32]; L.orc_quant_step_body.restype = C.c_int32
        L.orc_quant_step_tail32.argtypes = [f32, f32, C.c_int32, C.c_int32]; L.orc_quant_step_tail32.restype = C.c_int32
        L.orc_quant_step_scalar_nearest.argtypes = [f32, f32, i64, i64]; L.orc_quant_step_scalar_nearest.restype = i64
        L.orc_quant_step_scalar_stochastic.argtypes = [f32, f32, i64, i64, f32]; L.orc_quant_step_scalar_stochastic.restype = i64
        _lib = L
    return _lib


def packed_bytes(dt: int, numel: int) -> int:
    per = 8 // BITS[dt]
    return (numel + per - 1) // per


def _ptr(a: np.ndarray) -> int:
    assert a.flags.c_contiguous
    return a.ctypes.data


def dtype_of(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.uint16:
        return BF16
    raise TypeError(f"float tensors must be float32 or uint16(bf16 bits), got {a.dtype}")


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 bits (piquant.hpp:86-90); NaN forced quiet."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + (0x7FFF + ((u >> 16) & 1))) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 64).astype(np.uint16)
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def quantize(x: np.ndarray, dt_out: int, scale: float, zero_point: int, mode: int = NEAREST,
             xi: float = 0.0, semantics: int = SEM_BODY, nthreads: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    n = x.size
    if out is None:
        out = np.zeros(packed_bytes(dt_out, n), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size == packed_bytes(dt_out, n)
    rc = lib().orc_quantize(_ptr(x), dtype_of(x), _ptr(out), dt_out, n, scale, zero_point, mode, xi, semantics, nthreads)
    if rc != 0:
        raise ValueError("invalid dtype combination")
    return out


def quantize_sr(x: np.ndarray, dt_out: int, scale: float, zero_point: int, key: int, base: int = 0) -> np.ndarray:
    """Extension: per-element stochastic rounding (piquant_oracle.h), Philox key `key`, element 0 has index `base`."""
    out = np.zeros(packed_bytes(dt_out, x.size), dtype=np.uint8)
    rc = lib().orc_quantize_sr(_ptr(x), dtype_of(x), _ptr(out), dt_out, x.size, scale, zero_point, key & (2**64 - 1), base)
    if rc != 0:
        raise ValueError("invalid dtype combination")
    return out


def requantize_sr(x: np.ndarray, dt_quant: int, scale: float, zero_point: int, key: int, base: int = 0, op: int = SET,
                  out: np.ndarray | None = None, fma_add: bool = True) -> np.ndarray:
    if out is None:
        out = np.zeros(x.size, dtype=x.dtype)
    rc = lib().orc_requantize_sr(_ptr(x), dtype_of(x), _ptr(out), dt_quant, x.size, scale, zero_point, key & (2**64 - 1), base, op, int(fma_add))
    if rc != 0:
        raise ValueError("invalid dtype combination")
    return out


def philox4x32_10(ctr, key) -> list[int]:
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return [int(v) for v in o]


def dequantize(q: np.ndarray, dt_in: int, numel: int, dt_out: int, scale: float, zero_point: int, op: int = SET,
               out: np.ndarray | None = None, semantics: int = SEM_BODY, nthreads: int = 0) -> np.ndarray:
    assert q.dtype == np.uint8 and q.size == packed_bytes(dt_in, numel)
    if out is None:
        out = np.zeros(numel, dtype=NP_DTYPE[dt_out])
    assert out.size == numel and dtype_of(out) == dt_out
    rc = lib().orc_dequantize(_ptr(q), dt_in, _ptr(out), dt_out, numel, scale, zero_point, op, semantics, nthreads)
    if rc != 0:
        raise ValueError("invalid dtype combination")
    return out


def requantize(x: np.ndarray, dt_quant: int, scale: float, zero_point: int, mode: int = NEAREST, xi: float = 0.0,
               op: int = SET, out: np.ndarray | None = None, fma_add: bool = True) -> np.ndarray:
    if out is None:
        out = np.zeros(x.size, dtype=x.dtype)
    rc = lib().orc_requantize(_ptr(x), dtype_of(x), _ptr(out), dt_quant, x.size, scale, zero_point, mode, xi, op, int(fma_add))
    if rc != 0:
        raise ValueError("invalid dtype combination")
    return out


def minmax(x: np.ndarray) -> tuple[float, float]:
    out = np.zeros(2, dtype=np.float32)
    if dtype_of(x) == F32:
        lib().orc_minmax_f32(_ptr(x), x.size, _ptr(out))
    else:
        lib().orc_minmax_bf16(_ptr(x), x.size, _ptr(out))
    return float(out[0]), float(out[1])


def params_from_minmax(mn: float, mx: float, dt_quant: int) -> tuple[float, int]:
    s, z = C.c_float(), C.c_int64()
    if lib().orc_params_from_minmax(mn, mx, dt_quant, C.byref(s), C.byref(z)) != 0:
        raise ValueError("scale must be positive (reference aborts)")
    return s.value, z.value


def compute_quant_params(x: np.ndarray, dt_quant: int) -> tuple[float, int]:
    s, z = C.c_float(), C.c_int64()
    fn = lib().orc_compute_quant_params_f32 if dtype_of(x) == F32 else lib().orc_compute_quant_params_bf16
    if fn(_ptr(x), x.size, dt_quant, C.byref(s), C.byref(z)) != 0:
        raise ValueError("scale must be positive (reference aborts)")
    return s.value, z.value


def infer_stochastic_threshold(x_f32: np.ndarray, scale: float, zero_point: int, qmax: int, q_elems: np.ndarray) -> float | None:
    """The reference draws ONE stochastic threshold xi per call from a random_device-seeded RNG
    (piquant.cpp:194-201), so its output cannot be reproduced from a seed.  It can still be pinned:
    every element with dec > xi rounds away from zero and every element with dec <= xi truncates
    (quantize.inl:8-19), so the output brackets xi.  Returns a xi consistent with q_elems (one unpacked value per
    element), or None when no single threshold explains the output."""
    inv = np.float32(1.0) / np.float32(scale)
    with np.errstate(all="ignore"):
        r = x_f32.astype(np.float32) * inv
        tr = np.trunc(r)
        dec = np.abs(r - tr)
        trunc_q = np.clip(tr.astype(np.int64) + zero_point, 0, qmax)
        away_q = np.clip((tr + np.where(r < 0, -1.0, 1.0).astype(np.float32)).astype(np.int64) + zero_point, 0, qmax)
    informative = trunc_q != away_q
    went_away = informative & (q_elems.astype(np.int64) == away_q)
    stayed = informative & (q_elems.astype(np.int64) == trunc_q)
    lo = float(dec[stayed].max()) if stayed.any() else 0.0       # xi >= dec of every truncated element
    hi = float(dec[went_away].min()) if went_away.any() else 1.0  # xi <  dec of every rounded-away element
    if not lo < hi:
        return None
    return float(np.float32(lo)) if lo > 0.0 else 0.0


def set_fma_contract(on: bool) -> None:
    lib().orc_set_fma_contract(int(on))
