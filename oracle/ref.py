"""ctypes binding of oracle/_ref/libpiquant_ref.so -- the UNMODIFIED reference, compiled from
/root/reference by oracle/Makefile.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Binds exactly the six functions of the reference's include/piquant.h:42-85.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import REF_LIB, build, have_ref
from .port import BITS, NP_DTYPE, dtype_of, packed_bytes

_lib = None


def available() -> bool:
    if not have_ref():
        try:
            build()
        except Exception:
            return False
    return have_ref()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libpiquant_ref.so is missing (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(str(REF_LIB))
        vp = C.c_void_p
        L.piquant_context_create.argtypes = [C.c_size_t]; L.piquant_context_create.restype = vp
        L.piquant_context_destroy.argtypes = [vp]; L.piquant_context_destroy.restype = None
        L.piquant_quantize.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_size_t, C.c_float, C.c_int64, C.c_int]
        L.piquant_quantize.restype = None
        L.piquant_dequantize.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_size_t, C.c_float, C.c_int64, C.c_int]
        L.piquant_dequantize.restype = None
        for name in ("piquant_compute_quant_params_float32", "piquant_compute_quant_params_bfloat16"):
            fn = getattr(L, name)
            fn.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
            fn.restype = None
        if hasattr(L, "piquant_ref_shim_requantize"):   # oracle/ref_shim.cpp
            L.piquant_ref_shim_requantize.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_size_t, C.c_float, C.c_int64, C.c_int, C.c_int]
            L.piquant_ref_shim_requantize.restype = None
        _lib = L
    return _lib


def cpu_isa() -> str:
    """Which kernel set the reference's CPUID dispatch (piquant.cpp:178-188) picks on this host."""
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except Exception:
        return "unknown"
    if "avx512_bf16" in flags and "avx512f" in flags and "avx512bw" in flags:
        return "avx512f_bf16"
    if "avx512f" in flags:
        return "avx512f"
    if "avx2" in flags:
        return "avx2"
    if "sse4_2" in flags:
        return "sse42"
    return "generic"


class Context:
    """piquant::context through the reference's C ABI."""

    def __init__(self, num_threads: int | None = None) -> None:
        if num_threads is None:
            num_threads = max(1, os.cpu_count() or 1)
        self.num_threads = num_threads
        self._ctx = lib().piquant_context_create(num_threads)

    def close(self) -> None:
        if self._ctx:
            lib().piquant_context_destroy(self._ctx)
            self._ctx = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def quantize(self, x: np.ndarray, dt_out: int, scale: float, zero_point: int, mode: int = 0,
                 out: np.ndarray | None = None) -> np.ndarray:
        n = x.size
        if out is None:
            out = np.zeros(packed_bytes(dt_out, n), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size == packed_bytes(dt_out, n)
        lib().piquant_quantize(self._ctx, x.ctypes.data, dtype_of(x), out.ctypes.data, dt_out, n, scale, zero_point, mode)
        return out

    def dequantize(self, q: np.ndarray, dt_in: int, numel: int, dt_out: int, scale: float, zero_point: int,
                   op: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        assert q.dtype == np.uint8 and q.size == packed_bytes(dt_in, numel)
        if out is None:
            out = np.zeros(numel, dtype=NP_DTYPE[dt_out])
        lib().piquant_dequantize(self._ctx, q.ctypes.data, dt_in, out.ctypes.data, dt_out, numel, scale, zero_point, op)
        return out

    def requantize(self, x: np.ndarray, dt_quant: int, scale: float, zero_point: int, mode: int = 0, op: int = 0,
                   out: np.ndarray | None = None) -> np.ndarray:
        """context::quantize_dequantize_fused (piquant.hpp:276-285) via oracle/ref_shim.cpp."""
        if out is None:
            out = np.zeros(x.size, dtype=x.dtype)
        lib().piquant_ref_shim_requantize(self._ctx, x.ctypes.data, dtype_of(x), out.ctypes.data, dt_quant, x.size,
                                          scale, zero_point, mode, op)
        return out

    def compute_quant_params(self, x: np.ndarray, dt_quant: int) -> tuple[float, int]:
        s, z = C.c_float(), C.c_int64()
        name = "piquant_compute_quant_params_float32" if x.dtype == np.float32 else "piquant_compute_quant_params_bfloat16"
        getattr(lib(), name)(self._ctx, x.ctypes.data, x.size, dt_quant, C.byref(s), C.byref(z))
        return s.value, z.value


__all__ = ["Context", "available", "cpu_isa", "lib", "BITS"]
