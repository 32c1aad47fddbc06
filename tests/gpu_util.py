"""Drives libpiquant.so through its C ABI (the `piquant` package's pointer-level methods) with numpy
inputs staged into CUDA memory by torch.  Used by the `-m gpu` parity tests, smoke() and bench.py."""
from __future__ import annotations

import numpy as np
import torch

import piquant
from piquant import DataType, ReduceOp, RoundMode

from oracle.port import BF16, F32, INT2, INT4, INT8, UINT2, UINT4, UINT8, packed_bytes

DT = {F32: DataType.F32, BF16: DataType.BF16, UINT2: DataType.UINT2, UINT4: DataType.UINT4, UINT8: DataType.UINT8,
      INT2: DataType.INT2, INT4: DataType.INT4, INT8: DataType.INT8}
MODE = {0: RoundMode.NEAREST, 1: RoundMode.STOCHASTIC, 2: RoundMode.STOCHASTIC_PER_ELEMENT}
OP = {0: ReduceOp.SET, 1: ReduceOp.ADD}
CANARY = 0xCD
_BASE: dict = {}


def to_dev(a: np.ndarray, offset_bytes: int = 0, track: bool = False) -> torch.Tensor:
    """Copy a numpy array into CUDA memory as raw bytes, `offset_bytes` past a 256-byte aligned address."""
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    buf = torch.full((raw.size + offset_bytes + 64,), CANARY, dtype=torch.uint8, device="cuda")
    t = buf[offset_bytes:offset_bytes + raw.size]
    if raw.size:
        t.copy_(torch.from_numpy(raw))
    if track:
        _BASE[t.data_ptr() if raw.size else id(t)] = (buf, offset_bytes, raw.size)
    return t


def check_canary(t: torch.Tensor) -> None:
    """The bytes around an output buffer made by to_dev() must be untouched."""
    key = t.data_ptr() if t.numel() else id(t)
    buf, off, size = _BASE.pop(key)
    host = buf.cpu().numpy()
    assert (host[:off] == CANARY).all(), "kernel wrote in front of the output buffer"
    assert (host[off + size:] == CANARY).all(), "kernel wrote past the end of the output buffer"


def to_host(t: torch.Tensor, dtype) -> np.ndarray:
    return t.cpu().numpy().view(dtype).copy()


def np_dtype_of(dt: int):
    return {F32: np.float32, BF16: np.uint16}.get(dt, np.uint8)


class Gpu:
    """numpy in / numpy out wrapper over one piquant.Context."""

    def __init__(self, variant: int = 0) -> None:
        self.ctx = piquant.Context()
        self.ctx.set_kernel_variant(variant)

    def quantize(self, x: np.ndarray, dt_out: int, scale: float, zp: int, mode: int = 0, xi: float | None = None,
                 in_off: int = 0, out_off: int = 0) -> np.ndarray:
        dt_in = F32 if x.dtype == np.float32 else BF16
        n = x.size
        d_in = to_dev(x, in_off)
        d_out = to_dev(np.full(packed_bytes(dt_out, n), 0xAA, np.uint8), out_off, track=True)
        if mode == 1:
            self.ctx.set_stochastic_threshold(xi)
        self.ctx.quantize_ptr(d_in.data_ptr() if n else 1, DT[dt_in], d_out.data_ptr() if n else 1, DT[dt_out], n, scale, zp, MODE[mode])
        torch.cuda.synchronize()
        check_canary(d_out)
        return to_host(d_out, np.uint8)

    def quantize_sr(self, x: np.ndarray, dt_out: int, scale: float, zp: int, key: int, in_off: int = 0, out_off: int = 0) -> np.ndarray:
        """per-element stochastic rounding (extension), Philox key `key`"""
        self.ctx.set_sr_key(key)
        try:
            return self.quantize(x, dt_out, scale, zp, mode=2, in_off=in_off, out_off=out_off)
        finally:
            self.ctx.set_sr_key(None)

    def dequantize(self, q: np.ndarray, dt_in: int, numel: int, dt_out: int, scale: float, zp: int, op: int = 0,
                   prev: np.ndarray | None = None, in_off: int = 0, out_off: int = 0) -> np.ndarray:
        odt = np_dtype_of(dt_out)
        if prev is None:
            prev = np.zeros(numel, dtype=odt)
        d_in = to_dev(q, in_off)
        d_out = to_dev(prev, out_off, track=True)
        self.ctx.dequantize_ptr(d_in.data_ptr() if numel else 1, DT[dt_in], d_out.data_ptr() if numel else 1, DT[dt_out], numel, scale, zp, OP[op])
        torch.cuda.synchronize()
        check_canary(d_out)
        return to_host(d_out, odt)

    def requantize(self, x: np.ndarray, dt_q: int, scale: float, zp: int, mode: int = 0, xi: float | None = None, op: int = 0,
                   prev: np.ndarray | None = None, in_off: int = 0, out_off: int = 0) -> np.ndarray:
        dt_io = F32 if x.dtype == np.float32 else BF16
        if prev is None:
            prev = np.zeros(x.size, dtype=x.dtype)
        d_in = to_dev(x, in_off)
        d_out = to_dev(prev, out_off, track=True)
        if mode == 1:
            self.ctx.set_stochastic_threshold(xi)
        self.ctx.requantize_ptr(d_in.data_ptr(), DT[dt_io], d_out.data_ptr(), DT[dt_q], x.size, scale, zp, MODE[mode], OP[op])
        torch.cuda.synchronize()
        check_canary(d_out)
        return to_host(d_out, x.dtype)

    def requantize_sr(self, x: np.ndarray, dt_q: int, scale: float, zp: int, key: int, op: int = 0, prev: np.ndarray | None = None,
                      in_off: int = 0, out_off: int = 0) -> np.ndarray:
        self.ctx.set_sr_key(key)
        try:
            return self.requantize(x, dt_q, scale, zp, mode=2, op=op, prev=prev, in_off=in_off, out_off=out_off)
        finally:
            self.ctx.set_sr_key(None)

    def compute_quant_params(self, x: np.ndarray, dt_q: int, in_off: int = 0) -> tuple[float, int]:
        d_in = to_dev(x, in_off)
        if x.dtype == np.float32:
            return self.ctx.compute_quant_params_ptr_float32(d_in.data_ptr(), DT[dt_q], x.size)
        return self.ctx.compute_quant_params_ptr_bfloat16(d_in.data_ptr(), DT[dt_q], x.size)
