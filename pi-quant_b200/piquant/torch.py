"""piquant.torch -- tensor-level surface, device-aware.

Same three functions, keyword names, accepted dtypes and return conventions as the reference's
``piquant.torch`` (reference python/src/piquant/torch.py:9-129).  What changes for the B200 build:

* outputs are allocated on ``tensor.device`` (the reference always allocates on the CPU,
  reference torch.py:87,117);
* for CUDA tensors the work is enqueued on PyTorch's *current* stream of that device, so calls
  compose with surrounding torch ops without extra synchronisation; ``compute_quant_params``
  returns Python scalars and therefore synchronises;
* ``dequantize`` takes an optional ``out=`` so that ``reduce_op='add'`` has a defined accumulator
  (the reference accumulates into an uninitialised ``torch.empty``, reference torch.py:117);
* ``requantize`` exposes the fused quantize->dequantize pass (C++-only in the reference).

CPU tensors are accepted as well: the native library streams them through the GPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import Context, DataType, ReduceOp, RoundMode

_TORCH_DTYPE_MAP: dict = {
    torch.float32: DataType.F32,
    torch.bfloat16: DataType.BF16,
    torch.quint2x4: DataType.UINT2,
    torch.quint4x2: DataType.UINT4,
    torch.quint8: DataType.UINT8,
    torch.uint8: DataType.UINT8,
    # signed extension of the B200 build (the reference has no signed types at this commit)
    torch.qint8: DataType.INT8,
    torch.int8: DataType.INT8,
}

_QUANT_TYPES = {torch.quint2x4, torch.quint4x2, torch.quint8, torch.uint8, torch.qint8, torch.int8}
_DEQUANT_TYPES = {torch.float32, torch.bfloat16}
_ROUND_MODES = {"nearest": RoundMode.NEAREST, "stochastic": RoundMode.STOCHASTIC,
                "stochastic_per_element": RoundMode.STOCHASTIC_PER_ELEMENT}     # the last one: extension (quantize, requantize)
_REDUCE_OPS = {"set": ReduceOp.SET, "add": ReduceOp.ADD}


def torch_to_piquant_dtype(dtype: torch.dtype) -> DataType:
    if dtype not in _TORCH_DTYPE_MAP:
        raise ValueError(f"Unsupported quant_dtype: {dtype}")
    return _TORCH_DTYPE_MAP[dtype]


def piquant_to_torch_dtype(dtype: DataType) -> torch.dtype:
    for torch_dtype, piquant_dtype in _TORCH_DTYPE_MAP.items():
        if piquant_dtype == dtype:
            return torch_dtype
    raise ValueError(f"Unsupported quantized dtype: {dtype}")


def _bind_stream(ctx: Context, tensor: torch.Tensor) -> None:
    """Order the native call on torch's current stream of the tensor's device."""
    if tensor.is_cuda:
        ctx.set_stream(torch.cuda.current_stream(tensor.device).cuda_stream)


def _contiguous(tensor: torch.Tensor) -> torch.Tensor:
    return tensor if tensor.is_contiguous() else tensor.contiguous()


def compute_quant_params(tensor: torch.Tensor, *, dtype: torch.dtype, ctx: Context = Context.get()) -> Tuple[float, int]:
    """(scale, zero_point) that map [min(tensor), max(tensor)] onto the range of ``dtype``."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    tensor = _contiguous(tensor)
    _bind_stream(ctx, tensor)
    if tensor.dtype == torch.bfloat16:
        return ctx.compute_quant_params_ptr_bfloat16(tensor.data_ptr(), torch_to_piquant_dtype(dtype), tensor.numel())
    if tensor.dtype != torch.float32:
        raise ValueError(f"Unsupported input dtype: {tensor.dtype}. Must be one of {list(_DEQUANT_TYPES)}")
    return ctx.compute_quant_params_ptr_float32(tensor.data_ptr(), torch_to_piquant_dtype(dtype), tensor.numel())


def quantize(tensor: torch.Tensor, *, scale: float, zero_point: int, dtype: torch.dtype, round_mode: str = "nearest",
             ctx: Context = Context.get(), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Quantize ``tensor`` (float32 / bfloat16) to ``dtype``; the result has the input's shape and device."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    tensor = _contiguous(tensor)
    dtype_in = torch_to_piquant_dtype(tensor.dtype)
    dtype_out = torch_to_piquant_dtype(dtype)
    if out is None:
        out = torch.empty(tensor.shape, dtype=dtype, device=tensor.device)
    else:
        assert out.dtype == dtype and out.shape == tensor.shape and out.device == tensor.device and out.is_contiguous()
    _bind_stream(ctx, tensor)
    ctx.quantize_ptr(tensor.data_ptr(), dtype_in, out.data_ptr(), dtype_out, numel=tensor.numel(), scale=scale,
                     zero_point=zero_point, round_mode=_ROUND_MODES[round_mode])
    return out


def dequantize(tensor: torch.Tensor, *, scale: float, zero_point: int, dtype: torch.dtype, reduce_op: str = "set",
               ctx: Context = Context.get(), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dequantize ``tensor`` to ``dtype`` (float32 / bfloat16).  With ``reduce_op='add'`` the values
    are accumulated into ``out`` (pass the accumulator; without it the sum starts from zeros)."""
    if dtype not in _DEQUANT_TYPES:
        raise ValueError(f"Unsupported dequantized dtype: {dtype}. Must be one of {list(_DEQUANT_TYPES)}")
    tensor = _contiguous(tensor)
    if out is None:
        alloc = torch.zeros if reduce_op == "add" else torch.empty
        out = alloc(tensor.shape, dtype=dtype, device=tensor.device)
    else:
        assert out.dtype == dtype and out.shape == tensor.shape and out.device == tensor.device and out.is_contiguous()
    _bind_stream(ctx, tensor)
    ctx.dequantize_ptr(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), out.data_ptr(), torch_to_piquant_dtype(out.dtype),
                       numel=tensor.numel(), scale=scale, zero_point=zero_point, reduce_op=_REDUCE_OPS[reduce_op])
    return out


def requantize(tensor: torch.Tensor, *, scale: float, zero_point: int, dtype: torch.dtype, round_mode: str = "nearest",
               reduce_op: str = "set", ctx: Context = Context.get(), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused quantize->dequantize through quantized type ``dtype``; result has the input's float dtype."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    if tensor.dtype not in _DEQUANT_TYPES:
        raise ValueError(f"Unsupported input dtype: {tensor.dtype}. Must be one of {list(_DEQUANT_TYPES)}")
    tensor = _contiguous(tensor)
    if out is None:
        alloc = torch.zeros if reduce_op == "add" else torch.empty
        out = alloc(tensor.shape, dtype=tensor.dtype, device=tensor.device)
    else:
        assert out.dtype == tensor.dtype and out.shape == tensor.shape and out.device == tensor.device and out.is_contiguous()
    _bind_stream(ctx, tensor)
    ctx.requantize_ptr(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), out.data_ptr(), torch_to_piquant_dtype(dtype),
                       tensor.numel(), scale, zero_point, _ROUND_MODES[round_mode], _REDUCE_OPS[reduce_op])
    return out


def quantize_auto(tensor: torch.Tensor, *, dtype: torch.dtype, round_mode: str = "nearest", ctx: Context = Context.get(),
                  out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, float, int]:
    """``compute_quant_params`` followed by ``quantize`` without a host round trip in between: min/max kernel,
    parameter kernel and quantize kernel queue back to back (tensors that fit the 126 MB L2 are read from HBM
    once) and there is ONE synchronisation.  Returns ``(quantized, scale, zero_point)``; bit-identical to the
    two separate calls."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    tensor = _contiguous(tensor)
    if out is None:
        out = torch.empty(tensor.shape, dtype=dtype, device=tensor.device)
    else:
        assert out.dtype == dtype and out.shape == tensor.shape and out.device == tensor.device and out.is_contiguous()
    _bind_stream(ctx, tensor)
    scale, zero_point = ctx.quantize_auto_ptr(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), out.data_ptr(),
                                              torch_to_piquant_dtype(dtype), tensor.numel(), _ROUND_MODES[round_mode])
    return out, scale, zero_point


def new_meta(device: torch.device) -> torch.Tensor:
    """A 64-byte device block for parameters that never leave the GPU (``piquant_cuda_meta_t``)."""
    return torch.zeros(Context.META_BYTES, dtype=torch.uint8, device=device)


def meta_to_host(meta: torch.Tensor) -> Tuple[float, int]:
    """(scale, zero_point) of a meta block (synchronises)."""
    raw = meta.cpu().numpy().tobytes()
    import struct
    scale, error, zero_point = struct.unpack_from("<fiq", raw, 0)
    if error:
        raise ValueError("scale must be positive")
    return scale, zero_point
