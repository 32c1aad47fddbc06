"""piquant.torch -- tensor-level surface, device-aware.

Same three functions, keyword names, accepted dtypes and return conventions as the reference's
``piquant.torch`` (reference python/src/piquant/torch.py:9-129).  What changes for the B200 build:

* outputs are allocated on ``tensor.device`` (the reference always allocates on the CPU,
  reference torch.py:87,117);
* for CUDA tensors the work is enqueued on PyTorch's *current* stream of that device, so calls
  compose with surrounding torch ops without extra synchronisation; ``compute_quant_params``
  returns Python scalars and therefore synchronises.  Device index and stream travel WITH each call
  (``piquant_cuda_*_on_stream``): nothing is stored in the context, so threads can share one context, and
  the native side skips its per-call pointer classification;
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


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream        # int handle without building a Stream object
except AttributeError:                                     # pragma: no cover
    def _raw_stream(index: int) -> int:
        return torch.cuda.current_stream(index).cuda_stream


def _site(tensor: torch.Tensor) -> Tuple[int, int]:
    """(device, stream) of a native call on ``tensor``: a CUDA tensor runs on its own device and on torch's current stream
    of that device; a CPU tensor is classified by the library (``DEVICE_AUTO``: staged through the GPU, synchronous)."""
    if tensor.is_cuda:
        index = tensor.device.index
        return index, _raw_stream(index)
    return Context.DEVICE_AUTO, 0


def _contiguous(tensor: torch.Tensor) -> torch.Tensor:
    return tensor if tensor.is_contiguous() else tensor.contiguous()


def compute_quant_params(tensor: torch.Tensor, *, dtype: torch.dtype, ctx: Context = Context.get()) -> Tuple[float, int]:
    """(scale, zero_point) that map [min(tensor), max(tensor)] onto the range of ``dtype``."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    tensor = _contiguous(tensor)
    if tensor.dtype not in _DEQUANT_TYPES:
        raise ValueError(f"Unsupported input dtype: {tensor.dtype}. Must be one of {list(_DEQUANT_TYPES)}")
    device, stream = _site(tensor)
    return ctx.compute_quant_params_on_stream(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), tensor.numel(),
                                              torch_to_piquant_dtype(dtype), device, stream)


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
    device, stream = _site(tensor)
    ctx.quantize_on_stream(tensor.data_ptr(), dtype_in, out.data_ptr(), dtype_out, tensor.numel(), scale, zero_point,
                           _ROUND_MODES[round_mode], device, stream)
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
    device, stream = _site(tensor)
    ctx.dequantize_on_stream(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), out.data_ptr(), torch_to_piquant_dtype(out.dtype),
                             tensor.numel(), scale, zero_point, _REDUCE_OPS[reduce_op], device, stream)
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
    device, stream = _site(tensor)
    ctx.requantize_on_stream(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), out.data_ptr(), torch_to_piquant_dtype(dtype),
                             tensor.numel(), scale, zero_point, _ROUND_MODES[round_mode], _REDUCE_OPS[reduce_op], device, stream)
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
    device, stream = _site(tensor)
    scale, zero_point = ctx.quantize_auto_on_stream(tensor.data_ptr(), torch_to_piquant_dtype(tensor.dtype), out.data_ptr(),
                                                    torch_to_piquant_dtype(dtype), tensor.numel(), _ROUND_MODES[round_mode], device, stream)
    return out, scale, zero_point


def quantize_batch(tensors, *, scales, zero_points, dtype: torch.dtype, round_mode: str = "nearest", ctx: Context = Context.get(),
                   outs=None):
    """Quantize many (small) CUDA tensors of one float dtype on one device in ONE kernel launch per 256 tensors.

    The reference's own benchmark calls ``quantize`` a thousand times on 1e6 elements (reference python/benchmark/benchmark.py:16-23);
    on a GPU the launch costs more than that much data.  Every tensor gets exactly the bytes ``quantize(tensor, scale=..,
    zero_point=..)`` would produce.  ``round_mode``: ``nearest`` or ``stochastic`` (one threshold for the whole batch).
    Returns the list of quantized tensors."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    tensors = [_contiguous(t) for t in tensors]
    if not tensors:
        return []
    first = tensors[0]
    assert first.is_cuda, "quantize_batch serves CUDA tensors (host tensors: call quantize per tensor)"
    assert all(t.dtype == first.dtype and t.device == first.device for t in tensors), "one float dtype and one device per batch"
    assert len(scales) == len(tensors) and len(zero_points) == len(tensors)
    if outs is None:
        outs = [torch.empty(t.shape, dtype=dtype, device=t.device) for t in tensors]
    else:
        assert len(outs) == len(tensors) and all(o.dtype == dtype and o.shape == t.shape and o.device == t.device and o.is_contiguous()
                                                 for o, t in zip(outs, tensors))
    items = [(t.data_ptr(), o.data_ptr(), t.numel(), float(s), int(z)) for t, o, s, z in zip(tensors, outs, scales, zero_points)]
    device, stream = _site(first)
    ctx.quantize_batch(items, torch_to_piquant_dtype(first.dtype), torch_to_piquant_dtype(dtype), _ROUND_MODES[round_mode], device, stream)
    return outs


class QuantizeBatch:
    """A prepared ``quantize_batch``: the descriptor array is built once, every ``run()`` is one native call (one launch per 256
    tensors).  For sets of small tensors that are quantized again and again (the buckets of a model): the tensors, their outputs
    and their parameters are fixed at construction -- update parameters with ``set_params``."""

    def __init__(self, tensors, *, scales, zero_points, dtype: torch.dtype, ctx: Context = Context.get(), outs=None):
        assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
        self.tensors = [_contiguous(t) for t in tensors]
        assert self.tensors and self.tensors[0].is_cuda
        first = self.tensors[0]
        assert all(t.dtype == first.dtype and t.device == first.device for t in self.tensors), "one float dtype and one device per batch"
        self.outs = outs if outs is not None else [torch.empty(t.shape, dtype=dtype, device=t.device) for t in self.tensors]
        assert len(self.outs) == len(self.tensors) == len(scales) == len(zero_points)
        self.ctx, self.dtype_in, self.dtype_out = ctx, torch_to_piquant_dtype(first.dtype), torch_to_piquant_dtype(dtype)
        self.items = Context.make_batch([(t.data_ptr(), o.data_ptr(), t.numel(), float(s), int(z))
                                         for t, o, s, z in zip(self.tensors, self.outs, scales, zero_points)])

    def set_params(self, index: int, scale: float, zero_point: int) -> None:
        self.items[index].scale, self.items[index].zero_point = float(scale), int(zero_point)

    def run(self, round_mode: str = "nearest"):
        device, stream = _site(self.tensors[0])
        self.ctx.quantize_batch(self.items, self.dtype_in, self.dtype_out, _ROUND_MODES[round_mode], device, stream)
        return self.outs


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
