"""piquant -- B200-native drop-in for the pi-quant Python package.

Mirrors the public names of the reference package (reference python/src/piquant/__init__.py:20-142):
``RoundMode``, ``ReduceOp``, ``DataType`` and ``Context`` with its ``*_ptr`` methods, bound to
``libpiquant.so`` through the same C ABI.  Pointers may be CUDA device pointers (the normal case
here) or host pointers; see ``include/piquant.h`` for the semantics of each.
"""
from __future__ import annotations

__version__ = "0.1.0+b200"

import importlib.util
import multiprocessing
import weakref
from enum import Enum, unique
from functools import lru_cache
from typing import Optional, Tuple, Union

from piquant._bootstrap import C, ffi


@unique
class RoundMode(Enum):
    NEAREST = C.PIQUANT_NEAREST
    STOCHASTIC = C.PIQUANT_STOCHASTIC
    # extension of the B200 build (include/piquant_cuda.h): every element gets its own Philox4x32-10 random number
    # instead of the reference's one threshold per call; quantize only
    STOCHASTIC_PER_ELEMENT = 2


@unique
class ReduceOp(Enum):
    SET = C.PIQUANT_REDUCE_OP_SET
    ADD = C.PIQUANT_REDUCE_OP_ADD


@unique
class DataType(Enum):
    F32 = C.PIQUANT_DTYPE_F32
    BF16 = C.PIQUANT_DTYPE_BF16
    UINT2 = C.PIQUANT_DTYPE_UINT2
    UINT4 = C.PIQUANT_DTYPE_UINT4
    UINT8 = C.PIQUANT_DTYPE_UINT8
    # signed extension of the B200 build (include/piquant_cuda.h: PIQUANT_CUDA_DTYPE_INT2/4/8): two's complement
    # fields, same packing order; defined as the offset-binary view of the unsigned types
    INT2 = 5
    INT4 = 6
    INT8 = 7

    @property
    def bit_size(self) -> int:
        return {DataType.F32: 32, DataType.BF16: 16, DataType.UINT2: 2, DataType.UINT4: 4, DataType.UINT8: 8,
                DataType.INT2: 2, DataType.INT4: 4, DataType.INT8: 8}[self]

    @property
    def is_quantized(self) -> bool:
        return self in (DataType.UINT2, DataType.UINT4, DataType.UINT8, DataType.INT2, DataType.INT4, DataType.INT8)

    @property
    def is_signed(self) -> bool:
        return self in (DataType.INT2, DataType.INT4, DataType.INT8)

    @property
    def is_dequantized(self) -> bool:
        return self in (DataType.F32, DataType.BF16)

    @property
    def stride(self) -> int:
        """Bytes of one storage unit (a packed byte for the sub-byte types)."""
        return max(8, self.bit_size) >> 3

    def storage_bytes(self, numel: int) -> int:
        """Bytes a contiguous tensor of ``numel`` elements occupies (packed for UINT2/UINT4)."""
        if self.is_quantized:
            per = 8 // self.bit_size
            return (numel + per - 1) // per
        return numel * (self.bit_size >> 3)


class Context:
    """Owns the dispatcher state of the native library (stream, scratch, stochastic RNG, NCCL comm).

    ``num_threads`` is accepted for source compatibility with the reference
    (reference python/src/piquant/__init__.py:64-71); the CUDA grid replaces the thread pool.
    Creating a context never initialises CUDA -- the default one is created at import time.
    """

    def __init__(self, num_threads: Union[int, None] = None) -> None:
        if num_threads is None:
            num_threads = max(multiprocessing.cpu_count() - 1, 1)
        self._num_threads = num_threads
        self._ctx = C.piquant_context_create(self._num_threads)
        self._finalizer = weakref.finalize(self, C.piquant_context_destroy, self._ctx)

    @staticmethod
    @lru_cache(maxsize=1)
    def get() -> "Context":
        """Process-wide default context."""
        return Context()

    # ---- the reference's pointer-level API --------------------------------------------------------

    def quantize_ptr(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                     scale: float, zero_point: int, round_mode: RoundMode) -> None:
        assert dtype_in.is_dequantized, f"Input dtype must be a dequantized type, but is: {dtype_in}"
        assert dtype_out.is_quantized, f"Output dtype must be a quantized type, but is: {dtype_out}"
        assert ptr_in != 0, "Input arr pointer must not be NULL"
        assert ptr_out != 0, "Output arr pointer must not be NULL"
        C.piquant_quantize(self._ctx, ffi.cast("const void*", ptr_in), dtype_in.value, ffi.cast("void*", ptr_out),
                           dtype_out.value, numel, scale, zero_point, round_mode.value)

    def dequantize_ptr(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                       scale: float, zero_point: int, reduce_op: ReduceOp) -> None:
        assert dtype_in.is_quantized, f"Input dtype must be a quantized type, but is: {dtype_in}"
        assert dtype_out.is_dequantized, f"Output dtype must be a dequantized type, but is: {dtype_out}"
        assert ptr_in != 0, "Input arr pointer must not be NULL"
        assert ptr_out != 0, "Output arr pointer must not be NULL"
        C.piquant_dequantize(self._ctx, ffi.cast("const void*", ptr_in), dtype_in.value, ffi.cast("void*", ptr_out),
                             dtype_out.value, numel, scale, zero_point, reduce_op.value)

    def compute_quant_params_ptr_float32(self, ptr: int, target_quant_dtype: DataType, numel: int) -> Tuple[float, int]:
        assert target_quant_dtype.is_quantized, f"Target dtype must be a quantized type, but is: {target_quant_dtype}"
        assert ptr != 0, "Input arr pointer must not be NULL"
        scale, zero_point = ffi.new("float*"), ffi.new("int64_t*")
        C.piquant_compute_quant_params_float32(self._ctx, ffi.cast("const float*", ptr), numel, target_quant_dtype.value,
                                               scale, zero_point)
        return scale[0], zero_point[0]

    def compute_quant_params_ptr_bfloat16(self, ptr: int, target_quant_dtype: DataType, numel: int) -> Tuple[float, int]:
        assert target_quant_dtype.is_quantized, f"Target dtype must be a quantized type, but is: {target_quant_dtype}"
        assert ptr != 0, "Input arr pointer must not be NULL"
        scale, zero_point = ffi.new("float*"), ffi.new("int64_t*")
        C.piquant_compute_quant_params_bfloat16(self._ctx, ffi.cast("const uint16_t*", ptr), numel,
                                                target_quant_dtype.value, scale, zero_point)
        return scale[0], zero_point[0]

    # ---- CUDA extensions (include/piquant_cuda.h) -------------------------------------------------

    def requantize_ptr(self, ptr_in: int, dtype_in_out: DataType, ptr_out: int, quant_dtype: DataType, numel: int,
                       scale: float, zero_point: int, round_mode: RoundMode = RoundMode.NEAREST,
                       reduce_op: ReduceOp = ReduceOp.SET) -> None:
        """Fused quantize->dequantize (reference C++ API: context::quantize_dequantize_fused)."""
        assert dtype_in_out.is_dequantized and quant_dtype.is_quantized
        assert ptr_in != 0 and ptr_out != 0
        C.piquant_cuda_requantize(self._ctx, ffi.cast("const void*", ptr_in), dtype_in_out.value,
                                  ffi.cast("void*", ptr_out), quant_dtype.value, numel, scale, zero_point,
                                  round_mode.value, reduce_op.value)

    def set_stream(self, cuda_stream: int) -> None:
        """Order device-pointer calls on this ``cudaStream_t`` (0 = legacy default stream)."""
        C.piquant_cuda_set_stream(self._ctx, ffi.cast("void*", cuda_stream))

    def get_stream(self) -> int:
        return int(ffi.cast("uintptr_t", C.piquant_cuda_get_stream(self._ctx)))

    def synchronize(self) -> None:
        C.piquant_cuda_synchronize(self._ctx)

    def set_kernel_variant(self, variant: int) -> None:
        """0 = auto, 1 = direct LDG/STG kernels, 2 = TMA ring kernels."""
        C.piquant_cuda_set_kernel_variant(self._ctx, variant)

    @property
    def kernel_launches(self) -> int:
        return int(C.piquant_cuda_kernel_launches(self._ctx))

    def set_stochastic_threshold(self, xi: Optional[float]) -> None:
        """Fix the per-call stochastic threshold (None: draw a fresh one per call, like the reference)."""
        C.piquant_cuda_set_stochastic_threshold(self._ctx, -1.0 if xi is None else xi)

    def set_sr_key(self, key: Optional[int]) -> None:
        """Per-element stochastic rounding: use this 64-bit Philox key for every following call; None draws one per call again."""
        if key is None:
            C.piquant_cuda_clear_sr_key(self._ctx)
        else:
            C.piquant_cuda_set_sr_key(self._ctx, key & 0xFFFFFFFFFFFFFFFF)

    @property
    def last_sr_key(self) -> int:
        return int(C.piquant_cuda_last_sr_key(self._ctx))

    def seed(self, seed: int) -> None:
        C.piquant_cuda_seed(self._ctx, seed)

    @property
    def last_stochastic_threshold(self) -> float:
        return float(C.piquant_cuda_last_stochastic_threshold(self._ctx))

    def minmax_async_ptr(self, ptr: int, dtype: DataType, numel: int, ptr_out4: int) -> None:
        C.piquant_cuda_minmax_async(self._ctx, ffi.cast("const void*", ptr), dtype.value, numel,
                                    ffi.cast("float*", ptr_out4))

    @staticmethod
    def params_from_minmax(mn: float, mx: float, target_quant_dtype: DataType) -> Tuple[float, int]:
        scale, zero_point = ffi.new("float*"), ffi.new("int64_t*")
        C.piquant_cuda_params_from_minmax(mn, mx, target_quant_dtype.value, scale, zero_point)
        return scale[0], zero_point[0]

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ffi.new("char[128]")
        if C.piquant_cuda_nccl_unique_id(buf) != 0:
            raise RuntimeError("libnccl could not be loaded by libpiquant.so")
        return bytes(ffi.buffer(buf, 128))

    def comm_init_rank(self, unique_id: bytes, nranks: int, rank: int) -> None:
        assert len(unique_id) == 128
        C.piquant_cuda_comm_init_rank(self._ctx, ffi.from_buffer(unique_id), nranks, rank)

    def comm_destroy(self) -> None:
        C.piquant_cuda_comm_destroy(self._ctx)

    # ---- device-resident parameters (include/piquant_cuda.h: piquant_cuda_meta_t, 64 bytes of device memory) ----

    META_BYTES = 64

    def compute_meta_async_ptr(self, ptr: int, dtype: DataType, numel: int, target_quant_dtype: DataType, ptr_meta: int) -> None:
        assert dtype.is_dequantized and target_quant_dtype.is_quantized and ptr_meta != 0
        C.piquant_cuda_compute_meta_async(self._ctx, ffi.cast("const void*", ptr), dtype.value, numel, target_quant_dtype.value,
                                          ffi.cast("piquant_cuda_meta_t*", ptr_meta))

    def quantize_meta_async_ptr(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                round_mode: RoundMode, ptr_meta: int) -> None:
        assert dtype_in.is_dequantized and dtype_out.is_quantized and ptr_meta != 0
        C.piquant_cuda_quantize_meta_async(self._ctx, ffi.cast("const void*", ptr_in), dtype_in.value, ffi.cast("void*", ptr_out),
                                           dtype_out.value, numel, round_mode.value, ffi.cast("const piquant_cuda_meta_t*", ptr_meta))

    def dequantize_meta_async_ptr(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                  reduce_op: ReduceOp, ptr_meta: int) -> None:
        assert dtype_in.is_quantized and dtype_out.is_dequantized and ptr_meta != 0
        C.piquant_cuda_dequantize_meta_async(self._ctx, ffi.cast("const void*", ptr_in), dtype_in.value, ffi.cast("void*", ptr_out),
                                             dtype_out.value, numel, reduce_op.value, ffi.cast("const piquant_cuda_meta_t*", ptr_meta))

    def quantize_auto_ptr(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                          round_mode: RoundMode = RoundMode.NEAREST) -> Tuple[float, int]:
        """compute_quant_params + quantize in one call with a single synchronisation; returns (scale, zero_point)."""
        assert dtype_in.is_dequantized and dtype_out.is_quantized and ptr_in != 0 and ptr_out != 0
        scale, zero_point = ffi.new("float*"), ffi.new("int64_t*")
        C.piquant_cuda_quantize_auto(self._ctx, ffi.cast("const void*", ptr_in), dtype_in.value, ffi.cast("void*", ptr_out),
                                     dtype_out.value, numel, round_mode.value, scale, zero_point)
        return scale[0], zero_point[0]


    # ---- explicit device + stream per call (include/piquant_cuda.h: the *_on_stream family) ---------------------
    # device >= 0: the caller vouches that every pointer is device memory of that device (no driver query per call);
    # DEVICE_AUTO: classify the pointers like the reference-ABI calls do.  Nothing here touches context state, so
    # threads may share one context freely.

    DEVICE_AUTO = -1
    FLAG_LOCAL, FLAG_KEEP_IN_L2, FLAG_REVERSE = 1, 2, 4
    MAX_SUM_SOURCES = 8          # PIQUANT_CUDA_MAX_SUM_SOURCES

    def quantize_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int, scale: float,
                           zero_point: int, round_mode: RoundMode, device: int, stream: int) -> None:
        C.piquant_cuda_quantize_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, scale, zero_point,
                                          round_mode.value, device, stream)

    def dequantize_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int, scale: float,
                             zero_point: int, reduce_op: ReduceOp, device: int, stream: int) -> None:
        C.piquant_cuda_dequantize_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, scale, zero_point,
                                            reduce_op.value, device, stream)

    def requantize_on_stream(self, ptr_in: int, dtype_in_out: DataType, ptr_out: int, quant_dtype: DataType, numel: int, scale: float,
                             zero_point: int, round_mode: RoundMode, reduce_op: ReduceOp, device: int, stream: int) -> None:
        C.piquant_cuda_requantize_on_stream(self._ctx, ptr_in, dtype_in_out.value, ptr_out, quant_dtype.value, numel, scale, zero_point,
                                            round_mode.value, reduce_op.value, device, stream)

    def compute_quant_params_on_stream(self, ptr: int, dtype: DataType, numel: int, target_quant_dtype: DataType, device: int,
                                       stream: int) -> Tuple[float, int]:
        scale, zero_point = ffi.new("float*"), ffi.new("int64_t*")
        C.piquant_cuda_compute_quant_params_on_stream(self._ctx, ptr, dtype.value, numel, target_quant_dtype.value, scale, zero_point,
                                                      device, stream)
        return scale[0], zero_point[0]

    def minmax_on_stream(self, ptr: int, dtype: DataType, numel: int, ptr_out4: int, flags: int, device: int, stream: int) -> None:
        C.piquant_cuda_minmax_on_stream(self._ctx, ptr, dtype.value, numel, ptr_out4, flags, device, stream)

    def compute_meta_on_stream(self, ptr: int, dtype: DataType, numel: int, target_quant_dtype: DataType, ptr_meta: int, flags: int,
                               device: int, stream: int) -> None:
        C.piquant_cuda_compute_meta_on_stream(self._ctx, ptr, dtype.value, numel, target_quant_dtype.value, ptr_meta, flags, device, stream)

    def quantize_meta_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                round_mode: RoundMode, ptr_meta: int, flags: int, device: int, stream: int) -> None:
        C.piquant_cuda_quantize_meta_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, round_mode.value,
                                               ptr_meta, flags, device, stream)

    def dequantize_meta_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                  reduce_op: ReduceOp, ptr_meta: int, device: int, stream: int) -> None:
        C.piquant_cuda_dequantize_meta_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, reduce_op.value,
                                                 ptr_meta, device, stream)

    def quantize_auto_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                round_mode: RoundMode, device: int, stream: int) -> Tuple[float, int]:
        scale, zero_point = ffi.new("float*"), ffi.new("int64_t*")
        C.piquant_cuda_quantize_auto_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, round_mode.value,
                                               scale, zero_point, device, stream)
        return scale[0], zero_point[0]

    def dequantize_add_minmax_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                        ptr_meta: int, next_quant_dtype: DataType, ptr_meta_next: int, ptr_meta_next_copy: int,
                                        device: int, stream: int) -> None:
        """out += dequantize(in) and, in the same launch, the parameters of the sums for the next hop (ring reduce-scatter)."""
        C.piquant_cuda_dequantize_add_minmax_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, ptr_meta,
                                                       next_quant_dtype.value, ptr_meta_next, ptr_meta_next_copy, device, stream)

    def dequantize_forward_on_stream(self, ptr_in: int, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int,
                                     ptr_meta: int, ptr_forward: int, ptr_forward_meta: int, device: int, stream: int) -> None:
        """out = dequantize(in) and, in the same launch, the packed bytes (+ parameter block) stored on to the next rank (ring all-gather)."""
        C.piquant_cuda_dequantize_forward_on_stream(self._ctx, ptr_in, dtype_in.value, ptr_out, dtype_out.value, numel, ptr_meta,
                                                    ptr_forward, ptr_forward_meta, device, stream)

    def dequantize_sum_minmax_on_stream(self, ptrs_in, dtype_in: DataType, ptr_out: int, dtype_out: DataType, numel: int, ptrs_meta,
                                        next_quant_dtype: DataType, ptr_meta_next: int, ptr_meta_next_copy: int, device: int, stream: int) -> None:
        """``out += dequantize(ptrs_in[0]) + dequantize(ptrs_in[1]) + ...`` (sources folded in order, bit-identical to that many
        ADD calls) in ONE pass, plus min/max + parameters of the sums for ``next_quant_dtype`` -> ``ptr_meta_next``."""
        ins, metas = ffi.new("uintptr_t[]", list(ptrs_in)), ffi.new("uintptr_t[]", list(ptrs_meta))
        assert len(ins) == len(metas)
        C.piquant_cuda_dequantize_sum_minmax_on_stream(self._ctx, ins, metas, len(ins), dtype_in.value, ptr_out, dtype_out.value, numel,
                                                       next_quant_dtype.value, ptr_meta_next, ptr_meta_next_copy, device, stream)

    def wait_flag_on_stream(self, ptr_flag: int, device: int, stream: int) -> None:
        """Stream-ordered wait until the 4-byte flag at ``ptr_flag`` is non-zero (set by a copy engine after its payload); resets it."""
        C.piquant_cuda_wait_flag_on_stream(self._ctx, ptr_flag, device, stream)

    def copy_on_stream(self, ptr_dst: int, ptr_src: int, nbytes: int, device: int, stream: int) -> None:
        """Stream-ordered copy by a copy engine (cudaMemcpyAsync): local, peer-mapped or pinned memory on either side."""
        C.piquant_cuda_copy_on_stream(self._ctx, ptr_dst, ptr_src, nbytes, device, stream)

    @staticmethod
    def make_batch(items):
        """The native descriptor array of a batch: items = sequence of (ptr_in, ptr_out, numel, scale, zero_point).  Build it once
        for tensors that are quantized repeatedly (converting a thousand Python tuples costs more than quantizing them)."""
        return ffi.new("piquant_cuda_batch_item_t[]", items if isinstance(items, list) else list(items))     # tuples initialise the structs

    def quantize_batch(self, items, dtype_in: DataType, dtype_out: DataType, round_mode: RoundMode, device: int, stream: int) -> None:
        """items: sequence of (ptr_in, ptr_out, numel, scale, zero_point) or the result of ``make_batch``; ONE kernel launch per 256 tensors."""
        arr = items if isinstance(items, ffi.CData) else Context.make_batch(items)
        C.piquant_cuda_quantize_batch(self._ctx, arr, len(arr), dtype_in.value, dtype_out.value, round_mode.value, device, stream)

    def comm_set_transport(self, transport: int) -> None:
        """0 = peer-memory exchange inside the min/max kernel when available (default), 1 = ncclAllReduce, 2 = peer memory or abort."""
        C.piquant_cuda_comm_set_transport(self._ctx, transport)

    @property
    def comm_transport(self) -> int:
        return int(C.piquant_cuda_comm_transport(self._ctx))


def cuda_device_count() -> int:
    return int(C.piquant_cuda_device_count())


if importlib.util.find_spec("torch") is not None:
    try:
        from . import torch  # noqa: F401
    except ImportError:
        pass
