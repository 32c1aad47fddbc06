"""Loads libpiquant.so (the sm_100a CUDA build) through cffi in ABI mode.

Same mechanism and same exported names (``ffi``, ``C``) as the reference's loader
(reference python/src/piquant/_bootstrap.py:15-98): the library lives next to this file and is
``dlopen``-ed; nothing is compiled at import time.  The declarations are those of
``include/piquant.h`` (the reference ABI, unchanged) plus ``include/piquant_cuda.h`` (the CUDA
extensions).  There is no fallback: a missing library is an import error.
"""
from __future__ import annotations

import sys
from pathlib import Path

from cffi import FFI

LIB_NAME = "libpiquant.so"

# include/piquant.h (reference include/piquant.h:21-85) followed by include/piquant_cuda.h
_CDECLS = """
typedef struct piquant_context_t piquant_context_t;

typedef enum piquant_round_mode_t { PIQUANT_NEAREST, PIQUANT_STOCHASTIC } piquant_round_mode_t;
typedef enum piquant_reduce_op_t { PIQUANT_REDUCE_OP_SET, PIQUANT_REDUCE_OP_ADD } piquant_reduce_op_t;
typedef enum piquant_dtype_t {
    PIQUANT_DTYPE_F32 = 0, PIQUANT_DTYPE_BF16, PIQUANT_DTYPE_UINT2, PIQUANT_DTYPE_UINT4, PIQUANT_DTYPE_UINT8
} piquant_dtype_t;

piquant_context_t* piquant_context_create(size_t num_threads);
void piquant_context_destroy(piquant_context_t* ctx);
void piquant_quantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                      piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                      piquant_round_mode_t mode);
void piquant_dequantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                        piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                        piquant_reduce_op_t op);
void piquant_compute_quant_params_float32(piquant_context_t* ctx, const float* x, size_t n,
                                          piquant_dtype_t target_quant_dtype, float* out_scale, int64_t* out_zero_point);
void piquant_compute_quant_params_bfloat16(piquant_context_t* ctx, const uint16_t* x, size_t n,
                                           piquant_dtype_t target_quant_dtype, float* out_scale, int64_t* out_zero_point);

void  piquant_cuda_set_stream(piquant_context_t* ctx, void* cuda_stream);
void* piquant_cuda_get_stream(piquant_context_t* ctx);
void  piquant_cuda_synchronize(piquant_context_t* ctx);
void  piquant_cuda_set_kernel_variant(piquant_context_t* ctx, int variant);
uint64_t piquant_cuda_kernel_launches(piquant_context_t* ctx);
int   piquant_cuda_device_count(void);
void  piquant_cuda_set_stochastic_threshold(piquant_context_t* ctx, float xi);
void  piquant_cuda_seed(piquant_context_t* ctx, uint64_t seed);
float piquant_cuda_last_stochastic_threshold(piquant_context_t* ctx);
void  piquant_cuda_requantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in_out, void* out,
                              piquant_dtype_t quant_dtype, size_t numel, float scale, int64_t zero_point,
                              piquant_round_mode_t mode, piquant_reduce_op_t op);
void  piquant_cuda_minmax_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n, float* out4);
void  piquant_cuda_params_from_minmax(float min, float max, piquant_dtype_t target_quant_dtype,
                                      float* out_scale, int64_t* out_zero_point);
void  piquant_cuda_set_sr_key(piquant_context_t* ctx, uint64_t key);
void  piquant_cuda_clear_sr_key(piquant_context_t* ctx);
uint64_t piquant_cuda_last_sr_key(piquant_context_t* ctx);
int   piquant_cuda_nccl_unique_id(void* out128);
void  piquant_cuda_comm_init_rank(piquant_context_t* ctx, const void* unique_id128, int nranks, int rank);
void  piquant_cuda_comm_destroy(piquant_context_t* ctx);

typedef struct piquant_cuda_meta_t { float scale; int32_t error; int64_t zero_point; unsigned char opaque[48]; } piquant_cuda_meta_t;
void  piquant_cuda_compute_meta_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                      piquant_dtype_t target_quant_dtype, piquant_cuda_meta_t* d_meta);
void  piquant_cuda_quantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                       piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                       const piquant_cuda_meta_t* d_meta);
void  piquant_cuda_dequantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                         piquant_dtype_t dtype_out, size_t numel, piquant_reduce_op_t op,
                                         const piquant_cuda_meta_t* d_meta);
void  piquant_cuda_quantize_auto(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                 piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                 float* out_scale, int64_t* out_zero_point);
void  piquant_cuda_comm_set_transport(piquant_context_t* ctx, int transport);
int   piquant_cuda_comm_transport(piquant_context_t* ctx);

/* The *_on_stream family (include/piquant_cuda.h).  Data pointers are declared uintptr_t HERE: on x86-64 that is the same
   ABI as `const void*`, and Python ints (tensor.data_ptr()) then pass without an ffi.cast per argument. */
void  piquant_cuda_quantize_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel,
                                      float scale, int64_t zero_point, int mode, int device, uintptr_t stream);
void  piquant_cuda_dequantize_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel,
                                        float scale, int64_t zero_point, int op, int device, uintptr_t stream);
void  piquant_cuda_requantize_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in_out, uintptr_t out, int quant_dtype, size_t numel,
                                        float scale, int64_t zero_point, int mode, int op, int device, uintptr_t stream);
void  piquant_cuda_compute_quant_params_on_stream(piquant_context_t* ctx, uintptr_t x, int dtype, size_t n, int target_quant_dtype,
                                                  float* out_scale, int64_t* out_zero_point, int device, uintptr_t stream);
void  piquant_cuda_minmax_on_stream(piquant_context_t* ctx, uintptr_t x, int dtype, size_t n, uintptr_t out4, unsigned flags, int device, uintptr_t stream);
void  piquant_cuda_compute_meta_on_stream(piquant_context_t* ctx, uintptr_t x, int dtype, size_t n, int target_quant_dtype, uintptr_t d_meta,
                                          unsigned flags, int device, uintptr_t stream);
void  piquant_cuda_quantize_meta_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel, int mode,
                                           uintptr_t d_meta, unsigned flags, int device, uintptr_t stream);
void  piquant_cuda_dequantize_meta_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel, int op,
                                             uintptr_t d_meta, int device, uintptr_t stream);
void  piquant_cuda_quantize_auto_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel, int mode,
                                           float* out_scale, int64_t* out_zero_point, int device, uintptr_t stream);
void  piquant_cuda_dequantize_add_minmax_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel,
                                                   uintptr_t d_meta, int next_quant_dtype, uintptr_t d_meta_next, uintptr_t d_meta_next_copy,
                                                   int device, uintptr_t stream);
void  piquant_cuda_dequantize_forward_on_stream(piquant_context_t* ctx, uintptr_t in, int dtype_in, uintptr_t out, int dtype_out, size_t numel,
                                                uintptr_t d_meta, uintptr_t forward_to, uintptr_t forward_meta_to, int device, uintptr_t stream);
void  piquant_cuda_dequantize_sum_minmax_on_stream(piquant_context_t* ctx, const uintptr_t* ins, const uintptr_t* d_metas, size_t count, int dtype_in,
                                                   uintptr_t out, int dtype_out, size_t numel, int next_quant_dtype, uintptr_t d_meta_next,
                                                   uintptr_t d_meta_next_copy, int device, uintptr_t stream);
void  piquant_cuda_wait_flag_on_stream(piquant_context_t* ctx, uintptr_t flag, int device, uintptr_t stream);
void  piquant_cuda_copy_on_stream(piquant_context_t* ctx, uintptr_t dst, uintptr_t src, size_t nbytes, int device, uintptr_t stream);
typedef struct piquant_cuda_batch_item_t { uintptr_t in; uintptr_t out; size_t numel; float scale; int64_t zero_point; } piquant_cuda_batch_item_t;
void  piquant_cuda_quantize_batch(piquant_context_t* ctx, const piquant_cuda_batch_item_t* items, size_t count, int dtype_in, int dtype_out,
                                  int mode, int device, uintptr_t stream);
"""


def _load():
    assert sys.platform.startswith("linux"), f"the B200 build of piquant is Linux-only, not {sys.platform}"
    lib_path = Path(__file__).resolve().parent / LIB_NAME
    if not lib_path.exists():
        raise ImportError(
            f"{lib_path} not found: build it with `python pi-quant_b200/build.py` (needs nvcc). "
            "There is no CPU fallback behind the piquant CUDA package."
        )
    ffi = FFI()
    ffi.cdef(_CDECLS)
    return ffi, ffi.dlopen(str(lib_path))


ffi, C = _load()
