/*
 * piquant.h -- the C99 ABI of the B200-native pi-quant library (libpiquant.so).
 *
 * DROP-IN BOUNDARY.  The six entry points below have exactly the names, argument order, argument
 * types and enum values of the reference's C API (reference include/piquant.h:23-85, implemented by
 * reference src/capi.cpp:19-104), so the reference's own CFFI binding
 * (reference python/src/piquant/_bootstrap.py:15-98) and any C caller bind to this library
 * unchanged.  What differs is where the work runs: every call is dispatched to hand-written
 * sm_100a CUDA kernels on a CUDA stream instead of an AVX thread pool.
 *
 * Buffers may be CUDA device memory, pinned (page-locked) host memory or ordinary host memory:
 *   - device pointers : the kernels run in place, asynchronously, ordered on the context's stream
 *                       (default: the legacy default stream, i.e. ordered with PyTorch's default
 *                       stream); piquant_compute_quant_params_* synchronises because it returns
 *                       host scalars;
 *   - host pointers   : the call is synchronous like the reference's; data is streamed through
 *                       the GPU in double-buffered chunks (H2D copy | kernel | D2H copy overlap).
 * There is no CPU implementation behind this ABI: without a usable CUDA device the first compute
 * call aborts.  piquant_context_create itself never touches CUDA (the reference's Python package
 * creates a context at import time, reference python/src/piquant/torch.py:57).
 *
 * Errors: like the reference (reference src/piquant.cpp:88-98) every violated precondition and
 * every CUDA error prints a red message to stderr and calls abort(); all functions return void.
 *
 * CUDA-specific controls (stream, explicit stochastic threshold, NCCL communicator for sharded
 * tensors, fused quantize->dequantize) live in piquant_cuda.h so that this header stays
 * ABI-identical to the reference's.
 */
#ifndef PIQUANT_H
#define PIQUANT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_MSC_VER)
#define PIQUANT_EXPORT __declspec(dllexport)
#else
#define PIQUANT_EXPORT __attribute__((visibility("default")))
#endif

/* Opaque handle; replaces the reference's context (reference include/piquant.hpp:199-339). */
typedef struct piquant_context_t piquant_context_t;

/* reference include/piquant.h:23-26.  STOCHASTIC draws ONE threshold per call, shared by every
 * element (reference src/piquant.cpp:199-201). */
typedef enum piquant_round_mode_t {
    PIQUANT_NEAREST = 0,
    PIQUANT_STOCHASTIC = 1
} piquant_round_mode_t;

/* reference include/piquant.h:28-31.  ADD accumulates the dequantized value into `out`. */
typedef enum piquant_reduce_op_t {
    PIQUANT_REDUCE_OP_SET = 0,
    PIQUANT_REDUCE_OP_ADD = 1
} piquant_reduce_op_t;

/* reference include/piquant.h:33-40.  UINT2 / UINT4 are bit-packed, element k of a byte in bits
 * [k*b, k*b+b) (low element in the low bits); a tensor of n elements occupies ceil(n*b/8) bytes. */
typedef enum piquant_dtype_t {
    PIQUANT_DTYPE_F32 = 0,
    PIQUANT_DTYPE_BF16 = 1,
    PIQUANT_DTYPE_UINT2 = 2,
    PIQUANT_DTYPE_UINT4 = 3,
    PIQUANT_DTYPE_UINT8 = 4
} piquant_dtype_t;

/* reference include/piquant.h:42-43.  num_threads is accepted for compatibility; the GPU grid
 * replaces the thread pool. */
PIQUANT_EXPORT piquant_context_t* piquant_context_create(size_t num_threads);
PIQUANT_EXPORT void piquant_context_destroy(piquant_context_t* ctx);

/* reference include/piquant.h:45-55.
 * out[i] = clamp(round(in[i] / scale) + zero_point, 0, 2^bits - 1), packed.
 * dtype_in must be F32 or BF16, dtype_out UINT2/UINT4/UINT8; numel counts logical elements. */
PIQUANT_EXPORT void piquant_quantize(
    piquant_context_t* ctx,
    const void* in, piquant_dtype_t dtype_in,
    void* out, piquant_dtype_t dtype_out,
    size_t numel,
    float scale, int64_t zero_point,
    piquant_round_mode_t mode);

/* reference include/piquant.h:57-67.
 * SET: out[i] = (in[i] - zero_point) * scale;  ADD: out[i] += (in[i] - zero_point) * scale.
 * dtype_in must be UINT2/UINT4/UINT8, dtype_out F32 or BF16; numel counts logical elements. */
PIQUANT_EXPORT void piquant_dequantize(
    piquant_context_t* ctx,
    const void* in, piquant_dtype_t dtype_in,
    void* out, piquant_dtype_t dtype_out,
    size_t numel,
    float scale, int64_t zero_point,
    piquant_reduce_op_t op);

/* reference include/piquant.h:69-76.  min/max reduce of x -> (scale, zero_point) for the target
 * quantized dtype; aborts for n == 0 (like the reference). */
PIQUANT_EXPORT void piquant_compute_quant_params_float32(
    piquant_context_t* ctx,
    const float* x, size_t n,
    piquant_dtype_t target_quant_dtype,
    float* out_scale, int64_t* out_zero_point);

/* reference include/piquant.h:78-85.  x holds raw bf16 bit patterns. */
PIQUANT_EXPORT void piquant_compute_quant_params_bfloat16(
    piquant_context_t* ctx,
    const uint16_t* x, size_t n,
    piquant_dtype_t target_quant_dtype,
    float* out_scale, int64_t* out_zero_point);

#ifdef __cplusplus
}
#endif
#endif /* PIQUANT_H */
