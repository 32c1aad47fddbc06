/*
 * piquant_cuda.h -- CUDA-specific extensions of libpiquant.so.  Everything here is NEW surface:
 * the reference has no stream, no device, no collective.  piquant.h stays ABI-identical to the
 * reference; callers that only know the reference never need this header.
 *
 * All functions abort() on error like the rest of the library unless stated otherwise.
 */
#ifndef PIQUANT_CUDA_H
#define PIQUANT_CUDA_H

#include "piquant.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- stream dispatcher (replaces the reference's thread pool, reference src/piquant.cpp:197-211) */

/* CUDA stream (a cudaStream_t) that device-pointer calls are ordered on.  NULL = legacy default stream. */
PIQUANT_EXPORT void  piquant_cuda_set_stream(piquant_context_t* ctx, void* cuda_stream);
PIQUANT_EXPORT void* piquant_cuda_get_stream(piquant_context_t* ctx);
/* Block until everything this context has enqueued on its stream has finished. */
PIQUANT_EXPORT void  piquant_cuda_synchronize(piquant_context_t* ctx);
/* 0 = choose per cell (default), 1 = direct LDG/STG streaming kernels, 2 = TMA (cp.async.bulk) ring kernels. */
PIQUANT_EXPORT void  piquant_cuda_set_kernel_variant(piquant_context_t* ctx, int variant);
/* Number of CUDA kernels this context has launched so far (introspection for benchmarks and tests). */
PIQUANT_EXPORT uint64_t piquant_cuda_kernel_launches(piquant_context_t* ctx);
/* Number of usable CUDA devices; 0 when there is none (never aborts). */
PIQUANT_EXPORT int   piquant_cuda_device_count(void);

/* ---- signed quantized dtypes (extension) -----------------------------------------------------
 * The reference at this commit has only UINT2/4/8 (reference include/piquant.h:33-40), but its parameter code
 * already carries the signed case (compute_type_max's is_signed branch and type_min = -type_max - 1,
 * reference src/piquant.cpp:212-220,246-248).  These three values are accepted wherever piquant.h takes a
 * quantized dtype (quantize output, dequantize input, compute_quant_params target, requantize / *_meta_async / auto).
 *
 * Definition: intN is the offset-binary view of uintN --
 *     quantize(x -> intN; scale, zp)   == quantize(x -> uintN; scale, zp + 2^(N-1)) with the sign bit of every field flipped,
 *     dequantize(q: intN; scale, zp)   == dequantize(q with sign bits flipped: uintN; scale, zp + 2^(N-1)),
 * i.e. q = clamp(round(x / scale) + zp, -2^(N-1), 2^(N-1) - 1) stored in two's complement, same packing order as the
 * unsigned types (element k of a byte at bits [k*N, k*N+N)); every rounding rule and corner case of the unsigned kernels
 * carries over.  compute_quant_params follows the reference's formula with q_min = -2^(N-1) (constant input: zero point -1). */
#define PIQUANT_CUDA_DTYPE_INT2 ((piquant_dtype_t)5)
#define PIQUANT_CUDA_DTYPE_INT4 ((piquant_dtype_t)6)
#define PIQUANT_CUDA_DTYPE_INT8 ((piquant_dtype_t)7)

/* ---- stochastic rounding control.  The reference draws its per-call threshold from a
 * random_device-seeded generator that no API can reach (reference src/piquant.cpp:194-201). */

/* xi in [0,1): use this threshold for every following STOCHASTIC call; xi < 0: draw per call again. */
PIQUANT_EXPORT void  piquant_cuda_set_stochastic_threshold(piquant_context_t* ctx, float xi);
/* Reseed the per-call threshold generator (reproducible runs). */
PIQUANT_EXPORT void  piquant_cuda_seed(piquant_context_t* ctx, uint64_t seed);
/* The threshold used by the most recent STOCHASTIC call. */
PIQUANT_EXPORT float piquant_cuda_last_stochastic_threshold(piquant_context_t* ctx);

/* ---- per-element stochastic rounding (extension) ------------------------------------------------
 * The reference's STOCHASTIC mode compares every element of a call with ONE threshold (reference src/piquant.cpp:199-201):
 * cheap, but every call is biased.  This mode is accepted as the `mode` of piquant_quantize, piquant_cuda_quantize_meta_async,
 * piquant_cuda_quantize_auto and piquant_cuda_requantize and rounds every element with its own random number:
 *
 *     q_i = clamp(floor(x_i / scale + u_i) + zero_point, qmin, qmax),      u_i = (k_i + 1/2) * 2^-16
 *     k_i = 16 bits of Philox4x32-10(counter = {lo32(i / 8), hi32(i / 8), 0, 0}, key = {lo32(key), hi32(key)}):
 *           element i takes bits [16 * (i % 2), +16) of output word (i % 8) / 2,
 *
 * i = index of the element in the tensor passed to the call, x_i / scale = RN(x_i * (1.0f / scale)) as in every other mode.
 * E[q_i] = x_i / scale + zero_point inside the range (unbiased to 2^-17 of a step); integers never move.  The 64-bit key
 * of a call comes from the context's generator (reproducible after piquant_cuda_seed) unless one is set here. */
#define PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT ((piquant_round_mode_t)2)
/* Use this key for every following per-element call (tests, replay). */
PIQUANT_EXPORT void     piquant_cuda_set_sr_key(piquant_context_t* ctx, uint64_t key);
/* Draw a fresh key per call again. */
PIQUANT_EXPORT void     piquant_cuda_clear_sr_key(piquant_context_t* ctx);
/* The key used by the most recent per-element call. */
PIQUANT_EXPORT uint64_t piquant_cuda_last_sr_key(piquant_context_t* ctx);

/* ---- fused quantize -> dequantize (reference C++ API only: context::quantize_dequantize_fused,
 * reference include/piquant.hpp:276-285, src/piquant.cpp:342-369).  in and out have dtype
 * dtype_in_out (F32 / BF16) and numel elements each; the result is not packed. */
PIQUANT_EXPORT void piquant_cuda_requantize(
    piquant_context_t* ctx,
    const void* in, piquant_dtype_t dtype_in_out,
    void* out, piquant_dtype_t quant_dtype,
    size_t numel,
    float scale, int64_t zero_point,
    piquant_round_mode_t mode, piquant_reduce_op_t op);

/* ---- min/max pieces of compute_quant_params, for tensors sharded over several GPUs ------------- */

/* Asynchronous: reduce x (dtype F32 or BF16, device memory) to four floats {min, max, -min, max} in the
 * DEVICE buffer out4 (16 bytes, 16-byte aligned), ordered on the context stream.  {-min, max} is the
 * layout a single MAX all-reduce combines across shards. */
PIQUANT_EXPORT void piquant_cuda_minmax_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n, float* out4);
/* The double-precision scale / zero-point arithmetic of the reference (reference src/piquant.cpp:245-258)
 * on an already reduced {min, max}. */
PIQUANT_EXPORT void piquant_cuda_params_from_minmax(float min, float max, piquant_dtype_t target_quant_dtype,
                                                    float* out_scale, int64_t* out_zero_point);

/* ---- NCCL: whole-tensor quantization parameters for a tensor sharded across ranks ------------- */

/* Fills 128 bytes with an ncclUniqueId (rank 0 calls this and broadcasts the bytes).  Returns 0 on
 * success, -1 if libnccl could not be loaded (never aborts). */
PIQUANT_EXPORT int  piquant_cuda_nccl_unique_id(void* out128);
/* Join a communicator of nranks processes (one GPU each, the current CUDA device).  After this,
 * piquant_compute_quant_params_* on this context reduces the local shard and then combines
 * {-min, max} with ONE ncclAllReduce(count=2, float, max) over NVLink, so every rank returns the
 * parameters of the whole tensor. */
PIQUANT_EXPORT void piquant_cuda_comm_init_rank(piquant_context_t* ctx, const void* unique_id128, int nranks, int rank);
/* Leave the communicator; compute_quant_params is local again. */
PIQUANT_EXPORT void piquant_cuda_comm_destroy(piquant_context_t* ctx);
/* How the ranks combine {-min, max}.  comm_init_rank also has every rank map a 512-byte mailbox of every other rank
 * (CUDA IPC over NVLink / NVSwitch); when that succeeds on all ranks the exchange happens INSIDE the min/max kernel
 * (8-byte peer stores + a sequence tag, polled by the kernel's last CTA): one launch per compute_quant_params, no
 * NCCL call.  transport: 0 = peer memory when available (default), 1 = ncclAllReduce, 2 = peer memory (abort if it
 * is not available).  All ranks must choose the same.  Collectives of one context must be issued in the same order
 * on every rank (as with any communicator). */
PIQUANT_EXPORT void piquant_cuda_comm_set_transport(piquant_context_t* ctx, int transport);
/* The transport in use: 0 = no communicator, 1 = ncclAllReduce, 2 = peer memory. */
PIQUANT_EXPORT int  piquant_cuda_comm_transport(piquant_context_t* ctx);

/* ---- device-resident parameters: no host round trip between min/max, (scale, zero_point) and quantize ----
 * The reference's callers do  params = compute_quant_params(x); q = quantize(x, params)  with a join after
 * each step (reference python/benchmark/throughput_avg.py:18-21).  Here the parameters can stay on the GPU:
 * a one-thread kernel evaluates the reference's double-precision formula (bit-identical to the host version)
 * into a 64-byte block that the quantize / dequantize kernels read, so the three launches queue back to back
 * and a tensor smaller than the 126 MB L2 is read from HBM only once. */

/* 64 bytes of DEVICE memory (8-byte aligned).  The first 16 bytes are the public result; the rest is what the
 * kernels need (1/scale, bias, range flags) and is opaque.  error != 0: the reference would have aborted
 * ("scale must be positive"). */
typedef struct piquant_cuda_meta_t {
    float   scale;
    int32_t error;
    int64_t zero_point;
    unsigned char opaque[48];
} piquant_cuda_meta_t;

/* Asynchronous: d_meta <- quantization parameters of x (F32 / BF16 device tensor, n elements; whole-tensor
 * parameters if the context has a communicator) for the quantized dtype target_quant_dtype. */
PIQUANT_EXPORT void piquant_cuda_compute_meta_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                    piquant_dtype_t target_quant_dtype, piquant_cuda_meta_t* d_meta);
/* Asynchronous: like piquant_quantize / piquant_dequantize on device pointers, with (scale, zero_point) read
 * from d_meta by the kernel instead of passed by value. */
PIQUANT_EXPORT void piquant_cuda_quantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                     piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                                     const piquant_cuda_meta_t* d_meta);
PIQUANT_EXPORT void piquant_cuda_dequantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                       piquant_dtype_t dtype_out, size_t numel, piquant_reduce_op_t op,
                                                       const piquant_cuda_meta_t* d_meta);
/* One call = compute_quant_params + quantize: min/max kernel, parameter kernel, quantize kernel, ONE stream
 * synchronisation; returns the parameters it used.  Results are bit-identical to the two separate calls. */
PIQUANT_EXPORT void piquant_cuda_quantize_auto(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                               piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                               float* out_scale, int64_t* out_zero_point);

/* ---- explicit device and stream per call --------------------------------------------------------------------
 * piquant_cuda_set_stream is context state: two threads sharing a context would have to serialise "bind, then call".
 * Every function below has the meaning of the function it is named after, with two trailing arguments instead of that
 * state, so that no call mutates anything another thread depends on (the reference tolerates concurrent calls on one
 * context, reference src/piquant.cpp:194-211 -- so does this library: scratch memory is per stream):
 *   device >= 0 : the caller vouches that every pointer of the call is device-accessible memory of that CUDA device
 *                 (what a torch CUDA tensor knows about itself): no pointer classification, no driver query per call;
 *   device == PIQUANT_CUDA_DEVICE_AUTO : find out from the pointers like the piquant.h functions do (host pointers are
 *                 accepted wherever they are there);
 *   stream      : the cudaStream_t the work is ordered on (NULL = legacy default stream). */
#define PIQUANT_CUDA_DEVICE_AUTO (-1)
/* flags */
#define PIQUANT_CUDA_FLAG_LOCAL      1u   /* min/max of THIS tensor only, even if the context has a communicator */
#define PIQUANT_CUDA_FLAG_KEEP_IN_L2 2u   /* min/max: read with L2::evict_last, a pass over the same tensor follows */
#define PIQUANT_CUDA_FLAG_REVERSE    4u   /* quantize: process the tensor from its end (what the previous pass left in L2) */

PIQUANT_EXPORT void piquant_cuda_quantize_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                    piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                                    piquant_round_mode_t mode, int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_dequantize_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                      piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                                      piquant_reduce_op_t op, int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_requantize_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in_out, void* out,
                                                      piquant_dtype_t quant_dtype, size_t numel, float scale, int64_t zero_point,
                                                      piquant_round_mode_t mode, piquant_reduce_op_t op, int device, void* stream);
/* piquant_compute_quant_params_float32 / _bfloat16 (dtype says which); synchronises `stream`. */
PIQUANT_EXPORT void piquant_cuda_compute_quant_params_on_stream(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                                piquant_dtype_t target_quant_dtype, float* out_scale,
                                                                int64_t* out_zero_point, int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_minmax_on_stream(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n, float* out4,
                                                  unsigned flags, int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_compute_meta_on_stream(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                        piquant_dtype_t target_quant_dtype, piquant_cuda_meta_t* d_meta, unsigned flags,
                                                        int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_quantize_meta_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                         piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                                         const piquant_cuda_meta_t* d_meta, unsigned flags, int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_dequantize_meta_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                           piquant_dtype_t dtype_out, size_t numel, piquant_reduce_op_t op,
                                                           const piquant_cuda_meta_t* d_meta, int device, void* stream);
PIQUANT_EXPORT void piquant_cuda_quantize_auto_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                         piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                                         float* out_scale, int64_t* out_zero_point, int device, void* stream);

/* ---- the two fused passes of a quantized ring reduction (the caller the ADD store op exists for, reference README.md:29) ----
 *
 * Reduce-scatter hop, receiver side:  out += dequantize(in; d_meta)   (piquant_dequantize with PIQUANT_REDUCE_OP_ADD)
 * and, in the same launch, min/max of the sums just written -> the reference's (scale, zero_point) arithmetic for
 * next_quant_dtype -> d_meta_next (and a second copy at d_meta_next_copy unless NULL, e.g. the header of the next
 * receiver's slot in NVLink peer memory).  The chunk a rank accumulates at hop s is the chunk it quantizes and sends at
 * hop s + 1, so the separate min/max pass over it -- a full HBM read -- disappears; results are bit-identical to
 * dequantize-ADD followed by piquant_cuda_compute_meta (local).  Never uses the communicator. */
PIQUANT_EXPORT void piquant_cuda_dequantize_add_minmax_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                                 piquant_dtype_t dtype_out, size_t numel, const piquant_cuda_meta_t* d_meta,
                                                                 piquant_dtype_t next_quant_dtype, piquant_cuda_meta_t* d_meta_next,
                                                                 piquant_cuda_meta_t* d_meta_next_copy, int device, void* stream);
/* All-gather hop:  out = dequantize(in; d_meta)   (PIQUANT_REDUCE_OP_SET)  and, in the same launch, the packed bytes of
 * `in` are stored unchanged to forward_to and the 64-byte block d_meta to forward_meta_to (NULL: not forwarded) --
 * typically the next rank's receive slot in peer memory: the hop's dequantize IS its send. */
PIQUANT_EXPORT void piquant_cuda_dequantize_forward_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                              piquant_dtype_t dtype_out, size_t numel, const piquant_cuda_meta_t* d_meta,
                                                              void* forward_to, piquant_cuda_meta_t* forward_meta_to, int device, void* stream);

/* Reduce step of an all-to-all ("direct") quantized all-reduce on an NVSwitch box: every rank has received one packed chunk
 * from every other rank.  out += dequantize(ins[0]; d_metas[0]) + ... + dequantize(ins[count-1]; d_metas[count-1]), the
 * sources folded IN ORDER with exactly the arithmetic of `count` successive piquant_dequantize(PIQUANT_REDUCE_OP_ADD) calls
 * (bit-identical to them: a bf16 accumulator is rounded after every source), in ONE pass over `out` -- 2*sizeof(out
 * type) + count*bits/8 bytes per element instead of count * (2*sizeof + bits/8) -- and, like
 * piquant_cuda_dequantize_add_minmax_on_stream, min/max of the sums -> parameters for next_quant_dtype -> d_meta_next
 * (+ d_meta_next_copy unless NULL).  1 <= count <= PIQUANT_CUDA_MAX_SUM_SOURCES; every source holds `numel` elements of
 * dtype_in.  ins / d_metas are HOST arrays of device pointers (read before the call returns). */
#define PIQUANT_CUDA_MAX_SUM_SOURCES 8
PIQUANT_EXPORT void piquant_cuda_dequantize_sum_minmax_on_stream(piquant_context_t* ctx, const void* const* ins,
                                                                 const piquant_cuda_meta_t* const* d_metas, size_t count,
                                                                 piquant_dtype_t dtype_in, void* out, piquant_dtype_t dtype_out, size_t numel,
                                                                 piquant_dtype_t next_quant_dtype, piquant_cuda_meta_t* d_meta_next,
                                                                 piquant_cuda_meta_t* d_meta_next_copy, int device, void* stream);

/* Stream-ordered copy of nbytes between any two device-accessible buffers (local, peer-mapped, pinned host) by a COPY
 * ENGINE (cudaMemcpyAsync): the SM-free way to move a packed payload into a neighbour's slot.  A kernel that stores into
 * peer memory is link-bound (NVLink sustains ~0.6 TB/s against ~7 TB/s of HBM) and holds every SM it occupies for the
 * whole transfer; a DMA transfer leaves the SMs to the kernels of other streams. */
PIQUANT_EXPORT void piquant_cuda_copy_on_stream(piquant_context_t* ctx, void* dst, const void* src, size_t nbytes, int device, void* stream);

/* Stream-ordered wait for an arrival flag: a 4-byte word in device memory that a copy engine -- of this GPU or of a peer,
 * directly or through the NVSwitch multicast address of a symmetric buffer -- sets to a non-zero value AFTER the payload it
 * announces (two piquant_cuda_copy_on_stream on one stream).  A one-thread kernel polls the word, lowers it to 0 again and ends;
 * work queued behind it on `stream` sees the payload.  Nothing here blocks the host. */
PIQUANT_EXPORT void piquant_cuda_wait_flag_on_stream(piquant_context_t* ctx, void* flag, int device, void* stream);

/* ---- many small tensors, one launch -----------------------------------------------------------------------------
 * The reference's own Python benchmark quantizes a 1e6-element tensor 1000 times (reference python/benchmark/benchmark.py:16-23);
 * on a GPU a launch costs more than 1e6 elements of data.  One call quantizes `count` independent tensors -- each with
 * its own pointers, length, scale and zero point, all f32 or all bf16 into one quantized dtype -- in ONE kernel launch
 * per 256 tensors; every tensor gets exactly the bytes a piquant_quantize call on it would produce.  Device memory
 * only; mode: PIQUANT_NEAREST or PIQUANT_STOCHASTIC (one threshold for the whole batch). */
typedef struct piquant_cuda_batch_item_t {
    const void* in;
    void*       out;
    size_t      numel;
    float       scale;
    int64_t     zero_point;
} piquant_cuda_batch_item_t;
PIQUANT_EXPORT void piquant_cuda_quantize_batch(piquant_context_t* ctx, const piquant_cuda_batch_item_t* items, size_t count,
                                                piquant_dtype_t dtype_in, piquant_dtype_t dtype_out, piquant_round_mode_t mode,
                                                int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIQUANT_CUDA_H */
