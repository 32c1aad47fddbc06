/*
 * piquant_oracle.h -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C11 scalar restatement of the arithmetic of PrimeIntellect-ai/pi-quant's hot path
 * (quantize / dequantize / requantize / min-max -> quant params).  Nothing under pi-quant_b200/
 * may include, link or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it, and only as the checker.
 *
 * PARITY PINNED: this oracle is checked bit-for-bit against the unmodified reference compiled
 * from /root/reference into oracle/_ref/libpiquant_ref.so (see oracle/Makefile and
 * tests/test_oracle_vs_reference.py), against the committed golden vectors generated from that
 * library (tests/golden/, made by tests/golden/make_golden.py) and against the reference's own
 * known-answer tests (test/quant.cpp:198-217, test/naive.hpp:52-96, python/tests/test_torch.py).
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef PIQUANT_ORACLE_H
#define PIQUANT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as include/piquant.h:23-40 of the reference. */
enum { ORC_NEAREST = 0, ORC_STOCHASTIC = 1 };
enum { ORC_SET = 0, ORC_ADD = 1 };
enum { ORC_F32 = 0, ORC_BF16 = 1, ORC_UINT2 = 2, ORC_UINT4 = 3, ORC_UINT8 = 4 };
/*
 * EXTENSION, PARITY UNPINNED (there is nothing to pin it to): signed quantized dtypes.  The reference at this commit has
 * none (include/piquant.h:33-40) -- only the is_signed branches of compute_type_max / compute_quant_config
 * (src/piquant.cpp:212-220,246-248) hint at them.  The library under test (include/piquant_cuda.h) defines intN as the
 * offset-binary view of uintN, and so does this oracle, by calling its own pinned unsigned functions:
 *     quantize(x -> intN; scale, zp)   = quantize(x -> uintN; scale, zp + 2^(N-1)), sign bit of every existing field flipped
 *     dequantize(q: intN; scale, zp)   = dequantize(q with sign bits flipped: uintN; scale, zp + 2^(N-1))
 *     requantize(via intN; scale, zp)  = requantize(via uintN; scale, zp + 2^(N-1))
 * tests/test_oracle_signed.py checks that on ordinary inputs this equals the textbook
 * clamp(round_half_away(x / scale) + zp, -2^(N-1), 2^(N-1)-1) in two's complement.
 */
enum { ORC_INT2 = 5, ORC_INT4 = 6, ORC_INT8 = 7 };

/*
 * Which of the reference's (not fully self-consistent) per-element formulas to apply.
 *
 * ORC_SEM_BODY   every element goes through the formula of the reference's widest SIMD body
 *                (kernels_specialized.inl AVX-512 lanes) when the cell has one, otherwise through
 *                the generic scalar step (quantize.inl / dequantize.inl).  This is the semantics
 *                the CUDA library implements: it does not depend on thread count or alignment.
 * ORC_SEM_REF    emulate the AVX-512 build of the reference exactly: the per-thread partition of
 *                piquant.cpp:132-176 and, inside each partition, the scalar head / SIMD body /
 *                scalar tail split of each kernel (which uses std::round instead of trunc(p+-0.5)
 *                outside the body).  Needs `nthreads` (the context's thread count).
 *
 * The two differ only where trunc(p +- 0.5) != round(p), i.e. |p| == 0x1.fffffep-2f (and odd
 * |p| in [2^23, 2^24) before clamping), and in the rounding of float ADD accumulation.
 */
enum { ORC_SEM_BODY = 0, ORC_SEM_REF = 1 };

/* packed byte count for `numel` elements of a quantized dtype (piquant_internal.hpp:41-44) */
size_t orc_packed_bytes(int dtype, size_t numel);
/* storage bytes for `numel` elements of any dtype */
size_t orc_storage_bytes(int dtype, size_t numel);

/* context::quantize (piquant.cpp:277-308) -> quant_generic (quantize.inl:101-149).
 * rnd_threshold is the per-call xi of piquant.cpp:199-201 (ignored for ORC_NEAREST).
 * Returns 0, or -1 for an invalid dtype combination (where the reference aborts). */
int orc_quantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel,
                 float scale, int64_t zero_point, int round_mode, float rnd_threshold,
                 int semantics, int nthreads);

/*
 * EXTENSION, PARITY UNPINNED: per-element stochastic rounding (include/piquant_cuda.h,
 * PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT).  The reference only has the one-threshold-per-call mode above.  Spec:
 *     q_i = clamp(floor(p_i + u_i) + zero_point, qmin, qmax),  p_i = RN(x_i * (1.0f / scale)),  u_i = (k_i + 1/2) * 2^-16,
 *     k_i = 16 bits of Philox4x32-10(counter = {lo32(j / 8), hi32(j / 8), 0, 0}, key = {lo32(key), hi32(key)}), j = base + i:
 *           bits [16 * (j % 2), +16) of output word (j % 8) / 2.
 * |p_i| >= 2^23 (already an integer), NaN and inf take (int64)p_i with the x86 conversion rule, like the other int64 formulas.
 * The Philox4x32-10 core is pinned by the Random123 known-answer vectors (tests/test_oracle_sr.py).
 */
int orc_quantize_sr(const void* in, int dt_in, void* out, int dt_out, int64_t numel,
                    float scale, int64_t zero_point, uint64_t key, int64_t base);
/* fused quantize -> dequantize with the same per-element rule (the library's piquant_cuda_requantize with mode 2) */
int orc_requantize_sr(const void* in, int dt_inout, void* out, int dt_quant, int64_t numel,
                      float scale, int64_t zero_point, uint64_t key, int64_t base, int reduce_op, int fma_add);
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* context::dequantize (piquant.cpp:310-340) -> dequant_generic (dequantize.inl:89-140). */
int orc_dequantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel,
                   float scale, int64_t zero_point, int reduce_op,
                   int semantics, int nthreads);

/* context::quantize_dequantize_fused (piquant.cpp:342-369) -> requant_generic (kernels.inl:30-52).
 * `fma_add` selects how `o[i] += v*scale` is rounded for f32 ADD: 1 = contracted to one fma (what
 * GCC emits for the FMA-enabled translation units), 0 = separate multiply and add. */
int orc_requantize(const void* in, int dt_inout, void* out, int dt_quant, int64_t numel,
                   float scale, int64_t zero_point, int round_mode, float rnd_threshold,
                   int reduce_op, int fma_add);

/* find_min_max_f32 / find_min_max_bf16 (kernels_specialized.inl:1418-1516 / 1518-1607).
 * out[0] = min, out[1] = max, both starting from +-FLT_MAX; NaNs never win a comparison. */
void orc_minmax_f32(const float* x, int64_t n, float out[2]);
void orc_minmax_bf16(const uint16_t* x, int64_t n, float out[2]);

/* The double-precision scale / zero-point arithmetic of compute_quant_config
 * (piquant.cpp:245-258) on an already reduced {min,max}.  Returns -1 where the reference would
 * abort (piquant.cpp:373: scale NaN or negative). */
int orc_params_from_minmax(double r_min, double r_max, int dt_quant, float* scale, int64_t* zero_point);

/* context::compute_quant_config_from_data (piquant.cpp:371-381). n == 0 returns -1 (abort). */
int orc_compute_quant_params_f32(const float* x, int64_t n, int dt_quant, float* scale, int64_t* zero_point);
int orc_compute_quant_params_bf16(const uint16_t* x, int64_t n, int dt_quant, float* scale, int64_t* zero_point);

/* GCC contracts `mul + add(o)` into one fma in every FMA-enabled translation unit of the reference
 * (-ffp-contract=fast is the GNU-mode default).  1 (default) reproduces that build for the f32
 * accumulate of dequantize-ADD, 0 gives the source-level two-rounding result. */
void orc_set_fma_contract(int on);

/* bfp16_t conversions (piquant.hpp:86-90, :95) */
uint16_t orc_f32_to_bf16(float x);
float    orc_bf16_to_f32(uint16_t b);

/* single-element steps, exported so tests can probe special values */
int32_t orc_quant_step_body(float x, float inv_scale, int32_t zp32, int32_t qmax);
int64_t orc_quant_step_scalar_nearest(float x, float inv_scale, int64_t zp, int64_t qmax);
int64_t orc_quant_step_scalar_stochastic(float x, float inv_scale, int64_t zp, int64_t qmax, float xi);
int32_t orc_quant_step_tail32(float x, float inv_scale, int32_t zp32, int32_t qmax);

#ifdef __cplusplus
}
#endif
#endif
