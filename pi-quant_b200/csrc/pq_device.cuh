// pq_device.cuh -- device-side building blocks of the sm_100a pi-quant kernels.
//
// Everything here is per-element arithmetic or register-level packing; the streaming structure
// (who loads what, how many bytes are in flight) lives in the kernels that include this file.
//
// Arithmetic contract (citations are file:line under the reference tree, /root/reference):
//   * nearest quantize  == one lane of the reference's widest SIMD body
//       (src/kernels/kernels_specialized.inl:57-82 for f32->u8; :344-361 u4; :686-709 u2):
//       p = x*inv (RN, no fma), a = p +- 0.5 (RN), t = x86 cvtt(a) (NaN / |a| >= 2^31 -> INT32_MIN),
//       q = t + zp (32-bit wrap), clamp to [0, qmax].
//   * f32->u2 nearest   == the generic scalar step, std::round + int64 (src/kernels/quantize.inl:21-26).
//   * stochastic        == src/kernels/quantize.inl:8-19 with ONE threshold xi per call
//       (src/piquant.cpp:199-201).
//   * dequantize        == src/kernels/kernels_specialized.inl:729-1416 bodies, generic
//       src/kernels/dequantize.inl:8-11 for u2->f32.
// No fast-math, no FTZ, no implicit contraction: every rounding is spelled with an _rn intrinsic.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pq {

// same numeric values as include/piquant.h (reference include/piquant.h:33-40); 5..7 are this library's signed extension
// (include/piquant_cuda.h).  The kernels only ever see the unsigned types: intN is the offset-binary view of uintN,
//     quantize_intN(x; scale, zp)   = quantize_uintN(x; scale, zp + 2^(N-1))  XOR  the sign bit of every field
//     dequantize_intN(q; scale, zp) = dequantize_uintN(q XOR sign bits; scale, zp + 2^(N-1))
// so every rounding and corner case of the unsigned kernels (and their parity with the reference) carries over.
enum : int { DT_F32 = 0, DT_BF16 = 1, DT_U2 = 2, DT_U4 = 3, DT_U8 = 4, DT_I2 = 5, DT_I4 = 6, DT_I8 = 7 };
// which per-element formula a quantize cell uses
enum : int { STEP_BODY = 0, STEP_ROUND64 = 1, STEP_STOCH = 2, STEP_SRPE = 3 };   // SRPE: per-element stochastic rounding (extension)
enum : int { OP_SET = 0, OP_ADD = 1 };

struct QuantParams {
    float   inv_scale;   // 1.0f / scale, IEEE divide done once on the host (kernels_specialized.inl:42)
    float   scale;
    float   xi;          // per-call stochastic threshold
    float   bias;        // -(float)zp32 * scale, for the fma-form dequantize (kernels_specialized.inl:1204)
    int32_t zp32;        // (int32_t)zero_point, the truncation of quantize.inl:112-128
    int32_t bigzp;       // |zero_point| > 2^29: the int32 fast path of the int64 formulas is not valid
    int32_t spec_ok32;   // |zp32| <= 2^29: the speculative group path is valid for the int32 (SIMD-body) formula
    uint32_t sign_xor;   // signed dtypes: the sign bit of every packed field of a 32-bit word (0x80808080 / 0x88888888 / 0xAAAAAAAA), else 0
    int64_t zp64;
};

// Parameters produced on the device live inside a 64-byte DeviceMeta block (pq_kernels.h): {scale, error, zero_point, P}.
// `error` != 0 means the reference would have aborted in compute_quant_params (scale NaN or negative, src/piquant.cpp:373);
// a kernel handed such a block does no work at all -- the host cannot abort for it without a synchronisation -- and the
// flag stays in the block for whoever reads the parameters back (piquant_cuda_quantize_auto aborts on it like the reference).
__device__ __forceinline__ bool device_params_failed(const QuantParams* dP) {
    return *reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(dP) - 12) != 0;
}

// ------------------------------------------------------------------------------------------------
// programmatic dependent launch (see launch_kernel in pq_kernels.h)
// ------------------------------------------------------------------------------------------------

// Let the next kernel on the stream start scheduling its CTAs as SMs free up.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Block until everything launched before this kernel on the stream has completed and is visible.
// Must precede the first global-memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// global memory access: 4/8/16/32-byte vector loads and stores with streaming cache hints.
// 32-byte accesses assemble to LDG.E.256 / STG.E.256 on sm_100a: one full 32 B sector per thread.
// ------------------------------------------------------------------------------------------------

// read-only streaming load (data never written by this kernel): non-coherent path, no L1 allocation
__device__ __forceinline__ void ldg_stream(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg_stream(const void* p, uint32_t (&r)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
}
__device__ __forceinline__ void ldg_stream(const void* p, uint32_t (&r)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(p));
}
__device__ __forceinline__ void ldg_stream(const void* p, uint32_t (&r)[1]) {
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r[0]) : "l"(p));
}

// read-only load that asks L2 to KEEP the line (a following pass will read it again)
__device__ __forceinline__ void ldg_keep(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}

// coherent load of data this kernel will overwrite (the accumulator of dequantize-ADD)
__device__ __forceinline__ void ldg_rmw(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p) : "memory");
}
__device__ __forceinline__ void ldg_rmw(const void* p, uint32_t (&r)[4]) {
    asm volatile("ld.global.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p) : "memory");
}

__device__ __forceinline__ void stg_stream(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void stg_stream(void* p, const uint32_t (&r)[4]) {
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

// stores of data that a following pass will read again: ask L2 to keep the lines (SASS: STG.E.NA.ELL2)
__device__ __forceinline__ void stg_keep(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.L1::no_allocate.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// (the modifier exists for 32-byte stores only; the 16-byte-aligned variant of a kernel stores normally)

// Load NW 32-bit words from a NW*4-byte aligned address (A32: 32-byte LDG.256, else 16-byte LDG.128).
template <int NW, bool A32>
__device__ __forceinline__ void load_words(const void* p, uint32_t (&w)[NW]) {
    if constexpr (NW >= 8 && A32) {
        static_assert(NW % 8 == 0);
#pragma unroll
        for (int i = 0; i < NW / 8; ++i) {
            uint32_t t[8];
            ldg_stream(static_cast<const char*>(p) + 32 * i, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) w[8 * i + k] = t[k];
        }
    } else if constexpr (NW >= 4) {
        static_assert(NW % 4 == 0);
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            uint32_t t[4];
            ldg_stream(static_cast<const char*>(p) + 16 * i, t);
#pragma unroll
            for (int k = 0; k < 4; ++k) w[4 * i + k] = t[k];
        }
    } else if constexpr (NW == 2) {
        ldg_stream(p, w);
    } else {
        static_assert(NW == 1);
        ldg_stream(p, w);
    }
}

template <int NW, bool A32>
__device__ __forceinline__ void load_words_rmw(const void* p, uint32_t (&w)[NW]) {
    if constexpr (A32) {
        static_assert(NW % 8 == 0);
#pragma unroll
        for (int i = 0; i < NW / 8; ++i) {
            uint32_t t[8];
            ldg_rmw(static_cast<const char*>(p) + 32 * i, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) w[8 * i + k] = t[k];
        }
    } else {
        static_assert(NW % 4 == 0);
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            uint32_t t[4];
            ldg_rmw(static_cast<const char*>(p) + 16 * i, t);
#pragma unroll
            for (int k = 0; k < 4; ++k) w[4 * i + k] = t[k];
        }
    }
}

// KEEP: L2::evict_last (a following pass reads the data again) instead of the plain streaming store
template <int NW, bool A32, bool KEEP = false>
__device__ __forceinline__ void store_words(void* p, const uint32_t (&w)[NW]) {
    if constexpr (A32 && NW >= 8) {
        static_assert(NW % 8 == 0);
#pragma unroll
        for (int i = 0; i < NW / 8; ++i) {
            uint32_t t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = w[8 * i + k];
            if constexpr (KEEP) stg_keep(static_cast<char*>(p) + 32 * i, t);
            else stg_stream(static_cast<char*>(p) + 32 * i, t);
        }
    } else if constexpr (NW >= 4) {
        static_assert(NW % 4 == 0);
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            uint32_t t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = w[4 * i + k];
            stg_stream(static_cast<char*>(p) + 16 * i, t);
        }
    } else if constexpr (NW == 2) {
        asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
    } else {
        static_assert(NW == 1);
        asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// bf16 <-> f32
// ------------------------------------------------------------------------------------------------

// bfp16_t -> fp32_t is a 16-bit shift (include/piquant.hpp:95)
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }

// two f32 -> packed bf16x2, round-to-nearest-even (include/piquant.hpp:86-90,
// kernels_specialized.inl:15-32).  Identical to the reference for every non-NaN input; NaN comes
// out as the canonical quiet NaN (the reference's own two bf16 paths do not agree on NaN payloads).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float x) { return static_cast<uint16_t>(pack_bf16x2(x, 0.0f) & 0xffffu); }

// element `e` of an item whose input words are w[]: f32 = one word, bf16 = half a word
template <int IN_DT, int NW>
__device__ __forceinline__ float item_elem(const uint32_t (&w)[NW], int e) {
    if constexpr (IN_DT == DT_F32) return __uint_as_float(w[e]);
    // (Widening a half with the sm_100a mixed-precision add -- FHADD.BF16 Rd, Rw.H1, -RZ, no shift / mask -- was measured
    // too: 1-3 % slower on every bf16 cell, profiles/r1_cellbench_1e9_fhadd_variant.txt.)
    else return (e & 1) ? bf16_hi(w[e >> 1]) : bf16_lo(w[e >> 1]);
}

// ------------------------------------------------------------------------------------------------
// quantize steps
// ------------------------------------------------------------------------------------------------

// One lane of the reference's SIMD bodies.  copysign(0.5, p) equals the reference's
// (p >= 0 ? +0.5 : -0.5) after truncation for every p: they differ only at p == -0.0
// (-0.5 vs +0.5, both truncate to 0) and p == NaN (a is NaN either way).
__device__ __forceinline__ int32_t quant_step_body(float x, float inv, int32_t zp32, int32_t qmax) {
    const float p = __fmul_rn(x, inv);
    const float h = __uint_as_float(0x3f000000u | (__float_as_uint(p) & 0x80000000u));
    const float a = __fadd_rn(p, h);
    int32_t t = __float2int_rz(a);                 // saturating; x86 cvttps2dq gives INT32_MIN instead:
    if (!(a < 2147483648.0f)) t = INT32_MIN;       // NaN and a >= 2^31 ("integer indefinite"); a < -2^31 saturates to it already
    const int32_t q = static_cast<int32_t>(static_cast<uint32_t>(t) + static_cast<uint32_t>(zp32));   // vpaddd wraps
    return min(max(q, 0), qmax);
}

// static_cast<int64_t>(float) on x86-64 (cvttss2si r64): out of range and NaN give INT64_MIN
__device__ __forceinline__ long long x86_cvtt_i64(float a) {
    return (a >= -9223372036854775808.0f && a < 9223372036854775808.0f) ? __float2ll_rz(a) : LLONG_MIN;
}

// clamp(int64(rnd) + zp, 0, qmax) for an integer-valued (or NaN/inf) float `rnd`.
// Fast path in 32-bit arithmetic when it is provably exact: |rnd| < 2^30 and |zp| <= 2^29.
__device__ __forceinline__ int32_t finish_i64(float rnd, const QuantParams& P, int32_t qmax) {
    if (!P.bigzp && fabsf(rnd) < 1073741824.0f) {
        return min(max(__float2int_rz(rnd) + P.zp32, 0), qmax);
    }
    const long long t = x86_cvtt_i64(rnd);
    const long long q = static_cast<long long>(static_cast<unsigned long long>(t) + static_cast<unsigned long long>(P.zp64));
    return static_cast<int32_t>(q < 0 ? 0ll : (q > static_cast<long long>(qmax) ? static_cast<long long>(qmax) : q));
}

// quant_step_scalar_nearest (quantize.inl:21-26)
__device__ __forceinline__ int32_t quant_step_round64(float x, const QuantParams& P, int32_t qmax) {
    return finish_i64(roundf(__fmul_rn(x, P.inv_scale)), P, qmax);
}

// quant_step_scalar_stochastic (quantize.inl:8-19)
__device__ __forceinline__ int32_t quant_step_stochastic(float x, const QuantParams& P, int32_t qmax) {
    const float r = __fmul_rn(x, P.inv_scale);
    const float tr = truncf(r);
    const float dec = fabsf(__fsub_rn(r, tr));
    float adj = (P.xi < dec) ? 1.0f : 0.0f;
    if (r < 0.0f) adj = -adj;
    return finish_i64(__fadd_rn(tr, adj), P, qmax);
}

template <int STEP>
__device__ __forceinline__ int32_t quant_step(float x, const QuantParams& P, int32_t qmax) {
    if constexpr (STEP == STEP_BODY) return quant_step_body(x, P.inv_scale, P.zp32, qmax);
    else if constexpr (STEP == STEP_ROUND64) return quant_step_round64(x, P, qmax);
    else { static_assert(STEP == STEP_STOCH, "STEP_SRPE goes through quant_step_srpe"); return quant_step_stochastic(x, P, qmax); }
}

// ------------------------------------------------------------------------------------------------
// per-element stochastic rounding (extension, include/piquant_cuda.h: PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT)
//
//   q_i = clamp(floor(x_i / scale + u_i) + zp, 0, qmax),   u_i = (k_i + 1/2) * 2^-16,
//   k_i = 16 bits of Philox4x32-10(counter = {i / 8, 0, 0, 0} (64-bit i / 8 in the first two words), key = the call's 64-bit key):
//         element i takes bits [16 * (i % 2), +16) of output word (i % 8) / 2.
// E[q_i] = x_i / scale + zp inside the range -- unbiased per element, which the reference's one-threshold-per-call mode
// (src/piquant.cpp:199-201) is not.  x / scale is RN(x * (1.0f / scale)) like every other mode.
// ------------------------------------------------------------------------------------------------

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123): 10 rounds of
//   {c0, c1, c2, c3} <- {hi(M1*c2) ^ c1 ^ k0, lo(M1*c2), hi(M0*c0) ^ c3 ^ k1, lo(M0*c0)},  key += {W0, W1} per round.
// 2 IMAD.WIDE + 2 LOP3 per round; the round keys are uniform (key + r * W), so they cost nothing per element.
struct PhiloxKey { uint32_t k0, k1; };
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = static_cast<unsigned long long>(M0) * c0;
        const unsigned long long p1 = static_cast<unsigned long long>(M1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ (key.k0 + static_cast<uint32_t>(r) * W0);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ (key.k1 + static_cast<uint32_t>(r) * W1);
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 1 + u for the 16 random bits in half `hi` of `r`: the float 1.0 + (k + 1/2) * 2^-16, built in the mantissa (one PRMT / LOP3)
__device__ __forceinline__ float srpe_one_plus_u(uint32_t r, int hi) {
    const uint32_t k = hi ? (r >> 16) : (r & 0xffffu);
    return __uint_as_float(0x3f800040u | (k << 7));
}

// exact step, any input: floor(p + u) in double (p + 1 + u needs < 53 bits while |p| < 2^23; beyond that p is an integer,
// floor(p + u) == p, and NaN / inf / |p| >= 2^63 take the x86 conversion result like the other int64 formulas)
__device__ __forceinline__ int32_t quant_step_srpe(float x, const QuantParams& P, int32_t qmax, float one_plus_u) {
    const float p = __fmul_rn(x, P.inv_scale);
    long long t;
    if (fabsf(p) < 8388608.0f) t = __double2ll_rd(static_cast<double>(p) + static_cast<double>(one_plus_u)) - 1;
    else t = x86_cvtt_i64(p);
    const long long q = static_cast<long long>(static_cast<unsigned long long>(t) + static_cast<unsigned long long>(P.zp64));
    return static_cast<int32_t>(q < 0 ? 0ll : (q > static_cast<long long>(qmax) ? static_cast<long long>(qmax) : q));
}

// speculative step: floor(RM(p + (1 + u))) == floor(p + 1 + u) (an integer k <= y stays <= RM(y)); the "- 1" is folded into
// the zero point by the caller.  FMUL, FADD.RM, F2I.FLOOR: exact for |p| < 2^22.
__device__ __forceinline__ int32_t quant_spec_srpe(float x, const QuantParams& P, float one_plus_u, float& witness) {
    const float p = __fmul_rn(x, P.inv_scale);
    witness = p;
    return __float2int_rd(__fadd_rd(p, one_plus_u));
}

// the same step with the result kept as a float (requantize): floor(p + u) for |p| < 2^22, +0.0 for a zero result.
//   y = RM(p + (1 + u));  RM(y + 1.5*2^23) = 1.5*2^23 + floor(y)  (ulp 1 in [2^23, 2^24));  minus (1.5*2^23 + 1): exact.
__device__ __forceinline__ float requant_spec_srpe(float x, const QuantParams& P, float one_plus_u, float& witness) {
    const float p = __fmul_rn(x, P.inv_scale);
    witness = p;
    return __fadd_rn(__fadd_rd(__fadd_rd(p, one_plus_u), 12582912.0f), -12582913.0f);
}

// ------------------------------------------------------------------------------------------------
// quantize, group form: the hot loop.
//
// The exact steps above cost ~9-14 instructions per element (range check + select for the x86
// "integer indefinite", two clamps, shift/or packing) and the ALU pipe, not HBM, becomes the limiter
// under the power cap.  A group of elements owned by one thread is therefore quantized
// speculatively: per element FMUL, LOP3, FADD, F2I, IADD and half an I2IP (cvt.pack.sat: clamp to
// [0, 2^b-1] AND pack two elements in one instruction), plus half an FMNMX3.NAN that folds |a| of
// the group into one range witness.  If the witness shows every |a| < 2^30 (and |zp| <= 2^29, a
// host-checked flag) no conversion saturated and no add wrapped, so the speculative result IS the
// reference's; otherwise (NaN, inf, huge values, extreme zero points) the whole group is redone
// with the exact steps.  Results are bit-identical either way.
// ------------------------------------------------------------------------------------------------

// d = (c << 2*BITS) | (sat_uBITS(hi) << BITS) | sat_uBITS(lo)      (SASS: I2IP.U8/U4/U2.S32.SAT)
template <int BITS>
__device__ __forceinline__ uint32_t pack_sat2(int32_t hi, int32_t lo, uint32_t c) {
    uint32_t d;
    if constexpr (BITS == 8) asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
    else if constexpr (BITS == 4) asm("cvt.pack.sat.u4.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
    else asm("cvt.pack.sat.u2.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
    return d;
}

// max(|a|, |b|, |c|), NaN-propagating (SASS: FMNMX3.NAN with |.| operand modifiers)
__device__ __forceinline__ float max3_abs_nan(float a, float b, float c) {
    float d;
    asm("max.NaN.abs.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Speculative step: the integer BEFORE zero point and clamp, plus the float whose magnitude decides
// whether the speculation was valid (quant_spec_limit).
//
// STEP_ROUND64, std::round(p) == sign(p) * floor(|p| + 0.5):
//     z = p + copysign(0.5, p) rounded TOWARD ZERO, t = trunc(z).  floor(RZ(y)) == floor(y) for every y >= 0 (an integer
//     k <= y is representable, so k <= RZ(y) <= y), hence exact for every |p| < 2^31 -- unlike the SIMD body's
//     round-to-nearest add, which is what makes pred(0.5) round up there.  FMUL, LOP3, FADD.RZ, F2I.
// STEP_STOCH, trunc(r) + sign(r) * [xi < |r - trunc(r)|] == sign(r) * ceil(|r| - xi) for 0 <= xi < 1:
//     with m = |r|: frac(m) > xi  <=>  m - xi > trunc(m)  <=>  ceil(m - xi) == trunc(m) + 1, and frac(m) <= xi gives
//     trunc(m) - 1 < m - xi <= trunc(m).  The subtraction rounds TOWARD +INF and ceil(RP(y)) == ceil(y) by the same
//     argument, so one FADD.RP and one F2I.CEIL replace the reference's trunc / subtract / compare / add chain; the
//     sign goes back on with an integer multiply by +-1 that ptxas fuses with the zero-point add (IMAD, FMA pipe).
//     FMUL, FADD.RP, F2I.CEIL, SHF, LOP3, IMAD: 6 instructions where the literal formula took 11.
// Both were checked against the literal formulas on 4e7 adversarial values (ties, neighbours of ties, every binade,
// raw bit patterns) before they went in; tests/test_gpu_parity.py pins them on the GPU.
template <int STEP>
__device__ __forceinline__ int32_t quant_spec(float x, const QuantParams& P, float& witness) {
    const float p = __fmul_rn(x, P.inv_scale);
    const float h = __uint_as_float(0x3f000000u | (__float_as_uint(p) & 0x80000000u));
    if constexpr (STEP == STEP_BODY) {
        const float a = __fadd_rn(p, h);
        witness = a;
        return __float2int_rz(a);
    } else if constexpr (STEP == STEP_ROUND64) {
        witness = p;
        return __float2int_rz(__fadd_rz(p, h));
    } else {
        witness = p;
        const int32_t mag = __float2int_ru(__fadd_ru(fabsf(p), -P.xi));
        const int32_t sgn = (__float_as_int(p) >> 31) | 1;
        return mag * sgn;
    }
}

// largest |witness| for which quant_spec<STEP> followed by a 32-bit zero-point add is exact
template <int STEP>
__device__ __forceinline__ constexpr float quant_spec_limit() { return STEP == STEP_SRPE ? 4194304.0f : 1073741824.0f; }

// STEP_STOCH speculation also needs the threshold where the ceil identity holds (a context never passes anything else)
template <int STEP>
__device__ __forceinline__ bool quant_spec_params_ok(const QuantParams& P) {
    if constexpr (STEP == STEP_BODY) return P.spec_ok32 != 0;
    else if constexpr (STEP == STEP_ROUND64 || STEP == STEP_SRPE) return P.bigzp == 0;
    else return P.bigzp == 0 && P.xi >= 0.0f && P.xi < 1.0f;
}

// The same two steps with the result kept as a FLOAT (requantize never needs the integer): the rounded value of
// p = x/scale, exact for |p| < 2^22, with +0.0 for a zero result like the reference's float(q - zp).
//   ROUND64: f = RZ(RZ(p + h) + M) - M with M = copysign(2^23, p) = h * 2^24: the add of M drops the fraction.
//   STOCH:   f = RP(RP(|p| - xi) + 1.5*2^23) is 1.5*2^23 + ceil(|p| - xi) (the argument may be in (-1, 0]), then
//            (f - 1.5*2^23) * sign as one exact FFMA with s = +-1.
// No conversion instruction (XU pipe) at all.
template <int STEP>
__device__ __forceinline__ float requant_spec(float x, const QuantParams& P, float& witness) {
    static_assert(STEP == STEP_ROUND64 || STEP == STEP_STOCH);
    const float p = __fmul_rn(x, P.inv_scale);
    witness = p;
    if constexpr (STEP == STEP_ROUND64) {
        const float h = __uint_as_float(0x3f000000u | (__float_as_uint(p) & 0x80000000u));
        const float M = __fmul_rn(h, 16777216.0f);
        return __fadd_rn(__fadd_rz(__fadd_rz(p, h), M), -M);
    } else {
        const float s = __uint_as_float(0x3f800000u | (__float_as_uint(p) & 0x80000000u));
        const float f = __fadd_ru(__fadd_ru(fabsf(p), -P.xi), 12582912.0f);
        return __fmaf_rn(f, s, __fmul_rn(s, -12582912.0f));
    }
}

// Quantize the NE elements held in w[] (f32: one per word, bf16: two per word) and pack them,
// element 0 in the lowest bits, into o[NE*BITS/32 words] (at least one word; unused high bits are 0).
// STEP_SRPE: rnd[] holds the 16 random bits of element e in half (e & 1) of word e / 2 (unused otherwise).
template <int IN_DT, int BITS, int STEP, int NW>
__device__ __forceinline__ void quant_group(const uint32_t (&w)[NW], const QuantParams& P,
                                            uint32_t (&o)[(NW * (IN_DT == DT_F32 ? 1 : 2) * BITS + 31) / 32],
                                            const uint32_t (&rnd)[NW * (IN_DT == DT_F32 ? 1 : 2) / 2]) {
    constexpr int NE = NW * (IN_DT == DT_F32 ? 1 : 2);
    constexpr int OW = (NE * BITS + 31) / 32;
    constexpr int EPW = 32 / BITS;                      // elements per output word
    constexpr int QMAX = (1 << BITS) - 1;
    static_assert(NE % 2 == 0, "groups hold an even number of elements");
    int32_t t[NE];
    float wit[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if constexpr (STEP == STEP_SRPE) t[e] = quant_spec_srpe(item_elem<IN_DT, NW>(w, e), P, srpe_one_plus_u(rnd[e >> 1], e & 1), wit[e]);
        else t[e] = quant_spec<STEP>(item_elem<IN_DT, NW>(w, e), P, wit[e]);
    }
    float m = 0.0f;
#pragma unroll
    for (int e = 0; e < NE; e += 2) m = max3_abs_nan(m, wit[e], wit[e + 1]);
    const int32_t zp_add = STEP == STEP_SRPE ? P.zp32 - 1 : P.zp32;      // SRPE: the speculative step is floor(p + 1 + u)
    if (quant_spec_params_ok<STEP>(P) && m < quant_spec_limit<STEP>()) {
#pragma unroll
        for (int j = 0; j < OW; ++j) {
            uint32_t d = 0;
            constexpr int LAST = EPW < NE ? EPW : NE;
#pragma unroll
            for (int k = LAST - 2; k >= 0; k -= 2)
                d = pack_sat2<BITS>(t[j * EPW + k + 1] + zp_add, t[j * EPW + k] + zp_add, d);
            o[j] = d;
        }
    } else {
#pragma unroll
        for (int j = 0; j < OW; ++j) o[j] = 0u;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            uint32_t q;
            if constexpr (STEP == STEP_SRPE) q = static_cast<uint32_t>(quant_step_srpe(item_elem<IN_DT, NW>(w, e), P, QMAX, srpe_one_plus_u(rnd[e >> 1], e & 1)));
            else q = static_cast<uint32_t>(quant_step<STEP>(item_elem<IN_DT, NW>(w, e), P, QMAX));
            o[(e * BITS) / 32] |= q << ((e * BITS) % 32);
        }
    }
    // signed dtypes: offset binary -> two's complement, one LOP3 per packed word (fields that do not exist stay 0)
    constexpr uint32_t LAST_VALID = (NE * BITS) % 32 == 0 ? 0xffffffffu : ((1u << ((NE * BITS) % 32)) - 1u);
#pragma unroll
    for (int j = 0; j < OW; ++j) o[j] ^= P.sign_xor & (j == OW - 1 ? LAST_VALID : 0xffffffffu);
}

template <int IN_DT, int BITS, int STEP, int NW>
__device__ __forceinline__ void quant_group(const uint32_t (&w)[NW], const QuantParams& P,
                                            uint32_t (&o)[(NW * (IN_DT == DT_F32 ? 1 : 2) * BITS + 31) / 32]) {
    static_assert(STEP != STEP_SRPE, "per-element stochastic rounding needs the random words");
    const uint32_t none[NW * (IN_DT == DT_F32 ? 1 : 2) / 2] = {};
    quant_group<IN_DT, BITS, STEP, NW>(w, P, o, none);
}

// ------------------------------------------------------------------------------------------------
// dequantize steps
// ------------------------------------------------------------------------------------------------

// value of one dequantized element before the store op; HAS_BODY selects the SIMD-body formula
// family of the reference, the other branch is the generic dequant_step (u2->f32 only).
//   u8/u4 -> f32, u8 -> bf16 : float(int32(q) - zp32) * scale            (kernels_specialized.inl:748-761, :1030-1052, :946-966)
//   u4/u2 -> bf16            : fma(float(q), scale, -float(zp32)*scale)   (:1204,:1236-1243, :1318,:1361)
//   u2    -> f32             : float(int64(q) - zp64) * scale             (dequantize.inl:8-11)
template <int BITS, int OUT_DT>
__device__ __forceinline__ float dequant_term(uint32_t q, const QuantParams& P) {
    if constexpr (BITS == 2 && OUT_DT == DT_F32) {
        return P.bigzp ? __ll2float_rn(static_cast<long long>(static_cast<unsigned long long>(q) - static_cast<unsigned long long>(P.zp64)))
                       : static_cast<float>(static_cast<int32_t>(q) - P.zp32);
    } else {
        return static_cast<float>(static_cast<int32_t>(q - static_cast<uint32_t>(P.zp32)));
    }
}

// Final f32 value for out dtype f32 (`prev` used only for ADD).  GCC contracts the reference's
// mul + add(o) into one fma in its FMA-enabled translation units; we state that fma explicitly.
template <int BITS, int OP>
__device__ __forceinline__ float dequant_f32(uint32_t q, float prev, const QuantParams& P) {
    const float d = dequant_term<BITS, DT_F32>(q, P);
    if constexpr (OP == OP_ADD) return __fmaf_rn(d, P.scale, prev);
    else return __fmul_rn(d, P.scale);
}

// Final f32 value that is then rounded to bf16.
template <int BITS, int OP>
__device__ __forceinline__ float dequant_bf16_pre(uint32_t q, float prev, const QuantParams& P) {
    if constexpr (BITS == 8) {
        const float d = dequant_term<8, DT_BF16>(q, P);
        if constexpr (OP == OP_ADD) return __fmaf_rn(d, P.scale, prev);
        else return __fmul_rn(d, P.scale);
    } else {
        const float f = __fmaf_rn(static_cast<float>(q), P.scale, P.bias);
        if constexpr (OP == OP_ADD) return __fadd_rn(f, prev);
        else return f;
    }
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
// two bf16 lanes per instruction (SASS: HMNMX2.BF16); NaN never wins, like the f32 fminf / fmaxf
__device__ __forceinline__ uint32_t min_bf16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace pq
