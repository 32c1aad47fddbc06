// pq_reduce.cuh -- the tail every min/max-producing kernel ends with (sm_100a):
//
//   per-thread {min, max}  ->  warp shuffles  ->  one partial per CTA  ->  the LAST CTA to finish (atomic ticket)
//   folds the partials  ->  [sharded tensor: exchange {-min, max} with the other ranks through NVLink peer memory]
//   ->  {min, max, -min, max} to device (+ pinned-mapped host) memory
//   ->  [optionally the reference's (scale, zero_point) arithmetic, as a 64-byte parameter block]
//
// so that "reduce a tensor and know how to quantize it" is ONE launch: no second kernel for the fold, no one-thread
// parameter kernel, and for a tensor sharded over the GPUs of a box no NCCL call either.  Used by minmax_kernel
// (minmax.cu) and by the fused dequantize-ADD + min/max kernel of the ring reduction (dequantize.cu).
//
// Replaces, together, the partial gather of compute_quant_config (reference src/piquant.cpp:222-244) and its
// scale / zero-point arithmetic (reference src/piquant.cpp:245-258).
#pragma once

#include <cfloat>
#include <climits>

#include "pq_kernels.h"

namespace pq {

__device__ __forceinline__ long long dev_cvttsd_i64(double a) {     // x86 cvttsd2si: out of range / NaN -> INT64_MIN
    return (a >= -9223372036854775808.0 && a < 9223372036854775808.0) ? __double2ll_rz(a) : LLONG_MIN;
}

// compute_quant_config after the gather (reference src/piquant.cpp:245-258): same expressions, same order, same IEEE
// double operations as params_from_minmax() in context.cu -- bit-identical to the host version (tests sweep ranges x dtypes)
// -- plus everything the kernels derive from (scale, zero_point): 1/scale, bias, range flags (make_params() in context.cu).
__device__ __forceinline__ DeviceMeta meta_from_minmax(float neg_min, float max, int bits, int is_signed, uint32_t sign_xor) {
    const double r_min = -static_cast<double>(neg_min), r_max = static_cast<double>(max);
    const unsigned long long type_max = (1ull << (bits - (is_signed ? 1 : 0))) - 1;
    const long long type_min = is_signed ? -static_cast<long long>(type_max) - 1 : 0;
    float s;
    long long z;
    if (r_max == r_min) {
        s = 1.0f;
        z = is_signed ? -1 : static_cast<long long>(type_max >> 1);
    } else {
        const double q_min = static_cast<double>(type_min), q_max = static_cast<double>(type_max);
        const double sd = __ddiv_rn(__dsub_rn(r_max, r_min), __dsub_rn(q_max, q_min));
        double zp = __dsub_rn(q_min, __ddiv_rn(r_min, sd));
        zp = fmax(fmin(static_cast<double>(dev_cvttsd_i64(round(zp))), q_max), q_min);
        s = __double2float_rn(sd);
        z = dev_cvttsd_i64(zp);
    }
    DeviceMeta m;
    m.scale = s;
    m.error = (isnan(s) || !(s >= 0.0f)) ? 1 : 0;          // the reference aborts here (src/piquant.cpp:373)
    m.zero_point = z;
    // the kernels work on the unsigned (offset-binary) view of a signed type
    z = static_cast<long long>(static_cast<unsigned long long>(z) + (is_signed ? (1ull << (bits - 1)) : 0ull));
    m.P.sign_xor = sign_xor;
    m.P.scale = s;
    m.P.inv_scale = __fdiv_rn(1.0f, s);
    m.P.xi = 0.0f;
    m.P.zp64 = z;
    m.P.zp32 = static_cast<int32_t>(static_cast<uint32_t>(static_cast<unsigned long long>(z)));
    m.P.bias = __fmul_rn(-static_cast<float>(m.P.zp32), s);
    m.P.bigzp = (z > (1ll << 29) || z < -(1ll << 29)) ? 1 : 0;
    m.P.spec_ok32 = (m.P.zp32 <= (1 << 29) && m.P.zp32 >= -(1 << 29)) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(m.pad)); ++i) m.pad[i] = 0;
    return m;
}

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One warp: {neg_min, max} of this rank in, max over all ranks out (every lane returns the same pair).
__device__ __forceinline__ void peer_exchange_max(const PeerExchange& px, float& neg_min, float& max) {
    const int lane = threadIdx.x & 31;
    const unsigned long long tag = static_cast<unsigned long long>(px.seq) << 32;
    const int slot = (px.seq & 1u) * kMaxPeers;
    if (lane < px.nranks) {
        unsigned long long* dst = px.box[lane] + static_cast<size_t>(slot + px.rank) * 2;
        st_release_sys_u64(dst, tag | __float_as_uint(neg_min));
        st_release_sys_u64(dst + 1, tag | __float_as_uint(max));
    }
    float a = -FLT_MAX, b = -FLT_MAX;
    if (lane < px.nranks) {
        const unsigned long long* src = px.box[px.rank] + static_cast<size_t>(slot + lane) * 2;
        unsigned long long w0, w1;
        do { w0 = ld_acquire_sys_u64(src); } while ((w0 >> 32) != px.seq);
        do { w1 = ld_acquire_sys_u64(src + 1); } while ((w1 >> 32) != px.seq);
        a = __uint_as_float(static_cast<uint32_t>(w0));
        b = __uint_as_float(static_cast<uint32_t>(w1));
    }
    neg_min = warp_max(a);
    max = warp_max(b);
}

// Call with ALL threads of the CTA (kThreads of them), each holding the {min, max} of what it has seen (+-inf if nothing).
// Returns after the last CTA has published the result; other CTAs return as soon as their partial is stored.
__device__ __forceinline__ void cta_reduce_tail(float mn, float mx, const ReduceTail& t) {
    __shared__ float s_mn[kThreads / 32], s_mx[kThreads / 32];
    __shared__ bool s_last;
    mn = warp_min(mn);
    mx = warp_max(mx);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < kThreads / 32; ++i) { mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]); }
        t.partials[blockIdx.x] = make_float2(mn, mx);
        __threadfence();
        const unsigned tk = atomicAdd(t.ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    mn = __int_as_float(0x7f800000);
    mx = __int_as_float(0xff800000);
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) {
        const float2 p = __ldcg(t.partials + i);
        mn = fminf(mn, p.x);
        mx = fmaxf(mx, p.y);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    __syncthreads();
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) { mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]); }
    mn = fminf(mn, FLT_MAX);       // nothing below FLT_MAX seen -> the reference's start value (kernels_specialized.inl:1418-1607)
    mx = fmaxf(mx, -FLT_MAX);
    float neg_min = -mn;
    if (t.px.nranks > 1) {         // the one exchange step of a sharded tensor, without leaving the kernel
        peer_exchange_max(t.px, neg_min, mx);
        mn = -neg_min;
    }
    if (lane != 0) return;
    t.result[0] = mn; t.result[1] = mx; t.result[2] = neg_min; t.result[3] = mx;
    if (t.mapped_result) { t.mapped_result[0] = mn; t.mapped_result[1] = mx; t.mapped_result[2] = neg_min; t.mapped_result[3] = mx; }
    if (t.meta_out) {
        const DeviceMeta m = meta_from_minmax(neg_min, mx, t.q_bits, t.q_signed, t.q_sign_xor);
        *t.meta_out = m;
        if (t.meta_out2) *t.meta_out2 = m;
        if (t.meta_mapped) *t.meta_mapped = m;
    }
    if (t.mapped_result || t.meta_mapped || t.meta_out2) __threadfence_system();
    *t.ticket = 0u;                // ready for the next launch on this stream
}

}  // namespace pq
