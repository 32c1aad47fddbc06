// requantize.cu -- fused quantize -> dequantize ("fake quantization") for sm_100a, output unpacked
// in the input's float type.  Replaces requant_generic (src/kernels/kernels.inl:30-52), reached in
// the reference through context::quantize_dequantize_fused (src/piquant.cpp:342-369).
//
// Arithmetic is the reference's scalar path for every cell: quant_step_scalar (std::round or the
// per-call stochastic threshold, int64) followed by the generic dequant_step
// (src/kernels/dequantize.inl:8-11), which for bf16 multiplies in bf16 arithmetic with the scale
// rounded to bf16 (include/piquant.hpp:97-103).
//
// One thread owns 32 contiguous input bytes (LDG.256) and writes the matching 32 output bytes
// (STG.256); 4 items per thread per tile.  The quantized integer never leaves registers, so the
// pass costs one read and one write of the float tensor (plus one read of `out` for ADD).
#include <cstring>

#include "pq_kernels.h"

namespace pq {

#ifndef PQ_REQUANT_U
#define PQ_REQUANT_U 4
#endif
constexpr int kRequantItemsPerThread = PQ_REQUANT_U;

struct RequantArgs {
    const char* in;
    char*       out;
    int64_t     numel;
    int64_t     head;      // elements in front of the 32-byte aligned region
    int64_t     n_items;   // 32-byte items
    QuantParams P;
    float       scale_bf16;   // scale rounded to bf16 and widened again
    const QuantParams* dP;    // not null: parameters produced on the device (params_kernel), read from there
    PhiloxKey   sr_key;       // STEP_SRPE (per-element stochastic rounding): key of this call
    int64_t     sr_base;      // ... and the index of element 0 of this launch in the caller's tensor (multiple of 8)
};

// STEP_SRPE: 1 + u of element `e` of the caller's tensor (one Philox call; the vector path shares a call between 8 elements)
__device__ __forceinline__ float requant_one_plus_u(const RequantArgs& a, int64_t e) {
    const int64_t j = a.sr_base + e;
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(j >> 3), static_cast<uint32_t>(static_cast<uint64_t>(j >> 3) >> 32), 0u, 0u, a.sr_key, r);
    const int k = static_cast<int>(j & 7);
    return srpe_one_plus_u(r[k >> 1], k & 1);
}

__device__ __forceinline__ bool load_device_params(RequantArgs& a) {
    if (a.dP) {
        if (device_params_failed(a.dP)) return false;      // flagged block: no work (see pq_device.cuh)
        const float xi = a.P.xi;
        a.P = *a.dP;
        a.P.xi = xi;
        a.scale_bf16 = bf16_bits_to_f32(f32_to_bf16_bits(a.P.scale));
    }
    return true;
}

template <int DT, int STEP, int OP>
__device__ __forceinline__ uint32_t requant_elem(float x, uint32_t prev_bits, const RequantArgs& a, int32_t qmax, float one_plus_u = 0.0f) {
    int32_t q;
    if constexpr (STEP == STEP_SRPE) q = quant_step_srpe(x, a.P, qmax, one_plus_u);
    else q = quant_step<STEP>(x, a.P, qmax);
    const float d = a.P.bigzp
        ? __ll2float_rn(static_cast<long long>(static_cast<unsigned long long>(static_cast<long long>(q)) - static_cast<unsigned long long>(a.P.zp64)))
        : static_cast<float>(q - a.P.zp32);
    if constexpr (DT == DT_F32) {
        const float prev = __uint_as_float(prev_bits);
        return __float_as_uint(OP == OP_ADD ? __fmaf_rn(d, a.P.scale, prev) : __fmul_rn(d, a.P.scale));
    } else {
        const float db = bf16_bits_to_f32(f32_to_bf16_bits(d));
        uint16_t r = f32_to_bf16_bits(__fmul_rn(db, a.scale_bf16));
        if constexpr (OP == OP_ADD) r = f32_to_bf16_bits(__fadd_rn(bf16_bits_to_f32(static_cast<uint16_t>(prev_bits)), bf16_bits_to_f32(r)));
        return r;
    }
}

// One 32-byte vector (8 f32 / 16 bf16 elements), speculative form, entirely in floating point: requant_spec gives the
// rounded x/scale as a float (exact while |x/scale| < 2^22, one FMNMX3 witness per vector), clamp(q + zp, 0, qmax) - zp
// becomes clamp(r, -zp, qmax - zp) with two FMNMX, and that float IS the reference's float(q - zp): no F2I, no I2F.
// Valid while the bounds are exact floats (|zp| <= 2^22); for bf16 the first rounding of the reference's bf16
// arithmetic (float(q - zp) -> bf16) is the identity when 0 <= zp <= 255, where |q - zp| <= 255.  Anything else
// (huge values, NaN, extreme zero points, a threshold outside [0, 1)) redoes the vector with the exact per-element steps.
template <int DT, int STEP, int OP>
__device__ __forceinline__ void requant_vector(const uint32_t (&w)[8], const uint32_t (&p)[8], const RequantArgs& a, int32_t qmax,
                                               uint32_t (&o)[8], [[maybe_unused]] int64_t group) {
    constexpr int NE = DT == DT_F32 ? 8 : 16;
    float r[NE];
    float wit[NE];
    [[maybe_unused]] uint32_t rnd[NE / 2];             // STEP_SRPE: 16 random bits per element, `group` = index of element 0 / 8
    if constexpr (STEP == STEP_SRPE) {
#pragma unroll
        for (int g = 0; g < NE / 8; ++g) {
            uint32_t t4[4];
            philox4x32_10(static_cast<uint32_t>(group + g), static_cast<uint32_t>(static_cast<uint64_t>(group + g) >> 32), 0u, 0u, a.sr_key, t4);
#pragma unroll
            for (int k = 0; k < 4; ++k) rnd[4 * g + k] = t4[k];
        }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if constexpr (STEP == STEP_SRPE) r[e] = requant_spec_srpe(item_elem<DT, 8>(w, e), a.P, srpe_one_plus_u(rnd[e >> 1], e & 1), wit[e]);
        else r[e] = requant_spec<STEP>(item_elem<DT, 8>(w, e), a.P, wit[e]);
    }
    float m = 0.0f;
#pragma unroll
    for (int e = 0; e < NE; e += 2) m = max3_abs_nan(m, wit[e], wit[e + 1]);
    const bool zp_ok = DT == DT_F32 ? (a.P.zp64 >= -4194304 && a.P.zp64 <= 4194304) : (a.P.zp64 >= 0 && a.P.zp64 <= 255);
    const bool xi_ok = STEP != STEP_STOCH || (a.P.xi >= 0.0f && a.P.xi < 1.0f);      // (STEP_SRPE has no threshold)
    if (zp_ok && xi_ok && m < 4194304.0f) {
        const float lo = __fsub_rn(0.0f, static_cast<float>(a.P.zp32));     // +0.0, not -0.0, for zp == 0
        const float hi = static_cast<float>(qmax - a.P.zp32);
        float v[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const float d = fminf(fmaxf(r[e], lo), hi);
            if constexpr (DT == DT_F32) {
                v[e] = OP == OP_ADD ? __fmaf_rn(d, a.P.scale, __uint_as_float(p[e])) : __fmul_rn(d, a.P.scale);
            } else {
                v[e] = __fmul_rn(d, a.scale_bf16);                  // d is already a bf16 value here
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if constexpr (DT == DT_F32) {
                o[k] = __float_as_uint(v[k]);
            } else {
                uint32_t r2 = pack_bf16x2(v[2 * k], v[2 * k + 1]);
                if constexpr (OP == OP_ADD) r2 = pack_bf16x2(__fadd_rn(bf16_lo(p[k]), bf16_lo(r2)), __fadd_rn(bf16_hi(p[k]), bf16_hi(r2)));
                o[k] = r2;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if constexpr (DT == DT_F32) {
                const float u1 = STEP == STEP_SRPE ? srpe_one_plus_u(rnd[STEP == STEP_SRPE ? k >> 1 : 0], k & 1) : 0.0f;
                o[k] = requant_elem<DT, STEP, OP>(__uint_as_float(w[k]), p[k], a, qmax, u1);
            } else {
                const float u_lo = STEP == STEP_SRPE ? srpe_one_plus_u(rnd[STEP == STEP_SRPE ? k : 0], 0) : 0.0f;
                const float u_hi = STEP == STEP_SRPE ? srpe_one_plus_u(rnd[STEP == STEP_SRPE ? k : 0], 1) : 0.0f;
                const uint32_t lo = requant_elem<DT, STEP, OP>(bf16_lo(w[k]), p[k] & 0xffffu, a, qmax, u_lo);
                const uint32_t hi = requant_elem<DT, STEP, OP>(bf16_hi(w[k]), p[k] >> 16, a, qmax, u_hi);
                o[k] = lo | (hi << 16);
            }
        }
    }
}

template <int DT, int STEP, int OP>
__device__ __forceinline__ void requant_scalar(const RequantArgs& a, int64_t e, int32_t qmax) {
    if constexpr (DT == DT_F32) {
        const float x = __ldg(reinterpret_cast<const float*>(a.in) + e);
        uint32_t* o = reinterpret_cast<uint32_t*>(a.out) + e;
        *o = requant_elem<DT, STEP, OP>(x, OP == OP_ADD ? *o : 0u, a, qmax, STEP == STEP_SRPE ? requant_one_plus_u(a, e) : 0.0f);
    } else {
        const float x = bf16_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(a.in) + e));
        uint16_t* o = reinterpret_cast<uint16_t*>(a.out) + e;
        *o = static_cast<uint16_t>(requant_elem<DT, STEP, OP>(x, OP == OP_ADD ? *o : 0u, a, qmax, STEP == STEP_SRPE ? requant_one_plus_u(a, e) : 0.0f));
    }
}

template <int DT, int STEP, int OP>
__global__ void __launch_bounds__(kThreads) requant_stream_kernel(const RequantArgs a_in, const int32_t qmax) {
    RequantArgs a = a_in;
    constexpr int ISZ = DT == DT_F32 ? 4 : 2;
    constexpr int EPI = 32 / ISZ;
    constexpr int U = kRequantItemsPerThread;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * U;
    const char* in = a.in + a.head * ISZ;
    char* out = a.out + a.head * ISZ;
    const int64_t n_tiles = (a.n_items + TILE - 1) / TILE;
    pdl_launch_dependents();
    pdl_wait();
    if (!load_device_params(a)) return;

    if (const int64_t tile = blockIdx.x; tile < n_tiles) {      // one tile per CTA, hardware-scheduled (see quantize.cu)
        const int64_t first = tile * TILE + threadIdx.x;
        uint32_t w[U][8];
        uint32_t p[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t item = first + static_cast<int64_t>(u) * kThreads;
            if (item < a.n_items) {
                ldg_stream(in + item * 32, w[u]);
                if constexpr (OP == OP_ADD) ldg_rmw(out + item * 32, p[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t item = first + static_cast<int64_t>(u) * kThreads;
            if (item < a.n_items) {
                uint32_t o[8];
                requant_vector<DT, STEP, OP>(w[u], p[u], a, qmax, o, (a.sr_base + a.head + item * EPI) >> 3);
                stg_stream(out + item * 32, o);
            }
        }
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (int64_t e = threadIdx.x; e < a.head; e += kThreads) requant_scalar<DT, STEP, OP>(a, e, qmax);
        for (int64_t e = a.head + a.n_items * EPI + threadIdx.x; e < a.numel; e += kThreads) requant_scalar<DT, STEP, OP>(a, e, qmax);
    }
}

// in / out do not share a 32-byte phase: one element per thread
template <int DT, int STEP, int OP>
__global__ void __launch_bounds__(kThreads) requant_scalar_kernel(const RequantArgs a_in, const int32_t qmax) {
    RequantArgs a = a_in;
    pdl_launch_dependents();
    pdl_wait();
    if (!load_device_params(a)) return;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; e < a.numel;
         e += static_cast<int64_t>(gridDim.x) * kThreads)
        requant_scalar<DT, STEP, OP>(a, e, qmax);
}

template <int DT, int STEP, int OP>
static void launch_cell(RequantArgs a, int32_t qmax, bool vec, const LaunchCfg& cfg) {
    auto fn = vec ? requant_stream_kernel<DT, STEP, OP> : requant_scalar_kernel<DT, STEP, OP>;
    int64_t blocks_needed;
    if (vec) {
        const int64_t tile = static_cast<int64_t>(kThreads) * kRequantItemsPerThread;
        blocks_needed = (a.n_items + tile - 1) / tile;
    } else {
        a.head = 0;
        a.n_items = 0;
        blocks_needed = (a.numel + kThreads - 1) / kThreads;
    }
    int64_t grid = blocks_needed;                      // vector kernel: one tile per CTA
    if (!vec) {                                        // scalar kernel: grid-stride over a resident grid
        int per_sm = 0;
        PQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
        const int64_t resident = static_cast<int64_t>(cfg.sm_count) * (per_sm > 0 ? per_sm : 1);
        if (resident < grid) grid = resident;
    }
    if (grid < 1) grid = 1;
    launch_kernel(fn, static_cast<unsigned>(grid), kThreads, 0, cfg.stream, a, qmax);
    PQ_CUDA_CHECK(cudaGetLastError());
}

template <int DT>
static void launch_dt(const RequantArgs& a, int32_t qmax, int mode, int op, bool vec, const LaunchCfg& cfg) {
    if (mode == 2) {
        if (op == OP_ADD) launch_cell<DT, STEP_SRPE, OP_ADD>(a, qmax, vec, cfg);
        else launch_cell<DT, STEP_SRPE, OP_SET>(a, qmax, vec, cfg);
    } else if (mode == 1) {
        if (op == OP_ADD) launch_cell<DT, STEP_STOCH, OP_ADD>(a, qmax, vec, cfg);
        else launch_cell<DT, STEP_STOCH, OP_SET>(a, qmax, vec, cfg);
    } else {
        if (op == OP_ADD) launch_cell<DT, STEP_ROUND64, OP_ADD>(a, qmax, vec, cfg);
        else launch_cell<DT, STEP_ROUND64, OP_SET>(a, qmax, vec, cfg);
    }
}

static float round_to_bf16(float x) {   // include/piquant.hpp:86-90
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) u = ((u >> 16) | 64u) << 16;
    else u = ((u + (0x7fffu + ((u >> 16) & 1u))) >> 16) << 16;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int launch_requantize(const void* in, int dt_inout, void* out, int dt_quant, int64_t numel, const QuantParams& P,
                      int mode, int op, const LaunchCfg& cfg, const QuantParams* dP) {
    if (numel <= 0) return 0;
    const int isz = dtype_bits(dt_inout) / 8;
    RequantArgs a;
    a.in = static_cast<const char*>(in);
    a.out = static_cast<char*>(out);
    a.numel = numel;
    a.P = P;
    a.dP = dP;
    a.scale_bf16 = round_to_bf16(P.scale);
    const uintptr_t ia = reinterpret_cast<uintptr_t>(in), oa = reinterpret_cast<uintptr_t>(out);
    int64_t head = static_cast<int64_t>(((32 - (ia & 31u)) & 31u) / isz);
    if (head > numel) head = numel;
    a.head = head;
    a.n_items = (numel - head) / (32 / isz);
    bool vec = ((ia & 31u) == (oa & 31u)) && a.n_items > 0;
    a.sr_key = PhiloxKey{static_cast<uint32_t>(cfg.sr_key), static_cast<uint32_t>(cfg.sr_key >> 32)};
    a.sr_base = cfg.sr_base;
    if (mode == 2) {        // per-element stochastic rounding: a 32-byte item must start on a multiple of 8 elements of the tensor
        pq_assert((cfg.sr_base & 7) == 0, "sr_base must be a multiple of 8");
        if (head % 8 != 0) vec = false;
    }
    const int32_t qmax = (1 << dtype_bits(dt_quant)) - 1;
    if (dt_inout == DT_F32) launch_dt<DT_F32>(a, qmax, mode, op, vec, cfg);
    else launch_dt<DT_BF16>(a, qmax, mode, op, vec, cfg);
    return 1;
}

}  // namespace pq
