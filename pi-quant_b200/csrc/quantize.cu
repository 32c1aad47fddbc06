// quantize.cu -- f32|bf16 -> uint8|uint4|uint2 direct (LDG/STG) streaming kernels for sm_100a, and
// the per-cell dispatch between them and the TMA ring kernels of quantize_tma.cu.
//
// Replaces the reference's quant_generic router and its 5 SIMD quantize kernels
// (src/kernels/quantize.inl:101-149, src/kernels/kernels_specialized.inl:35-727).
//
// Work decomposition (HBM-bound, every element crosses HBM exactly once):
//   vector = 32 contiguous input bytes (8 f32 / 16 bf16): ONE LDG.256 = one full DRAM sector per
//            thread.  Vectors are dealt to threads warp-interleaved -- lane l of a warp reads vector
//            (j*kThreads + tid) -- so each load instruction of a warp covers 1 KiB of contiguous
//            memory (8 cache lines = 8 L1 wavefronts).  [Measured on B200: giving each thread
//            64-256 contiguous bytes instead makes one instruction touch 16-32 lines and the L1
//            wavefront rate, not HBM, bounds the kernel: 36 % of peak for f32->u4.]
//   tile   = kThreads * 2 vectors (16 KiB of input): every thread has 2 x 32 B of loads in flight before it converts.
//            [Measured: 4 vectors per thread cost registers (48 -> 40, 5 -> 6 CTAs per SM) and tail granularity: equal at
//            f32->u8 @1e9, 2-11 % slower on the bf16 cells and at 27 M elements; 1 vector per thread is 4-40 % slower.]
//   pack   = the 8/16 elements of a vector become 2..16 packed bytes in registers (quant_group: clamp
//            and pack fused in I2IP) and leave with one STG.{16,32,64,128}; a warp's stores are
//            contiguous, full sectors.
//   grid   = one tile per CTA, dealt by the hardware scheduler (dynamic balance across the two dies).
//   ragged = output bytes before the 16-byte aligned region and after the last full vector are
//            produced byte-by-byte by the last CTA of the same launch (no second kernel).
// Inputs whose alignment rules out vector loads go through the byte-granular kernel.
#include <cmath>
#include <cstring>

#include "quantize_common.cuh"

namespace pq {

namespace {
#ifndef PQ_VEC_PER_THREAD
#define PQ_VEC_PER_THREAD 2
#endif
// 32-byte vectors per thread: 2, measured against 1, 3 and 4 (profiles/r1_vectors_per_thread_sweep.txt).  Per-element
// stochastic rounding amortises its per-thread set-up (18 Philox round keys on the uniform datapath) over 4: measured
// 83-98 % of the copy peak for f32->u8 against 75 % with 2.
constexpr int kVecPerThread = PQ_VEC_PER_THREAD;
template <int STEP> constexpr int kVecPerThreadOf = STEP == STEP_SRPE ? 4 : kVecPerThread;

template <int OB>
__device__ __forceinline__ void store_packed(uint8_t* p, const uint32_t* o) {
    if constexpr (OB == 16) {
        const uint32_t t[4] = {o[0], o[1], o[2], o[3]};
        stg_stream(p, t);
    } else if constexpr (OB == 8) {
        asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(o[0]), "r"(o[1]) : "memory");
    } else if constexpr (OB == 4) {
        asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(o[0]) : "memory");
    } else {
        static_assert(OB == 2);
        asm volatile("st.global.L1::no_allocate.b16 [%0], %1;" ::"l"(p), "h"(static_cast<uint16_t>(o[0])) : "memory");
    }
}

// One 32-byte vector (8 f32 / 16 bf16 elements) -> packed words; `group` = index of its first element / 8 (STEP_SRPE only:
// one Philox4x32-10 call per 8 elements, 16 random bits each).
template <int IN_DT, int BITS, int STEP>
__device__ __forceinline__ void quant_vector(const QuantArgs& a, int64_t group, const uint32_t (&w)[8],
                                             uint32_t (&o)[((IN_DT == DT_F32 ? 8 : 16) * BITS + 31) / 32]) {
    if constexpr (STEP == STEP_SRPE) {
        constexpr int NG = IN_DT == DT_F32 ? 1 : 2;
        uint32_t rnd[4 * NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            uint32_t r[4];
            srpe_words(a, group + g, r);
#pragma unroll
            for (int k = 0; k < 4; ++k) rnd[4 * g + k] = r[k];
        }
        quant_group<IN_DT, BITS, STEP, 8>(w, a.P, o, rnd);
    } else {
        quant_group<IN_DT, BITS, STEP, 8>(w, a.P, o);
    }
}
}  // namespace

// One tile (kThreads * J vectors) of the vectorised region, and -- for the CTA flagged `last` -- the ragged head / tail bytes.
template <int IN_DT, int BITS, int STEP>
__device__ __forceinline__ void quant_tile(const QuantArgs& a, uint32_t tile, bool last) {
    constexpr int PER = 8 / BITS;                       // elements per packed byte
    constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    constexpr int EV = 32 / ISZ;                        // elements per 32-byte vector
    constexpr int OB = EV * BITS / 8;                   // packed output bytes per vector: 2..16
    constexpr int J = kVecPerThreadOf<STEP>;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * J;

    const char* in = a.in_body;
    uint8_t* out = a.out_body;
    // STEP_SRPE: Philox counter (= element index / 8) of vector 0; the host guarantees that vectors start on multiples of 8
    [[maybe_unused]] const int64_t vec_group0 = (a.sr_base + a.head_bytes * PER) >> 3;
    const int64_t first = static_cast<int64_t>(tile) * TILE + threadIdx.x;
    uint32_t w[J][8];
    if (tile < a.n_full_tiles) {
#pragma unroll
        for (int j = 0; j < J; ++j) ldg_stream(in + (first + static_cast<int64_t>(j) * kThreads) * 32, w[j]);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            uint32_t o[(OB + 3) / 4];
            quant_vector<IN_DT, BITS, STEP>(a, vec_group0 + (first + static_cast<int64_t>(j) * kThreads) * (EV / 8), w[j], o);
            store_packed<OB>(out + (first + static_cast<int64_t>(j) * kThreads) * OB, o);
        }
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int64_t v = first + static_cast<int64_t>(j) * kThreads;
            if (v < a.n_vecs) {
                ldg_stream(in + v * 32, w[j]);
                uint32_t o[(OB + 3) / 4];
                quant_vector<IN_DT, BITS, STEP>(a, vec_group0 + v * (EV / 8), w[j], o);
                store_packed<OB>(out + v * OB, o);
            }
        }
    }

    if (last) {
        const int64_t total = (a.numel + PER - 1) / PER;
        for (int64_t b = threadIdx.x; b < a.head_bytes; b += kThreads) quant_one_byte<IN_DT, BITS, STEP>(a, b);
        for (int64_t b = a.head_bytes + a.n_items * 16 + threadIdx.x; b < total; b += kThreads)
            quant_one_byte<IN_DT, BITS, STEP>(a, b);
    }
}

// 8 CTAs per SM = 32 registers per thread: what the speculative path of every cell needs (bf16 stochastic wanted 39-40
// for its exact fallback and ran at 6 CTAs per SM); the Philox cells (4 vectors per thread) get what keeps their hot path free of spills: <= 85 (f32) / <= 128 (bf16).
template <int IN_DT, int BITS, int STEP>
__global__ void __launch_bounds__(kThreads, STEP != STEP_SRPE ? 8 : (IN_DT == DT_F32 ? 3 : 2)) quant_stream_kernel(const QuantArgs a_in) {
    QuantArgs a = a_in;
    pdl_launch_dependents();
    pdl_wait();
    if (!load_device_params(a)) return;
    // ONE tile per CTA (grid == number of tiles): the hardware CTA scheduler deals tiles to whichever SM is free.  (A
    // persistent grid with a static tile -> CTA map waits for its slowest SM -- the two dies differ by ~10 % -- and measured
    // 7 % slower: profiles/r1_sched_probe_static_vs_dynamic_tiles.txt.)  a.reverse: last tile first (LaunchCfg::reverse).
    const uint32_t tile = a.reverse ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    quant_tile<IN_DT, BITS, STEP>(a, tile, blockIdx.x == gridDim.x - 1);
}

// ---------------------------------------------------------------------------------------------
// Many small tensors in ONE launch.
//
// The reference's own Python benchmark quantizes a 1e6-element tensor 1000 times (python/benchmark/benchmark.py:16-23): at
// that size one tensor is 0.7 us of HBM time and a launch costs more than the data.  Here up to kBatchMax tensors -- each
// with its own pointers, length, scale and zero point -- travel in the kernel's parameter space (40 bytes per tensor, no
// staging copy, no device-side table to keep alive), a prefix sum of tiles maps blockIdx.x to (tensor, tile) with a
// 9-step uniform binary search, and every CTA then runs exactly the single-tensor tile code above on a QuantArgs it
// derives from the descriptor: same head / body / tail split, same bytes as a piquant_quantize call per tensor.
// ---------------------------------------------------------------------------------------------
constexpr int kBatchMax = 256;

struct BatchDesc {
    const char* in;
    uint8_t*    out;
    int64_t     numel;
    int64_t     zp64;        // kernel view (signed dtypes: + 2^(bits-1))
    float       inv_scale;
    float       scale;
};

struct BatchArgs {
    int       count;
    float     xi;
    uint32_t  sign_xor;
    uint32_t  tile_start[kBatchMax + 1];    // tile_start[i] .. tile_start[i+1]: CTAs of tensor i
    BatchDesc t[kBatchMax];
};

// what launch_quantize() derives on the host for one tensor, on the device (both must agree: same split, same bytes)
template <int IN_DT, int BITS>
__host__ __device__ inline void batch_split(const char* in, const uint8_t* out, int64_t numel, int64_t& head_bytes, int64_t& n_items, bool& vec) {
    constexpr int PER = 8 / BITS;
    constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    const int64_t full_bytes = numel / PER;
    int64_t head = static_cast<int64_t>((16 - (reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
    if (head > full_bytes) head = full_bytes;
    head_bytes = head;
    n_items = (full_bytes - head) / 16;
    const uintptr_t in_vec = reinterpret_cast<uintptr_t>(in) + static_cast<uintptr_t>(head) * PER * ISZ;
    vec = n_items > 0 && (in_vec & 31u) == 0;
}

template <int IN_DT, int BITS>
__host__ __device__ inline uint32_t batch_tiles(const char* in, const uint8_t* out, int64_t numel) {
    constexpr int PER = 8 / BITS;
    constexpr int OB = (32 / (IN_DT == DT_F32 ? 4 : 2)) * BITS / 8;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * kVecPerThread;
    int64_t head_bytes, n_items;
    bool vec;
    batch_split<IN_DT, BITS>(in, out, numel, head_bytes, n_items, vec);
    int64_t tiles;
    if (vec) tiles = (n_items * 16 / OB + TILE - 1) / TILE;
    else tiles = ((numel + PER - 1) / PER + TILE * OB - 1) / (TILE * OB);      // byte-granular: the same bytes per CTA
    return static_cast<uint32_t>(tiles < 1 ? 1 : tiles);
}

template <int IN_DT, int BITS, int STEP>
__global__ void __launch_bounds__(kThreads, 8) quant_batch_kernel(const __grid_constant__ BatchArgs b) {
    static_assert(STEP != STEP_SRPE, "the batch entry point serves nearest and per-call stochastic rounding");
    constexpr int PER = 8 / BITS;
    constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    constexpr int OB = (32 / ISZ) * BITS / 8;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * kVecPerThread;
    pdl_launch_dependents();
    // which tensor does this CTA serve?  largest i with tile_start[i] <= blockIdx.x (uniform across the CTA)
    int lo = 0, hi = b.count;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (b.tile_start[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    const BatchDesc& d = b.t[lo];
    const uint32_t tile = blockIdx.x - b.tile_start[lo];
    const bool last = blockIdx.x + 1 == b.tile_start[lo + 1];
    QuantArgs a;
    a.in = d.in;
    a.out = d.out;
    a.numel = d.numel;
    a.P.inv_scale = d.inv_scale;
    a.P.scale = d.scale;
    a.P.xi = b.xi;
    a.P.bias = 0.0f;
    a.P.zp64 = d.zp64;
    a.P.zp32 = static_cast<int32_t>(static_cast<uint32_t>(static_cast<unsigned long long>(d.zp64)));
    a.P.bigzp = (d.zp64 > (1ll << 29) || d.zp64 < -(1ll << 29)) ? 1 : 0;
    a.P.spec_ok32 = (a.P.zp32 <= (1 << 29) && a.P.zp32 >= -(1 << 29)) ? 1 : 0;
    a.P.sign_xor = b.sign_xor;
    a.dP = nullptr;
    a.sr_base = 0;
    bool vec;
    batch_split<IN_DT, BITS>(d.in, d.out, d.numel, a.head_bytes, a.n_items, vec);
    pdl_wait();
    if (vec) {
        a.n_vecs = a.n_items * 16 / OB;
        a.n_full_tiles = static_cast<uint32_t>(a.n_vecs / TILE);
        a.in_body = a.in + a.head_bytes * PER * ISZ;
        a.out_body = a.out + a.head_bytes;
        quant_tile<IN_DT, BITS, STEP>(a, tile, last);
    } else {
        const int64_t total = (d.numel + PER - 1) / PER;
        const int64_t b0 = static_cast<int64_t>(tile) * TILE * OB;
        const int64_t b1 = b0 + TILE * OB < total ? b0 + TILE * OB : total;
        for (int64_t x = b0 + threadIdx.x; x < b1; x += kThreads) quant_one_byte<IN_DT, BITS, STEP>(a, x);
    }
}

// ---------------------------------------------------------------------------------------------
// bf16 -> 2-bit by threshold compares.
//
// At 2.25 bytes per element the exact per-element steps (7-8 instructions, one of them an XU-pipe conversion) leave this
// cell bound by instruction issue, not by HBM, whenever the 1 kW cap holds the SMs near 1.5 GHz (87 % of the copy peak).
// But a 2-bit quantizer is a step function with three steps, and -- for a positive scale and away from the wrap-around of
// the int32 add -- a monotone one: x -> RN(x * inv) -> RN(. +- 0.5) -> trunc -> + zp -> clamp never decreases (the stochastic
// step trunc(r) + sign * [xi < frac] with its one threshold per call is monotone too).  So q(x) = #{k : x >= T_k} with three
// bf16 thresholds T_1 <= T_2 <= T_3, found on the host by bisection over the ordered bf16 values with a bit-exact host
// replica of the step (quant_thresholds below) -- and bf16 compares come two per instruction:
//   PRMT (pair element e with element e + 8), 3 x HSET2.BF16.GE, 3 x HADD2.BF16 (count), IMAD.SHL + LOP3 (place) per TWO
//   elements, plus 1/2 HMNMX2.|abs|.NaN per element for the witness "every |x| <= X" (X = where |x * inv| reaches 2^29; NaN, inf and
//   huge values fail it and the vector is redone by the exact steps): ~4 instructions per element instead of 7-8, no XU.
// Exhaustively checked against the oracle over all 65536 bf16 inputs (tests/test_gpu_parity.py).
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t bf16x2_ge_one(uint32_t a, uint32_t b) {       // bf16 1.0 per half where a >= b, else 0.0
    uint32_t d;
    asm("set.ge.bf16x2.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_max_abs_nan(uint32_t a, uint32_t b) {  // magnitudes: max(|a|, |b|) per half, NaN wins
    uint32_t d;
    asm("max.NaN.xorsign.abs.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// The compares (HSET2), the pairing permute, the witness (HMNMX2) and the final merge all run on the half-rate ALU pipe,
// which a first version with mask-valued compares saturated (80 % ALU, FMA pipe idle: profiles/r1_ncu_bf16_u2_threshold_kernel.txt).
// So the count of passed thresholds is formed ARITHMETICALLY on the FMA pipe: the compares return bf16 1.0 / 0.0,
// 128.0 + s3 + s2 + s1 is exact in bf16 (ulp 1 in [128, 256)) and carries the count in the two lowest mantissa bits of each
// half; one integer multiply (IMAD.SHL, FMA pipe too) moves both fields to their place and one LOP3 merges them:
// per pair of elements 5 ALU-pipe + 4 FMA-pipe instructions instead of 7 ALU-pipe ones.
template <int STEP>
__device__ __forceinline__ uint32_t quant_vector_bf16_u2_thr(const QuantArgs& a, const uint32_t (&w)[8]) {
    uint32_t mx = bf16x2_max_abs_nan(w[0], w[1]);
#pragma unroll
    for (int k = 2; k < 8; ++k) mx = bf16x2_max_abs_nan(mx, w[k]);
    mx &= 0x7fff7fffu;
    uint32_t o;
    if ((mx & 0xffffu) <= a.thr_xlim && (mx >> 16) <= a.thr_xlim) {
        o = 0u;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            // low half: element e, high half: element e + 8 -> their fields sit at bits 2e and 16 + 2e
            const uint32_t v = __byte_perm(w[e >> 1], w[(e >> 1) + 4], (e & 1) ? 0x7632u : 0x5410u);
            uint32_t cnt = bf16x2_add(bf16x2_ge_one(v, a.thr[2]), 0x43004300u);      // 128.0 + [x >= T3]
            cnt = bf16x2_add(cnt, bf16x2_ge_one(v, a.thr[1]));
            cnt = bf16x2_add(cnt, bf16x2_ge_one(v, a.thr[0]));
            o |= (cnt << (2 * e)) & (0x00030003u << (2 * e));
        }
        o ^= a.P.sign_xor;
    } else {
        uint32_t oo[1];
        quant_group<DT_BF16, 2, STEP, 8>(w, a.P, oo);
        o = oo[0];
    }
    return o;
}

// 8 CTAs per SM (32 registers): the fast path needs no more, and the rarely taken exact fallback may spill
template <int STEP>
__global__ void __launch_bounds__(kThreads, 8) quant_bf16_u2_threshold_kernel(const QuantArgs a_in) {
    QuantArgs a = a_in;
    constexpr int J = kVecPerThread;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * J;
    const char* in = a.in_body;
    uint8_t* out = a.out_body;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tile = a.reverse ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    const int64_t first = static_cast<int64_t>(tile) * TILE + threadIdx.x;
    uint32_t w[J][8];
    if (tile < a.n_full_tiles) {
#pragma unroll
        for (int j = 0; j < J; ++j) ldg_stream(in + (first + static_cast<int64_t>(j) * kThreads) * 32, w[j]);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t o[1] = {quant_vector_bf16_u2_thr<STEP>(a, w[j])};
            store_packed<4>(out + (first + static_cast<int64_t>(j) * kThreads) * 4, o);
        }
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int64_t v = first + static_cast<int64_t>(j) * kThreads;
            if (v < a.n_vecs) {
                ldg_stream(in + v * 32, w[j]);
                const uint32_t o[1] = {quant_vector_bf16_u2_thr<STEP>(a, w[j])};
                store_packed<4>(out + v * 4, o);
            }
        }
    }
    if (blockIdx.x == gridDim.x - 1) {
        const int64_t total = (a.numel + 3) / 4;
        for (int64_t b = threadIdx.x; b < a.head_bytes; b += kThreads) quant_one_byte<DT_BF16, 2, STEP>(a, b);
        for (int64_t b = a.head_bytes + a.n_items * 16 + threadIdx.x; b < total; b += kThreads) quant_one_byte<DT_BF16, 2, STEP>(a, b);
    }
}

// ---- host side of the threshold kernel: a bit-exact replica of the two steps (x86-64 SSE float arithmetic is IEEE, the
// library is compiled with -ffp-contract=off) used ONLY to place the three thresholds; no data goes through it.
namespace {
inline float host_bf16(uint32_t bits16) {
    const uint32_t u = bits16 << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
// q in the unsigned kernel view, valid while |x * inv| < 2^29, |zp| <= 2^29 (the caller guarantees both)
inline int32_t host_step_u2(float x, const QuantParams& P, int mode) {
    volatile float p = x * P.inv_scale;
    long long t;
    if (mode == 1) {                                    // quantize.inl:8-19
        const float r = p;
        const float tr = truncf(r);
        volatile float diff = r - tr;
        const float dec = fabsf(diff);
        float adj = (P.xi < dec) ? 1.0f : 0.0f;
        if (r < 0.0f) adj = -adj;
        volatile float sum = tr + adj;
        t = static_cast<long long>(sum);
    } else {                                            // kernels_specialized.inl:57-82 (one SIMD lane)
        volatile float a = p + copysignf(0.5f, p);
        t = static_cast<long long>(truncf(a));
    }
    const long long q = t + P.zp64;
    return static_cast<int32_t>(q < 0 ? 0 : (q > 3 ? 3 : q));
}
}  // namespace

// Fills a.thr / a.thr_xlim; false when the step function is not known to be monotone (then the generic kernel runs).
static bool quant_thresholds(QuantArgs& a, int mode) {
    const QuantParams& P = a.P;
    if (a.dP != nullptr) return false;                                  // parameters live on the device
    if (!(P.inv_scale > 0.0f) || !std::isfinite(P.inv_scale) || !(P.scale > 0.0f) || !std::isfinite(P.scale)) return false;
    if (P.bigzp || !P.spec_ok32) return false;
    if (mode == 1 && !(P.xi >= 0.0f && P.xi < 1.0f)) return false;
    if (mode != 0 && mode != 1) return false;
    // X: the largest bf16 magnitude (bit pattern, finite) with |x * inv| < 2^29; magnitudes are ordered like their bits
    uint32_t lo = 0, hi = 0x7f7f;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) / 2;
        volatile float p = host_bf16(mid) * P.inv_scale;
        if (p < 536870912.0f) lo = mid; else hi = mid - 1;
    }
    const uint32_t X = lo;
    if (X == 0) return false;
    // zero of either sign and every denormal must quantize alike: then no threshold falls inside (-2^-126, 2^-126) and the
    // compares do not depend on how the hardware treats signed zeros or denormal operands
    const int32_t q0 = host_step_u2(host_bf16(0u), P, mode);
    if (host_step_u2(host_bf16(0x8000u), P, mode) != q0 || host_step_u2(host_bf16(0x807fu), P, mode) != q0 ||
        host_step_u2(host_bf16(0x007fu), P, mode) != q0 || X < 0x0080u)
        return false;
    // ordered values: index i in [0, 2X]: i < X -> -(X - i), i >= X -> +(i - X)
    const auto at = [&](uint32_t i) { return i < X ? (0x8000u | (X - i)) : (i - X); };
    for (int k = 1; k <= 3; ++k) {
        uint32_t bits;
        if (host_step_u2(host_bf16(at(0)), P, mode) >= k) bits = 0xff80u;              // every x in the domain passes: -inf
        else if (host_step_u2(host_bf16(at(2 * X)), P, mode) < k) bits = 0x7f80u;      // none does: +inf
        else {
            uint32_t l = 0, h = 2 * X;                  // q(at(l)) < k <= q(at(h))
            while (h - l > 1) {
                const uint32_t mid = l + (h - l) / 2;
                if (host_step_u2(host_bf16(at(mid)), P, mode) >= k) h = mid; else l = mid;
            }
            bits = at(h);
        }
        a.thr[k - 1] = bits | (bits << 16);
    }
    a.thr_xlim = X;
    return true;
}

// Any alignment: one thread per packed output byte (loads stay sector-coalesced through L1).
template <int IN_DT, int BITS, int STEP>
__global__ void __launch_bounds__(kThreads) quant_bytes_kernel(const QuantArgs a_in) {
    QuantArgs a = a_in;
    constexpr int PER = 8 / BITS;
    const int64_t total = (a.numel + PER - 1) / PER;
    pdl_launch_dependents();
    pdl_wait();
    if (!load_device_params(a)) return;
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; b < total;
         b += static_cast<int64_t>(gridDim.x) * kThreads)
        quant_one_byte<IN_DT, BITS, STEP>(a, b);
}

// ---------------------------------------------------------------------------------------------
// dispatch (replaces the constexpr fn-pointer tables of src/kernels/kernels.inl:108-121)
// ---------------------------------------------------------------------------------------------

using QuantKernel = void (*)(const QuantArgs);

template <int IN_DT, int BITS, int STEP>
static void launch_cell(const QuantArgs& a0, bool vec, const LaunchCfg& cfg) {
    QuantArgs a = a0;
    constexpr int PER = 8 / BITS;
    constexpr int OB = (32 / (IN_DT == DT_F32 ? 4 : 2)) * BITS / 8;
    QuantKernel fn;
    int64_t blocks_needed;
    if (vec) {
        fn = quant_stream_kernel<IN_DT, BITS, STEP>;
        if constexpr (IN_DT == DT_BF16 && BITS == 2 && (STEP == STEP_BODY || STEP == STEP_STOCH)) {
            if (quant_thresholds(a, STEP == STEP_STOCH ? 1 : 0)) fn = quant_bf16_u2_threshold_kernel<STEP>;
        }
        const int64_t tile = static_cast<int64_t>(kThreads) * kVecPerThreadOf<STEP>;
        a.n_vecs = a.n_items * 16 / OB;
        blocks_needed = (a.n_vecs + tile - 1) / tile;
        pq_assert(blocks_needed < (int64_t{1} << 31), "tensor too large for one launch (%lld tiles)", static_cast<long long>(blocks_needed));
        a.n_full_tiles = static_cast<uint32_t>(a.n_vecs / tile);
        a.in_body = a.in + a.head_bytes * PER * (IN_DT == DT_F32 ? 4 : 2);
        a.out_body = a.out + a.head_bytes;
    } else {
        fn = quant_bytes_kernel<IN_DT, BITS, STEP>;
        a.head_bytes = 0;
        a.n_items = 0;
        const int64_t total = (a.numel + PER - 1) / PER;
        blocks_needed = (total + kThreads - 1) / kThreads;
    }
    int64_t grid = blocks_needed;                      // vector kernel: one tile per CTA
    if (!vec) {                                        // byte kernel: grid-stride over a resident grid
        int per_sm = 0;
        PQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
        const int64_t resident = static_cast<int64_t>(cfg.sm_count) * (per_sm > 0 ? per_sm : 1);
        if (resident < grid) grid = resident;
    }
    if (grid < 1) grid = 1;
    launch_kernel(fn, static_cast<unsigned>(grid), kThreads, 0, cfg.stream, a);
    PQ_CUDA_CHECK(cudaGetLastError());
}

template <int IN_DT, int BITS>
static void launch_mode(const QuantArgs& a, int mode, bool vec, const LaunchCfg& cfg) {
    if (mode == 2) launch_cell<IN_DT, BITS, STEP_SRPE>(a, vec, cfg);
    else if (mode == 1) launch_cell<IN_DT, BITS, STEP_STOCH>(a, vec, cfg);
    else if (IN_DT == DT_F32 && BITS == 2) launch_cell<IN_DT, BITS, STEP_ROUND64>(a, vec, cfg);   // no SIMD body in the reference: quantize.inl:132-148
    else launch_cell<IN_DT, BITS, STEP_BODY>(a, vec, cfg);
}

template <int IN_DT>
static void launch_out(const QuantArgs& a, int dt_out, int mode, bool vec, const LaunchCfg& cfg) {
    switch (dt_out) {
        case DT_U8: launch_mode<IN_DT, 8>(a, mode, vec, cfg); break;
        case DT_U4: launch_mode<IN_DT, 4>(a, mode, vec, cfg); break;
        default:    launch_mode<IN_DT, 2>(a, mode, vec, cfg); break;
    }
}

int launch_quantize_tma(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int mode,
                        const LaunchCfg& cfg, const QuantParams* dP);   // quantize_tma.cu; returns 0 when the cell / alignment is not covered

int launch_quantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int mode,
                    const LaunchCfg& cfg, const QuantParams* dP) {
    if (numel <= 0) return 0;
    const int per = 8 / dtype_bits(dt_out);
    const int isz = dtype_bits(dt_in) / 8;
    QuantArgs a;
    a.in = static_cast<const char*>(in);
    a.out = static_cast<uint8_t*>(out);
    a.numel = numel;
    a.P = P;
    a.dP = dP;
    a.sched = nullptr;           // the direct kernels are scheduled by the hardware, one tile per CTA
    a.sr_key = PhiloxKey{static_cast<uint32_t>(cfg.sr_key), static_cast<uint32_t>(cfg.sr_key >> 32)};
    a.sr_base = cfg.sr_base;
    a.reverse = cfg.reverse ? 1u : 0u;
    const int64_t full_bytes = numel / per;                      // bytes whose elements all exist
    int64_t head = static_cast<int64_t>((16 - (reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
    if (head > full_bytes) head = full_bytes;
    a.head_bytes = head;
    a.n_items = (full_bytes - head) / 16;
    const uintptr_t in_vec = reinterpret_cast<uintptr_t>(in) + static_cast<uintptr_t>(head) * per * isz;
    const bool a16 = a.n_items > 0 && (in_vec & 15u) == 0;      // what cp.async.bulk needs
    bool a32 = a.n_items > 0 && (in_vec & 31u) == 0;            // what LDG.256 needs
    if (mode == 2) {
        // per-element stochastic rounding: direct kernels only, and the vector kernel only when every vector starts on a
        // multiple of 8 elements of the caller's tensor (one Philox call per 8 elements); the byte kernel serves the rest
        pq_assert((cfg.sr_base & 7) == 0, "sr_base must be a multiple of 8");
        if ((head * per) % 8 != 0) a32 = false;
        if (dt_in == DT_F32) launch_out<DT_F32>(a, dt_out, mode, a32, cfg);
        else launch_out<DT_BF16>(a, dt_out, mode, a32, cfg);
        return 1;
    }
    // variants 0 / 1 = direct kernels, 2 = TMA ring where alignment allows (selection notes: pq_kernels.h); an input that is
    // 16- but not 32-byte aligned goes to the TMA ring in every variant (bulk copies need 16, LDG.256 needs 32)
    const bool want_tma = cfg.variant == 2 || !a32;
    if (a16 && want_tma) {
        const int n = launch_quantize_tma(in, dt_in, out, dt_out, numel, P, mode, cfg, dP);
        if (n) return n;
    }
    if (dt_in == DT_F32) launch_out<DT_F32>(a, dt_out, mode, a32, cfg);
    else launch_out<DT_BF16>(a, dt_out, mode, a32, cfg);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// batch launch
// ---------------------------------------------------------------------------------------------

template <int IN_DT, int BITS, int STEP>
static int launch_batch_cell(const BatchItem* items, int count, int dt_out, float xi, const LaunchCfg& cfg) {
    int launches = 0;
    for (int base = 0; base < count;) {
        BatchArgs b;
        b.xi = xi;
        b.sign_xor = dtype_sign_xor(dt_out);
        int k = 0;
        uint64_t tiles = 0;
        while (base < count && k < kBatchMax) {
            const BatchItem& it = items[base];
            if (it.numel <= 0) { ++base; continue; }
            const QuantParams P = make_params(it.scale, it.zero_point, xi, dt_out);
            BatchDesc& d = b.t[k];
            d.in = static_cast<const char*>(it.in);
            d.out = static_cast<uint8_t*>(it.out);
            d.numel = it.numel;
            d.zp64 = P.zp64;
            d.inv_scale = P.inv_scale;
            d.scale = P.scale;
            const uint32_t t = batch_tiles<IN_DT, BITS>(d.in, d.out, d.numel);
            if (tiles + t >= (uint64_t{1} << 31)) {
                pq_assert(k > 0, "tensor too large for the batch entry point (%llu tiles)", static_cast<unsigned long long>(t));
                break;
            }
            b.tile_start[k] = static_cast<uint32_t>(tiles);
            tiles += t;
            ++k;
            ++base;
        }
        if (k == 0) break;
        b.count = k;
        for (int i = k; i <= kBatchMax; ++i) b.tile_start[i] = static_cast<uint32_t>(tiles);
        launch_kernel(quant_batch_kernel<IN_DT, BITS, STEP>, static_cast<unsigned>(tiles), kThreads, 0, cfg.stream, b);
        PQ_CUDA_CHECK(cudaGetLastError());
        ++launches;
    }
    return launches;
}

template <int IN_DT, int BITS>
static int launch_batch_mode(const BatchItem* items, int count, int dt_out, int mode, float xi, const LaunchCfg& cfg) {
    if (mode == 1) return launch_batch_cell<IN_DT, BITS, STEP_STOCH>(items, count, dt_out, xi, cfg);
    if constexpr (IN_DT == DT_F32 && BITS == 2) return launch_batch_cell<IN_DT, BITS, STEP_ROUND64>(items, count, dt_out, xi, cfg);   // quantize.inl:132-148
    else return launch_batch_cell<IN_DT, BITS, STEP_BODY>(items, count, dt_out, xi, cfg);
}

template <int IN_DT>
static int launch_batch_out(const BatchItem* items, int count, int dt_out_view, int dt_out, int mode, float xi, const LaunchCfg& cfg) {
    switch (dt_out_view) {
        case DT_U8: return launch_batch_mode<IN_DT, 8>(items, count, dt_out, mode, xi, cfg);
        case DT_U4: return launch_batch_mode<IN_DT, 4>(items, count, dt_out, mode, xi, cfg);
        default:    return launch_batch_mode<IN_DT, 2>(items, count, dt_out, mode, xi, cfg);
    }
}

int launch_quantize_batch(const BatchItem* items, int count, int dt_in, int dt_out_view, int dt_out, int mode, float xi, const LaunchCfg& cfg) {
    pq_assert(mode == 0 || mode == 1, "the batch entry point serves nearest and per-call stochastic rounding, not mode %d", mode);
    if (dt_in == DT_F32) return launch_batch_out<DT_F32>(items, count, dt_out_view, dt_out, mode, xi, cfg);
    return launch_batch_out<DT_BF16>(items, count, dt_out_view, dt_out, mode, xi, cfg);
}

}  // namespace pq
