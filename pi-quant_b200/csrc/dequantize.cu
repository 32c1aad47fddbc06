// dequantize.cu -- uint8|uint4|uint2 -> f32|bf16 streaming kernels with SET / ADD store ops (sm_100a).
//
// Replaces the reference's dequant_generic router and its 5 SIMD dequantize kernels
// (src/kernels/dequantize.inl:89-140, src/kernels/kernels_specialized.inl:729-1416).
//
// Work decomposition: the float side is the wide stream here, so an item is 64 output bytes
// (16 f32 / 32 bf16 elements): one thread reads the item's 4..32 packed bytes with a single
// vector load, unpacks in registers and writes 2 x STG.256 (full 32-byte sectors).  For ADD the
// accumulator is read with 2 x LDG.256 first and the sum is formed in registers, so `out` crosses
// HBM once in each direction.  U = 2 items per thread per tile keeps 128 B of stores (and, for ADD,
// 128 B of loads) in flight per thread.  Ragged head/tail bytes are handled by the last CTA.
#include "dequantize_common.cuh"

namespace pq {

#ifndef PQ_DEQUANT_U
#define PQ_DEQUANT_U 2
#endif
constexpr int kDequantItemsPerThread = PQ_DEQUANT_U;

template <int BITS, int OUT_DT, int OP, bool A32>
__global__ void __launch_bounds__(kThreads) dequant_stream_kernel(const DequantArgs a_in) {
    DequantArgs a = a_in;
    constexpr int PER = 8 / BITS;
    constexpr int V = OUT_DT == DT_F32 ? 16 : 32;       // elements per item (64 output bytes)
    constexpr int OSZ = OUT_DT == DT_F32 ? 4 : 2;
    constexpr int IB = V * BITS / 8;                    // packed input bytes per item: 4..32
    constexpr int NWI = IB / 4;
    constexpr int NWO = 16;
    constexpr int U = kDequantItemsPerThread;
    constexpr uint32_t QMAX = (1u << BITS) - 1u;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * U;

    const uint8_t* in = a.in + a.head_bytes;
    char* out = a.out + a.head_bytes * PER * OSZ;
    const int64_t n_tiles = (a.n_items + TILE - 1) / TILE;
    pdl_launch_dependents();
    pdl_wait();
    load_device_params<BITS, OUT_DT>(a);

    if (const int64_t tile = blockIdx.x; tile < n_tiles) {      // one tile per CTA, hardware-scheduled (see quantize.cu)
        const int64_t first = tile * TILE + threadIdx.x;
        uint32_t wi[U][NWI];
        uint32_t wp[U][OP == OP_ADD ? NWO : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t item = first + static_cast<int64_t>(u) * kThreads;
            if (item < a.n_items) {
                load_words<NWI, A32>(in + item * IB, wi[u]);
                if constexpr (OP == OP_ADD) load_words_rmw<NWO, A32>(out + item * 64, wp[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t item = first + static_cast<int64_t>(u) * kThreads;
            if (item < a.n_items) {
                uint32_t wo[NWO];
#pragma unroll
                for (int k = 0; k < NWI; ++k) wi[u][k] ^= a.P.sign_xor;     // signed dtypes: two's complement -> offset binary
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const uint32_t q = (wi[u][(e * BITS) / 32] >> ((e * BITS) % 32)) & QMAX;
                    if constexpr (OUT_DT == DT_F32) {
                        const float prev = OP == OP_ADD ? __uint_as_float(wp[u][OP == OP_ADD ? e : 0]) : 0.0f;
                        wo[e] = __float_as_uint(dequant_f32<BITS, OP>(q, prev, a.P));
                    } else if ((e & 1) == 0) {
                        const uint32_t q1 = (wi[u][((e + 1) * BITS) / 32] >> (((e + 1) * BITS) % 32)) & QMAX;
                        const uint32_t pw = OP == OP_ADD ? wp[u][OP == OP_ADD ? (e >> 1) : 0] : 0u;
                        const float lo = dequant_bf16_pre<BITS, OP>(q, bf16_lo(pw), a.P);
                        const float hi = dequant_bf16_pre<BITS, OP>(q1, bf16_hi(pw), a.P);
                        wo[e >> 1] = pack_bf16x2(lo, hi);
                    }
                }
                store_words<NWO, A32>(out + item * 64, wo);
            }
        }
    }

    if (blockIdx.x == gridDim.x - 1) {
        const int64_t total = (a.numel + PER - 1) / PER;
        for (int64_t b = threadIdx.x; b < a.head_bytes; b += kThreads) dequant_one_byte<BITS, OUT_DT, OP>(a, b);
        for (int64_t b = a.head_bytes + a.n_items * IB + threadIdx.x; b < total; b += kThreads)
            dequant_one_byte<BITS, OUT_DT, OP>(a, b);
    }
}

// Any alignment: one thread per packed input byte.
template <int BITS, int OUT_DT, int OP>
__global__ void __launch_bounds__(kThreads) dequant_bytes_kernel(const DequantArgs a_in) {
    DequantArgs a = a_in;
    constexpr int PER = 8 / BITS;
    const int64_t total = (a.numel + PER - 1) / PER;
    pdl_launch_dependents();
    pdl_wait();
    load_device_params<BITS, OUT_DT>(a);
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; b < total;
         b += static_cast<int64_t>(gridDim.x) * kThreads)
        dequant_one_byte<BITS, OUT_DT, OP>(a, b);
}

// ---------------------------------------------------------------------------------------------
// dispatch (replaces the [op][dt_out][dt_in] table of src/kernels/kernels.inl:123-137)
// ---------------------------------------------------------------------------------------------

using DequantKernel = void (*)(const DequantArgs);

template <int BITS, int OUT_DT, int OP>
static void launch_cell(const void* in, void* out, int64_t numel, const QuantParams& P, const LaunchCfg& cfg, const QuantParams* dP) {
    constexpr int PER = 8 / BITS;
    constexpr int V = OUT_DT == DT_F32 ? 16 : 32;
    constexpr int OSZ = OUT_DT == DT_F32 ? 4 : 2;
    constexpr int IB = V * BITS / 8;
    DequantArgs a;
    a.in = static_cast<const uint8_t*>(in);
    a.out = static_cast<char*>(out);
    a.numel = numel;
    a.P = P;
    a.dP = dP;
    a.sched = nullptr;           // the direct kernels are scheduled by the hardware, one tile per CTA
    a.head_bytes = 0;
    a.n_items = 0;
    set_dequant_fast(a, BITS, OUT_DT);
    const int64_t full_bytes = numel / PER;
    // smallest head (in packed bytes) after which `out` is 32- (else 16-) byte aligned and `in` is
    // aligned for its vector load
    bool vec = false, a32 = false;
    for (int pass = 0; pass < 2 && !vec; ++pass) {
        const uintptr_t oalign = pass == 0 ? 32 : 16;
        const uintptr_t ialign = (pass == 0 || IB < 16) ? IB : 16;
        for (int64_t h = 0; h < 64 && h <= full_bytes; ++h) {
            const uintptr_t o = reinterpret_cast<uintptr_t>(out) + static_cast<uintptr_t>(h) * PER * OSZ;
            const uintptr_t i = reinterpret_cast<uintptr_t>(in) + static_cast<uintptr_t>(h);
            if (o % oalign == 0 && i % ialign == 0) {
                const int64_t items = (full_bytes - h) / IB;
                if (items > 0) {
                    vec = true;
                    a32 = pass == 0;
                    a.head_bytes = h;
                    a.n_items = items;
                }
                break;
            }
        }
    }
    DequantKernel fn;
    int64_t blocks_needed;
    if (vec) {
        fn = a32 ? dequant_stream_kernel<BITS, OUT_DT, OP, true> : dequant_stream_kernel<BITS, OUT_DT, OP, false>;
        const int64_t tile = static_cast<int64_t>(kThreads) * kDequantItemsPerThread;
        blocks_needed = (a.n_items + tile - 1) / tile;
    } else {
        fn = dequant_bytes_kernel<BITS, OUT_DT, OP>;
        const int64_t total = (numel + PER - 1) / PER;
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

template <int BITS, int OUT_DT>
static void launch_op(const void* in, void* out, int64_t numel, const QuantParams& P, int op, const LaunchCfg& cfg, const QuantParams* dP) {
    if (op == OP_ADD) launch_cell<BITS, OUT_DT, OP_ADD>(in, out, numel, P, cfg, dP);
    else launch_cell<BITS, OUT_DT, OP_SET>(in, out, numel, P, cfg, dP);
}

template <int OUT_DT>
static void launch_in(const void* in, int dt_in, void* out, int64_t numel, const QuantParams& P, int op, const LaunchCfg& cfg, const QuantParams* dP) {
    switch (dt_in) {
        case DT_U8: launch_op<8, OUT_DT>(in, out, numel, P, op, cfg, dP); break;
        case DT_U4: launch_op<4, OUT_DT>(in, out, numel, P, op, cfg, dP); break;
        default:    launch_op<2, OUT_DT>(in, out, numel, P, op, cfg, dP); break;
    }
}

int launch_dequantize_tma(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                          const LaunchCfg& cfg, const QuantParams* dP);   // dequantize_tma.cu; returns 0 when alignment rules out bulk copies

int launch_dequantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                      const LaunchCfg& cfg, const QuantParams* dP) {
    if (numel <= 0) return 0;
    // variant 2: the TMA ring kernel whenever both streams can be 16-byte aligned; 1: direct kernels;
    // 0 (auto): TMA for large tensors (measured crossover, see dequantize_prefers_tma)
    const int64_t traffic = numel * (dtype_bits(dt_out) / 8) * (op == OP_ADD ? 2 : 1) + numel * dtype_bits(dt_in) / 8;
    if (cfg.variant == 2 || (cfg.variant == 0 && dequantize_prefers_tma(traffic))) {
        const int n = launch_dequantize_tma(in, dt_in, out, dt_out, numel, P, op, cfg, dP);
        if (n) return n;
    }
    if (dt_out == DT_F32) launch_in<DT_F32>(in, dt_in, out, numel, P, op, cfg, dP);
    else launch_in<DT_BF16>(in, dt_in, out, numel, P, op, cfg, dP);
    return 1;
}

}  // namespace pq
