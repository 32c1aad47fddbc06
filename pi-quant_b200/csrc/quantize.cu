// quantize.cu -- f32|bf16 -> uint8|uint4|uint2 streaming kernels for sm_100a.
//
// Replaces the reference's quant_generic router and its 5 SIMD quantize kernels
// (src/kernels/quantize.inl:101-149, src/kernels/kernels_specialized.inl:35-727).
//
// Work decomposition (HBM-bound, every element crosses HBM exactly once):
//   item  = the elements that produce 16 packed output bytes (16 u8 / 32 u4 / 64 u2 elements);
//           one thread owns one item: it reads the item's V*sizeof(In) contiguous input bytes with
//           32-byte LDG.256 (one full DRAM sector per instruction, so no sector is fetched twice)
//           and writes one 16-byte STG.128 -- a warp writes 512 contiguous bytes.
//   tile  = kThreads * U items; U is chosen so every thread has >= 128 B of loads in flight.
//   grid  = persistent: (resident CTAs per SM) x 148 SMs, tiles dealt round-robin so that
//           concurrently running CTAs stream neighbouring DRAM pages.
//   ragged= output bytes before the 16-byte aligned region and after the last full item are
//           produced byte-by-byte by the last CTA of the same launch (no second kernel).
// Inputs whose alignment rules out vector loads go through the byte-granular kernel.
#include "quantize_common.cuh"

namespace pq {

template <int IN_DT, int BITS, int STEP, bool A32>
__global__ void __launch_bounds__(kThreads) quant_stream_kernel(const QuantArgs a) {
    constexpr int PER = 8 / BITS;                       // elements per packed byte
    constexpr int V = 16 * PER;                         // elements per item (16 output bytes)
    constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    constexpr int NW = V * ISZ / 4;                     // input words per item
    constexpr int U = (NW * 4 >= 128) ? 1 : 128 / (NW * 4);
    constexpr int QMAX = (1 << BITS) - 1;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * U;

    const char* in = a.in + a.head_bytes * PER * ISZ;
    uint8_t* out = a.out + a.head_bytes;
    const int64_t n_tiles = (a.n_items + TILE - 1) / TILE;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t first = tile * TILE + threadIdx.x;
        uint32_t w[U][NW];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t item = first + static_cast<int64_t>(u) * kThreads;
            if (item < a.n_items) load_words<NW, A32>(in + item * (NW * 4), w[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t item = first + static_cast<int64_t>(u) * kThreads;
            if (item < a.n_items) {
                uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const uint32_t q = static_cast<uint32_t>(quant_step<STEP>(item_elem<IN_DT, NW>(w[u], e), a.P, QMAX));
                    o[(e * BITS) / 32] |= q << ((e * BITS) % 32);
                }
                stg_stream(out + item * 16, o);
            }
        }
    }

    if (blockIdx.x == gridDim.x - 1) {
        const int64_t total = (a.numel + PER - 1) / PER;
        for (int64_t b = threadIdx.x; b < a.head_bytes; b += kThreads) quant_one_byte<IN_DT, BITS, STEP>(a, b);
        for (int64_t b = a.head_bytes + a.n_items * 16 + threadIdx.x; b < total; b += kThreads)
            quant_one_byte<IN_DT, BITS, STEP>(a, b);
    }
}

// Any alignment: one thread per packed output byte (loads stay sector-coalesced through L1).
template <int IN_DT, int BITS, int STEP>
__global__ void __launch_bounds__(kThreads) quant_bytes_kernel(const QuantArgs a) {
    constexpr int PER = 8 / BITS;
    const int64_t total = (a.numel + PER - 1) / PER;
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; b < total;
         b += static_cast<int64_t>(gridDim.x) * kThreads)
        quant_one_byte<IN_DT, BITS, STEP>(a, b);
}

// ---------------------------------------------------------------------------------------------
// dispatch (replaces the constexpr fn-pointer tables of src/kernels/kernels.inl:108-121)
// ---------------------------------------------------------------------------------------------

using QuantKernel = void (*)(const QuantArgs);

template <int IN_DT, int BITS, int STEP>
static void launch_cell(const QuantArgs& a0, bool vec, bool a32, const LaunchCfg& cfg) {
    QuantArgs a = a0;
    constexpr int PER = 8 / BITS;
    constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    constexpr int NW = 16 * PER * ISZ / 4;
    constexpr int U = (NW * 4 >= 128) ? 1 : 128 / (NW * 4);
    QuantKernel fn;
    int64_t blocks_needed;
    if (vec) {
        fn = a32 ? quant_stream_kernel<IN_DT, BITS, STEP, true> : quant_stream_kernel<IN_DT, BITS, STEP, false>;
        const int64_t tile = static_cast<int64_t>(kThreads) * U;
        blocks_needed = (a.n_items + tile - 1) / tile;
    } else {
        fn = quant_bytes_kernel<IN_DT, BITS, STEP>;
        a.head_bytes = 0;
        a.n_items = 0;
        const int64_t total = (a.numel + PER - 1) / PER;
        blocks_needed = (total + kThreads - 1) / kThreads;
    }
    int per_sm = 0;
    PQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    int64_t grid = static_cast<int64_t>(cfg.sm_count) * (per_sm > 0 ? per_sm : 1);
    if (blocks_needed < grid) grid = blocks_needed;
    if (grid < 1) grid = 1;
    fn<<<static_cast<unsigned>(grid), kThreads, 0, cfg.stream>>>(a);
    PQ_CUDA_CHECK(cudaGetLastError());
}

template <int IN_DT, int BITS>
static void launch_mode(const QuantArgs& a, int mode, bool vec, bool a32, const LaunchCfg& cfg) {
    if (mode == 1) launch_cell<IN_DT, BITS, STEP_STOCH>(a, vec, a32, cfg);
    else if (IN_DT == DT_F32 && BITS == 2) launch_cell<IN_DT, BITS, STEP_ROUND64>(a, vec, a32, cfg);   // no SIMD body in the reference: quantize.inl:132-148
    else launch_cell<IN_DT, BITS, STEP_BODY>(a, vec, a32, cfg);
}

template <int IN_DT>
static void launch_out(const QuantArgs& a, int dt_out, int mode, bool vec, bool a32, const LaunchCfg& cfg) {
    switch (dt_out) {
        case DT_U8: launch_mode<IN_DT, 8>(a, mode, vec, a32, cfg); break;
        case DT_U4: launch_mode<IN_DT, 4>(a, mode, vec, a32, cfg); break;
        default:    launch_mode<IN_DT, 2>(a, mode, vec, a32, cfg); break;
    }
}

int launch_quantize_tma(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int mode,
                        const LaunchCfg& cfg);   // quantize_tma.cu; returns 0 when the cell / alignment is not covered

int launch_quantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int mode,
                    const LaunchCfg& cfg) {
    if (numel <= 0) return 0;
    if (cfg.variant == 2) {
        const int n = launch_quantize_tma(in, dt_in, out, dt_out, numel, P, mode, cfg);
        if (n) return n;
    }
    const int per = 8 / dtype_bits(dt_out);
    const int isz = dtype_bits(dt_in) / 8;
    QuantArgs a;
    a.in = static_cast<const char*>(in);
    a.out = static_cast<uint8_t*>(out);
    a.numel = numel;
    a.P = P;
    const int64_t total_bytes = (numel + per - 1) / per;
    const int64_t full_bytes = numel / per;                      // bytes whose elements all exist
    int64_t head = static_cast<int64_t>((16 - (reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
    if (head > full_bytes) head = full_bytes;
    a.head_bytes = head;
    a.n_items = (full_bytes - head) / 16;
    const uintptr_t in_vec = reinterpret_cast<uintptr_t>(in) + static_cast<uintptr_t>(head) * per * isz;
    const bool vec = a.n_items > 0 && (in_vec & 15u) == 0;
    const bool a32 = (in_vec & 31u) == 0;
    (void)total_bytes;
    if (dt_in == DT_F32) launch_out<DT_F32>(a, dt_out, mode, vec, a32, cfg);
    else launch_out<DT_BF16>(a, dt_out, mode, vec, a32, cfg);
    return 1;
}

}  // namespace pq
