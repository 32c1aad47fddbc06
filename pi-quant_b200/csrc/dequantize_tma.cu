// dequantize_tma.cu -- TMA (cp.async.bulk) ring variant of the dequantize kernels for sm_100a.
//
// The float side of dequantize is the wide stream (80 % of the bytes of u8->f32 are writes), and HBM
// likes few, long, sequential request streams better than thousands of warps each storing 1-2 KiB:
// here every byte of the output leaves the SM as part of a 16 KiB bulk store issued by one thread.
//   stage  = {packed input tile, 16 KiB output tile, full mbarrier, empty mbarrier}, S stages per CTA;
//   producer (warp 0, one lane): waits `empty`, bulk-loads the packed tile -- and, for ADD, the 16 KiB
//            accumulator tile straight into the stage's OUTPUT buffer -- completion on `full`;
//   consumers (8 warps): wait `full`; each thread reads 1..8 packed bytes (LDS) [+ the accumulator, LDS.128],
//            dequantizes in registers (byte-permute conversion, dequantize_common.cuh), writes 16 bytes back
//            to the same place of the output tile (STS.128, conflict-free);
//   store  : after a consumer barrier one elected thread bulk-stores the tile (UBLKCP.G.S); when the
//            PREVIOUS tile's store has finished reading shared memory its stage is released (`empty`).
// For ADD `out` crosses HBM exactly once in each direction and never touches a register file twice.
#include <atomic>

#include "dequantize_common.cuh"
#include "pq_tma.cuh"

namespace pq {

namespace {

constexpr int kDqStages = 3;
constexpr int kDqConsumers = 256;
constexpr int kDqThreads = kDqConsumers + 32;   // warp 0 = producer
constexpr int kOutTile = 16384;                 // output bytes per tile
constexpr int kOutVecs = kOutTile / 16;

template <int BITS, int OUT_DT>
struct DqShape {
    static constexpr int OSZ = OUT_DT == DT_F32 ? 4 : 2;
    static constexpr int EV = 16 / OSZ;                     // elements per 16-byte output vector
    static constexpr int IBV = EV * BITS / 8;               // packed input bytes per output vector: 1..8
    static constexpr int IN_TILE = kOutVecs * IBV;          // 1..8 KiB
    static constexpr int STAGE = kOutTile + IN_TILE;
    static constexpr int SMEM = kDqStages * STAGE + 3 * kDqStages * 8;      // + full, empty, tile index per stage
};

template <int NB>
__device__ __forceinline__ uint32_t lds_packed(const unsigned char* p, uint32_t (&w)[(NB + 3) / 4]) {
    if constexpr (NB == 8) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        w[0] = t.x;
        w[1] = t.y;
    } else if constexpr (NB == 4) {
        w[0] = *reinterpret_cast<const uint32_t*>(p);
    } else if constexpr (NB == 2) {
        w[0] = *reinterpret_cast<const uint16_t*>(p);
    } else {
        w[0] = *p;
    }
    return 0;
}

template <int BITS, int OUT_DT, int OP>
__global__ void __launch_bounds__(kDqThreads) dequant_tma_kernel(const DequantArgs a_in) {
    DequantArgs a = a_in;
    using S = DqShape<BITS, OUT_DT>;
    constexpr int PER = 8 / BITS;
    constexpr int NWI = (S::IBV + 3) / 4;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kDqStages * S::STAGE);
    uint64_t* empty = full + kDqStages;
    long long* s_tile = reinterpret_cast<long long*>(empty + kDqStages);   // tile index of each stage, -1 = no more work

    const uint8_t* in = a.in + a.head_bytes;
    char* out = a.out + a.head_bytes * PER * S::OSZ;
    const int64_t n_vecs = a.n_items * 16 / S::IBV;          // 16-byte output vectors in the vectorised region
    const int64_t n_tiles = (n_vecs + kOutVecs - 1) / kOutVecs;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kDqStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    if (!load_device_params<BITS, OUT_DT>(a)) return;

    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            long long tile = sched_next_chunk(a.sched), chunk_end = tile + kSchedChunk;
            long long next_chunk = sched_next_chunk(a.sched);     // requested one chunk ahead: its latency is never waited for
            for (int i = 0;; ++i) {
                const int s = i % kDqStages;
                mbar_wait(empty + s, ((i / kDqStages) & 1) ^ 1);
                if (tile >= n_tiles) {                          // tell the consumers and stop
                    s_tile[s] = -1;
                    mbar_arrive(full + s);
                    break;
                }
                s_tile[s] = tile;
                const int64_t v0 = tile * kOutVecs;
                const int64_t rem = n_vecs - v0;
                const uint32_t vecs = static_cast<uint32_t>(rem < kOutVecs ? rem : kOutVecs);
                unsigned char* st = smem + s * S::STAGE;
                mbar_expect_tx(full + s, vecs * S::IBV + (OP == OP_ADD ? vecs * 16 : 0));
                tma_load_1d(st + kOutTile, in + v0 * S::IBV, vecs * S::IBV, full + s);
                if constexpr (OP == OP_ADD) tma_load_1d(st, out + v0 * 16, vecs * 16, full + s);
                if (++tile == chunk_end) {
                    tile = next_chunk;
                    chunk_end = tile + kSchedChunk;
                    next_chunk = sched_next_chunk(a.sched);
                }
            }
            sched_cta_done(a.sched);
        }
        return;
    }

    const int t = threadIdx.x - 32;
    for (int i = 0;; ++i) {
        const int s = i % kDqStages;
        mbar_wait(full + s, (i / kDqStages) & 1);
        const long long tile = s_tile[s];
        if (tile < 0) break;
        const int64_t v0 = tile * kOutVecs;
        const int64_t rem = n_vecs - v0;
        const int vecs = static_cast<int>(rem < kOutVecs ? rem : kOutVecs);
        unsigned char* st = smem + s * S::STAGE;
        uint4* ot = reinterpret_cast<uint4*>(st);
        const unsigned char* it = st + kOutTile;
        constexpr int NV = kOutVecs / kDqConsumers;      // 16-byte output vectors per thread per tile
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int vi = j * kDqConsumers + t;
            if (vecs == kOutVecs || vi < vecs) {
                uint32_t w[NWI];
                lds_packed<S::IBV>(it + vi * S::IBV, w);
                uint32_t prev[4] = {0u, 0u, 0u, 0u};
                if constexpr (OP == OP_ADD) {
                    const uint4 p = ot[vi];
                    prev[0] = p.x; prev[1] = p.y; prev[2] = p.z; prev[3] = p.w;
                }
                uint32_t o[4];
                dequant_words<BITS, OUT_DT, OP, NWI, 4>(w, prev, a, o);
                ot[vi] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        fence_proxy_async();
        consumer_barrier<kDqConsumers>();
        if (t == 0) {
            tma_store_1d(out + v0 * 16, st, static_cast<uint32_t>(vecs) * 16u);
            tma_store_commit();
            if (i > 0) {
                tma_store_wait_read<1>();                 // the previous tile's store no longer reads its stage
                mbar_arrive(empty + (i - 1) % kDqStages);
            }
        }
    }
    if (t == 0) tma_store_wait_all<0>();

    if (blockIdx.x == gridDim.x - 1) {
        const int64_t total = (a.numel + PER - 1) / PER;
        for (int64_t b = t; b < a.head_bytes; b += kDqConsumers) dequant_one_byte<BITS, OUT_DT, OP>(a, b);
        for (int64_t b = a.head_bytes + a.n_items * 16 + t; b < total; b += kDqConsumers) dequant_one_byte<BITS, OUT_DT, OP>(a, b);
    }
}

template <int BITS, int OUT_DT, int OP>
void launch_tma_cell(DequantArgs a, const LaunchCfg& cfg) {
    using S = DqShape<BITS, OUT_DT>;
    auto fn = dequant_tma_kernel<BITS, OUT_DT, OP>;
    set_dequant_fast(a, BITS, OUT_DT);
    static std::atomic<unsigned long long> configured{0};   // one bit per device (the attribute is per device); contexts on
    int dev = 0;                                             // different threads may race here: setting it twice is harmless
    PQ_CUDA_CHECK(cudaGetDevice(&dev));
    if (!(configured.load(std::memory_order_relaxed) >> (dev & 63) & 1ull)) {
        PQ_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
    int per_sm = 0;
    PQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kDqThreads, S::SMEM));
    const int64_t n_vecs = a.n_items * 16 / S::IBV;
    const int64_t n_tiles = (n_vecs + kOutVecs - 1) / kOutVecs;
    int64_t grid = static_cast<int64_t>(cfg.sm_count) * (per_sm > 0 ? per_sm : 1);
    if (n_tiles < grid) grid = n_tiles;
    if (grid < 1) grid = 1;
    launch_kernel(fn, static_cast<unsigned>(grid), kDqThreads, S::SMEM, cfg.stream, a);
    PQ_CUDA_CHECK(cudaGetLastError());
}

template <int BITS, int OUT_DT>
void launch_tma_op(const DequantArgs& a, int op, const LaunchCfg& cfg) {
    if (op == OP_ADD) launch_tma_cell<BITS, OUT_DT, OP_ADD>(a, cfg);
    else launch_tma_cell<BITS, OUT_DT, OP_SET>(a, cfg);
}

template <int OUT_DT>
void launch_tma_in(const DequantArgs& a, int dt_in, int op, const LaunchCfg& cfg) {
    switch (dt_in) {
        case DT_U8: launch_tma_op<8, OUT_DT>(a, op, cfg); break;
        case DT_U4: launch_tma_op<4, OUT_DT>(a, op, cfg); break;
        default:    launch_tma_op<2, OUT_DT>(a, op, cfg); break;
    }
}

}  // namespace

// returns 0 when the alignment of the buffers rules out bulk copies (the direct kernels run then)
int launch_dequantize_tma(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                          const LaunchCfg& cfg, const QuantParams* dP) {
    const int per = 8 / dtype_bits(dt_in);
    const int osz = dtype_bits(dt_out) / 8;
    DequantArgs a;
    a.in = static_cast<const uint8_t*>(in);
    a.out = static_cast<char*>(out);
    a.numel = numel;
    a.P = P;
    a.dP = dP;
    a.sched = cfg.sched;
    a.head_bytes = 0;
    a.n_items = 0;
    const int64_t full_bytes = numel / per;
    // smallest head (packed bytes) after which both streams are 16-byte aligned; units of 16 packed bytes follow
    for (int64_t h = 0; h < 16 && h <= full_bytes; ++h) {
        const uintptr_t o = reinterpret_cast<uintptr_t>(out) + static_cast<uintptr_t>(h) * per * osz;
        const uintptr_t i = reinterpret_cast<uintptr_t>(in) + static_cast<uintptr_t>(h);
        if (o % 16 == 0 && i % 16 == 0) {
            a.head_bytes = h;
            a.n_items = (full_bytes - h) / 16;
            break;
        }
    }
    if (a.n_items <= 0) return 0;
    if (dt_out == DT_F32) launch_tma_in<DT_F32>(a, dt_in, op, cfg);
    else launch_tma_in<DT_BF16>(a, dt_in, op, cfg);
    return 1;
}

}  // namespace pq
