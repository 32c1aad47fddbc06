// quantize_tma.cu -- TMA (cp.async.bulk) ring variant of the quantize kernels for sm_100a.
//
// Same arithmetic and same item / ragged decomposition as quantize.cu; what changes is how bytes move:
//   producer (warp 0, one lane)  : cp.async.bulk global -> shared, 16 KiB input tiles into an
//                                  S-stage ring, completion counted on a `full` mbarrier per stage
//                                  (SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64);
//   consumers (8 warps)          : wait `full`, read 16-byte vectors from shared memory
//                                  (conflict-free LDS.128), quantize + pack in registers, write the
//                                  packed bytes to a shared-memory output tile, arrive on `empty`;
//   store                        : one elected consumer issues cp.async.bulk shared -> global of the
//                                  packed tile (UBLKCP.G.S), double-buffered with bulk-group waits.
// No thread computes a global address per element and the loads in flight are bounded by shared
// memory (S x 16 KiB per CTA, several CTAs per SM), not by registers.
#include <atomic>

#include "pq_tma.cuh"
#include "quantize_common.cuh"

namespace pq {

namespace {

constexpr int kStages = 4;
constexpr int kTileVecs = 1024;                 // 16-byte vectors per input tile (16 KiB)
constexpr int kConsumers = 256;
constexpr int kTmaThreads = kConsumers + 32;    // warp 0 = producer

template <int IN_DT, int BITS>
struct TmaShape {
    static constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    static constexpr int EV = 16 / ISZ;                 // elements per 16-byte vector
    static constexpr int OBV = EV * BITS / 8;           // packed output bytes per vector: 1, 2, 4 or 8
    static constexpr int IN_TILE = kTileVecs * 16;
    static constexpr int OUT_TILE = kTileVecs * OBV;
    static constexpr int SMEM = kStages * IN_TILE + 2 * OUT_TILE + 3 * kStages * 8;     // + full, empty, tile index per stage
};

template <int IN_DT, int BITS, int STEP>
__global__ void __launch_bounds__(kTmaThreads) quant_tma_kernel(const QuantArgs a_in) {
    QuantArgs a = a_in;
    using S = TmaShape<IN_DT, BITS>;
    constexpr int PER = 8 / BITS;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_in = smem;
    unsigned char* s_out = smem + kStages * S::IN_TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_out + 2 * S::OUT_TILE);
    uint64_t* empty = full + kStages;
    long long* s_tile = reinterpret_cast<long long*>(empty + kStages);      // tile index of each stage, -1 = no more work

    const char* in = a.in + a.head_bytes * PER * S::ISZ;
    uint8_t* out = a.out + a.head_bytes;
    const int64_t n_vecs = a.n_items * 16 / S::OBV;     // 16-byte input vectors in the vectorised region
    const int64_t n_tiles = (n_vecs + kTileVecs - 1) / kTileVecs;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kConsumers); }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();                 // set-up above overlaps the previous kernel's tail; no global access before this line
    if (!load_device_params(a)) return;

    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            long long tile = sched_next_chunk(a.sched), chunk_end = tile + kSchedChunk;
            long long next_chunk = sched_next_chunk(a.sched);     // requested one chunk ahead: its latency is never waited for
            for (int i = 0;; ++i) {
                const int s = i % kStages;
                mbar_wait(empty + s, ((i / kStages) & 1) ^ 1);
                if (tile >= n_tiles) {                          // tell the consumers and stop
                    s_tile[s] = -1;
                    mbar_arrive(full + s);
                    break;
                }
                s_tile[s] = tile;
                const int64_t v0 = tile * kTileVecs;
                const int64_t rem = n_vecs - v0;
                const uint32_t bytes = static_cast<uint32_t>((rem < kTileVecs ? rem : kTileVecs) * 16);
                mbar_expect_tx(full + s, bytes);
                tma_load_1d(s_in + s * S::IN_TILE, in + v0 * 16, bytes, full + s);
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
        const int s = i % kStages;
        mbar_wait(full + s, (i / kStages) & 1);
        const long long tile = s_tile[s];
        if (tile < 0) break;
        const int64_t v0 = tile * kTileVecs;
        const int64_t rem = n_vecs - v0;
        const int vecs = static_cast<int>(rem < kTileVecs ? rem : kTileVecs);
        unsigned char* ob = s_out + (i & 1) * S::OUT_TILE;
        const uint4* src = reinterpret_cast<const uint4*>(s_in + s * S::IN_TILE);
        constexpr int NV = kTileVecs / kConsumers;       // 16-byte vectors per thread per tile
        uint4 v[NV];
        if (vecs == kTileVecs) {
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] = src[j * kConsumers + t];
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] = (j * kConsumers + t < vecs) ? src[j * kConsumers + t] : make_uint4(0u, 0u, 0u, 0u);
        }
        // every consumer thread releases the input stage itself.  (One elected arrive per warp after __syncwarp() is
        // equivalent in the PTX memory model but compute-sanitizer's racecheck does not chain the two and reported the
        // stage reuse as a hazard; with per-thread arrives it reports none -- and the kernel measured 0-7 % faster.)
        mbar_arrive(empty + s);
#pragma unroll
        for (int j = 0; j < NV; j += 2) {                // two vectors per speculative group
            const uint32_t w[8] = {v[j].x, v[j].y, v[j].z, v[j].w, v[j + 1].x, v[j + 1].y, v[j + 1].z, v[j + 1].w};
            uint32_t o[(2 * S::OBV + 3) / 4];
            quant_group<IN_DT, BITS, STEP, 8>(w, a.P, o);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int vi = (j + h) * kConsumers + t;
                if (vecs == kTileVecs || vi < vecs) {
                    if constexpr (S::OBV == 8) *reinterpret_cast<uint2*>(ob + vi * 8) = make_uint2(o[2 * h], o[2 * h + 1]);
                    else if constexpr (S::OBV == 4) *reinterpret_cast<uint32_t*>(ob + vi * 4) = o[h];
                    else if constexpr (S::OBV == 2) *reinterpret_cast<uint16_t*>(ob + vi * 2) = static_cast<uint16_t>(o[0] >> (16 * h));
                    else ob[vi] = static_cast<uint8_t>(o[0] >> (8 * h));
                }
            }
        }
        fence_proxy_async();
        if (t == 0) tma_store_wait_read<0>();           // the previous tile's store has drained its buffer
        consumer_barrier<kConsumers>();
        if (t == 0) {
            tma_store_1d(out + v0 * S::OBV, ob, static_cast<uint32_t>(vecs * S::OBV));
            tma_store_commit();
        }
    }
    if (t == 0) tma_store_wait_all<0>();

    if (blockIdx.x == gridDim.x - 1) {
        const int64_t total = (a.numel + PER - 1) / PER;
        for (int64_t b = t; b < a.head_bytes; b += kConsumers) quant_one_byte<IN_DT, BITS, STEP>(a, b);
        for (int64_t b = a.head_bytes + a.n_items * 16 + t; b < total; b += kConsumers) quant_one_byte<IN_DT, BITS, STEP>(a, b);
    }
}

template <int IN_DT, int BITS, int STEP>
void launch_tma_cell(const QuantArgs& a, const LaunchCfg& cfg) {
    using S = TmaShape<IN_DT, BITS>;
    auto fn = quant_tma_kernel<IN_DT, BITS, STEP>;
    static std::atomic<unsigned long long> configured{0};   // one bit per device (the attribute is per device); contexts on
    int dev = 0;                                             // different threads may race here: setting it twice is harmless
    PQ_CUDA_CHECK(cudaGetDevice(&dev));
    if (!(configured.load(std::memory_order_relaxed) >> (dev & 63) & 1ull)) {
        PQ_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
    int per_sm = 0;
    PQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kTmaThreads, S::SMEM));
    const int64_t n_vecs = a.n_items * 16 / S::OBV;
    const int64_t n_tiles = (n_vecs + kTileVecs - 1) / kTileVecs;
    int64_t grid = static_cast<int64_t>(cfg.sm_count) * (per_sm > 0 ? per_sm : 1);
    if (n_tiles < grid) grid = n_tiles;
    if (grid < 1) grid = 1;
    launch_kernel(fn, static_cast<unsigned>(grid), kTmaThreads, S::SMEM, cfg.stream, a);
    PQ_CUDA_CHECK(cudaGetLastError());
}

template <int IN_DT, int BITS>
void launch_tma_mode(const QuantArgs& a, int mode, const LaunchCfg& cfg) {
    if (mode == 1) launch_tma_cell<IN_DT, BITS, STEP_STOCH>(a, cfg);
    else if (IN_DT == DT_F32 && BITS == 2) launch_tma_cell<IN_DT, BITS, STEP_ROUND64>(a, cfg);
    else launch_tma_cell<IN_DT, BITS, STEP_BODY>(a, cfg);
}

}  // namespace

int launch_quantize_tma(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int mode,
                        const LaunchCfg& cfg, const QuantParams* dP) {
    const int per = 8 / dtype_bits(dt_out);
    const int isz = dtype_bits(dt_in) / 8;
    QuantArgs a;
    a.in = static_cast<const char*>(in);
    a.out = static_cast<uint8_t*>(out);
    a.numel = numel;
    a.P = P;
    a.dP = dP;
    a.sched = cfg.sched;
    const int64_t full_bytes = numel / per;
    int64_t head = static_cast<int64_t>((16 - (reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
    if (head > full_bytes) head = full_bytes;
    a.head_bytes = head;
    a.n_items = (full_bytes - head) / 16;
    const uintptr_t in_vec = reinterpret_cast<uintptr_t>(in) + static_cast<uintptr_t>(head) * per * isz;
    if (a.n_items <= 0 || (in_vec & 15u) != 0) return 0;     // bulk copies need 16-byte aligned addresses
    if (dt_in == DT_F32) {
        if (dt_out == DT_U8) launch_tma_mode<DT_F32, 8>(a, mode, cfg);
        else if (dt_out == DT_U4) launch_tma_mode<DT_F32, 4>(a, mode, cfg);
        else launch_tma_mode<DT_F32, 2>(a, mode, cfg);
    } else {
        if (dt_out == DT_U8) launch_tma_mode<DT_BF16, 8>(a, mode, cfg);
        else if (dt_out == DT_U4) launch_tma_mode<DT_BF16, 4>(a, mode, cfg);
        else launch_tma_mode<DT_BF16, 2>(a, mode, cfg);
    }
    return 1;
}

}  // namespace pq
