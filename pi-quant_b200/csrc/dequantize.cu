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
#include "pq_reduce.cuh"

namespace pq {

#ifndef PQ_DEQUANT_U
#define PQ_DEQUANT_U 2
#endif
constexpr int kDequantItemsPerThread = PQ_DEQUANT_U;

// What a fused launch does besides dequantizing (the two halves of a quantized ring reduction, reference README.md:29):
//   FUSE_MINMAX  (with ADD): min/max of the values this launch WROTE, folded by the pq_reduce.cuh tail into the parameter
//                block of the next hop -- the chunk a rank accumulates at hop s is exactly the chunk it quantizes and sends at
//                hop s + 1, so the separate min/max pass over it (a full HBM read) disappears.  The sums are stored with
//                L2::evict_last: the quantize pass that follows finds the end of the chunk still in the 126 MB L2.
//   FUSE_FORWARD (with SET): the packed input words are also stored, unchanged, to `fwd` -- the next rank's receive slot in
//                NVLink peer memory -- together with the 64-byte parameter block: one kernel is both the all-gather
//                hop's dequantize and its send.
enum : int { FUSE_NONE = 0, FUSE_MINMAX = 1, FUSE_FORWARD = 2 };

struct DequantFuse {
    ReduceTail        tail;           // FUSE_MINMAX
    uint8_t*          fwd;            // FUSE_FORWARD: same byte phase (mod 32) as the packed input
    const DeviceMeta* fwd_meta_src;   // FUSE_FORWARD: parameter block to pass on ...
    DeviceMeta*       fwd_meta_dst;   // ... and where (nullptr: none)
};

template <int BITS, int OUT_DT, int OP, bool A32, int FUSE>
__device__ __forceinline__ void dequant_stream_body(DequantArgs& a, [[maybe_unused]] const DequantFuse* f) {
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
    if (!load_device_params<BITS, OUT_DT>(a)) {
        // flagged parameters: nothing is dequantized; a fused min/max passes the flag on to the block it was to produce
        if constexpr (FUSE == FUSE_MINMAX) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (f->tail.meta_out) f->tail.meta_out->error = 1;
                if (f->tail.meta_out2) { f->tail.meta_out2->error = 1; __threadfence_system(); }
            }
        }
        return;
    }

    [[maybe_unused]] float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    [[maybe_unused]] uint32_t pmn = 0x7f807f80u, pmx = 0xff80ff80u;      // packed bf16x2 accumulators

    auto do_tile = [&](const int64_t tile) {
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
                if constexpr (FUSE == FUSE_FORWARD) store_words<NWI, A32>(f->fwd + a.head_bytes + item * IB, wi[u]);
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
                if constexpr (FUSE == FUSE_MINMAX) {
#pragma unroll
                    for (int k = 0; k < NWO; ++k) {
                        if constexpr (OUT_DT == DT_F32) {
                            mn = fminf(mn, __uint_as_float(wo[k]));
                            mx = fmaxf(mx, __uint_as_float(wo[k]));
                        } else {            // of the ROUNDED bf16 values: what a min/max pass over the stored tensor would see
                            pmn = min_bf16x2(pmn, wo[k]);
                            pmx = max_bf16x2(pmx, wo[k]);
                        }
                    }
                }
                store_words<NWO, A32, FUSE == FUSE_MINMAX>(out + item * 64, wo);
            }
        }
    };
    // one tile per CTA, hardware-scheduled (see quantize.cu).  Several tiles per CTA -- what pays in reduce_sum.cu, where the ticketed
    // tail is as expensive as a tile -- was measured here too (FUSE_MINMAX, 33.5 M elements): 52.1 us against 50.7 us, the loop costs
    // 12 registers and with them the fourth CTA per SM.
    if (blockIdx.x < n_tiles) do_tile(blockIdx.x);

    if (blockIdx.x == gridDim.x - 1) {
        const int64_t total = (a.numel + PER - 1) / PER;
        auto ragged = [&](int64_t b) {
            dequant_one_byte<BITS, OUT_DT, OP>(a, b);
            if constexpr (FUSE == FUSE_FORWARD) f->fwd[b] = a.in[b];
            if constexpr (FUSE == FUSE_MINMAX) {        // fold what this thread has just stored
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int64_t e = b * PER + k;
                    if (e < a.numel) {
                        float v;
                        if constexpr (OUT_DT == DT_F32) v = reinterpret_cast<const float*>(a.out)[e];
                        else v = bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(a.out)[e]);
                        mn = fminf(mn, v);
                        mx = fmaxf(mx, v);
                    }
                }
            }
        };
        for (int64_t b = threadIdx.x; b < a.head_bytes; b += kThreads) ragged(b);
        for (int64_t b = a.head_bytes + a.n_items * IB + threadIdx.x; b < total; b += kThreads) ragged(b);
        if constexpr (FUSE == FUSE_FORWARD) {
            if (f->fwd_meta_dst && threadIdx.x < 16)       // 64 bytes = 16 words; the block was written by an earlier kernel on the stream
                reinterpret_cast<uint32_t*>(f->fwd_meta_dst)[threadIdx.x] = reinterpret_cast<const uint32_t*>(f->fwd_meta_src)[threadIdx.x];
        }
    }
    if constexpr (FUSE == FUSE_MINMAX) {
        if constexpr (OUT_DT == DT_BF16) {
            mn = fminf(mn, fminf(bf16_lo(pmn), bf16_hi(pmn)));
            mx = fmaxf(mx, fmaxf(bf16_lo(pmx), bf16_hi(pmx)));
        }
        cta_reduce_tail(mn, mx, f->tail);
    }
}

template <int BITS, int OUT_DT, int OP, bool A32>
__global__ void __launch_bounds__(kThreads) dequant_stream_kernel(const DequantArgs a_in) {
    DequantArgs a = a_in;
    dequant_stream_body<BITS, OUT_DT, OP, A32, FUSE_NONE>(a, nullptr);
}

// the same pass with one of the fused epilogues; `f` lives in the kernel's parameter space (constant bank)
template <int BITS, int OUT_DT, int OP, bool A32, int FUSE>
__global__ void __launch_bounds__(kThreads) dequant_fused_kernel(const DequantArgs a_in, const DequantFuse f) {
    static_assert((FUSE == FUSE_MINMAX && OP == OP_ADD) || (FUSE == FUSE_FORWARD && OP == OP_SET));
    DequantArgs a = a_in;
    dequant_stream_body<BITS, OUT_DT, OP, A32, FUSE>(a, &f);
}

// Any alignment: one thread per packed input byte.
template <int BITS, int OUT_DT, int OP>
__global__ void __launch_bounds__(kThreads) dequant_bytes_kernel(const DequantArgs a_in) {
    DequantArgs a = a_in;
    constexpr int PER = 8 / BITS;
    const int64_t total = (a.numel + PER - 1) / PER;
    pdl_launch_dependents();
    pdl_wait();
    if (!load_device_params<BITS, OUT_DT>(a)) return;
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; b < total;
         b += static_cast<int64_t>(gridDim.x) * kThreads)
        dequant_one_byte<BITS, OUT_DT, OP>(a, b);
}

// ---------------------------------------------------------------------------------------------
// dispatch (replaces the [op][dt_out][dt_in] table of src/kernels/kernels.inl:123-137)
// ---------------------------------------------------------------------------------------------

using DequantKernel = void (*)(const DequantArgs);
using DequantFusedKernel = void (*)(const DequantArgs, const DequantFuse);

// host-side request for a fused epilogue (see DequantFuse)
struct FuseRequest {
    int                  kind = FUSE_NONE;
    const MinMaxScratch* scratch = nullptr;   // FUSE_MINMAX
    const ReduceOut*     reduce = nullptr;    // FUSE_MINMAX
    void*                fwd = nullptr;       // FUSE_FORWARD
    const DeviceMeta*    fwd_meta_src = nullptr;
    DeviceMeta*          fwd_meta_dst = nullptr;
};

template <int BITS, int OUT_DT, int OP>
static int launch_cell(const void* in, void* out, int64_t numel, const QuantParams& P, const LaunchCfg& cfg, const QuantParams* dP,
                       const FuseRequest& fr) {
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
    const int64_t tile = static_cast<int64_t>(kThreads) * kDequantItemsPerThread;
    // a fused epilogue rides on the vector kernel only; the caller splits the work when this returns 0
    if (fr.kind != FUSE_NONE) {
        if (!vec) return 0;
        DequantFuse f{};
        DequantFusedKernel fn = nullptr;
        const int64_t grid = (a.n_items + tile - 1) / tile;
        if constexpr (OP == OP_ADD) {
            if (fr.kind != FUSE_MINMAX || grid > fr.scratch->max_blocks) return 0;
            f.tail = make_reduce_tail(*fr.scratch, *fr.reduce);
            fn = a32 ? dequant_fused_kernel<BITS, OUT_DT, OP_ADD, true, FUSE_MINMAX> : dequant_fused_kernel<BITS, OUT_DT, OP_ADD, false, FUSE_MINMAX>;
        } else {
            if (fr.kind != FUSE_FORWARD) return 0;
            // the forwarded copy uses the input's vector stores: same phase modulo 32 bytes required
            if ((reinterpret_cast<uintptr_t>(fr.fwd) & 31u) != (reinterpret_cast<uintptr_t>(in) & 31u)) return 0;
            f.fwd = static_cast<uint8_t*>(fr.fwd);
            f.fwd_meta_src = fr.fwd_meta_src;
            f.fwd_meta_dst = fr.fwd_meta_dst;
            fn = a32 ? dequant_fused_kernel<BITS, OUT_DT, OP_SET, true, FUSE_FORWARD> : dequant_fused_kernel<BITS, OUT_DT, OP_SET, false, FUSE_FORWARD>;
        }
        launch_kernel(fn, static_cast<unsigned>(grid < 1 ? 1 : grid), kThreads, 0, cfg.stream, a, f);
        PQ_CUDA_CHECK(cudaGetLastError());
        return 1;
    }
    DequantKernel fn;
    int64_t blocks_needed;
    if (vec) {
        fn = a32 ? dequant_stream_kernel<BITS, OUT_DT, OP, true> : dequant_stream_kernel<BITS, OUT_DT, OP, false>;
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
    return 1;
}

template <int BITS, int OUT_DT>
static int launch_op(const void* in, void* out, int64_t numel, const QuantParams& P, int op, const LaunchCfg& cfg, const QuantParams* dP,
                     const FuseRequest& fr) {
    if (op == OP_ADD) return launch_cell<BITS, OUT_DT, OP_ADD>(in, out, numel, P, cfg, dP, fr);
    return launch_cell<BITS, OUT_DT, OP_SET>(in, out, numel, P, cfg, dP, fr);
}

template <int OUT_DT>
static int launch_in(const void* in, int dt_in, void* out, int64_t numel, const QuantParams& P, int op, const LaunchCfg& cfg, const QuantParams* dP,
                     const FuseRequest& fr) {
    switch (dt_in) {
        case DT_U8: return launch_op<8, OUT_DT>(in, out, numel, P, op, cfg, dP, fr);
        case DT_U4: return launch_op<4, OUT_DT>(in, out, numel, P, op, cfg, dP, fr);
        default:    return launch_op<2, OUT_DT>(in, out, numel, P, op, cfg, dP, fr);
    }
}

static int launch_direct(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                         const LaunchCfg& cfg, const QuantParams* dP, const FuseRequest& fr) {
    if (dt_out == DT_F32) return launch_in<DT_F32>(in, dt_in, out, numel, P, op, cfg, dP, fr);
    return launch_in<DT_BF16>(in, dt_in, out, numel, P, op, cfg, dP, fr);
}

int launch_dequantize_tma(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                          const LaunchCfg& cfg, const QuantParams* dP);   // dequantize_tma.cu; returns 0 when alignment rules out bulk copies

int launch_dequantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                      const LaunchCfg& cfg, const QuantParams* dP) {
    if (numel <= 0) return 0;
    // variant 2: the TMA ring kernel whenever both streams can be 16-byte aligned; 0 (auto) and 1: the direct kernels
    // (selection table: pq_kernels.h)
    if (cfg.variant == 2) {
        const int n = launch_dequantize_tma(in, dt_in, out, dt_out, numel, P, op, cfg, dP);
        if (n) return n;
    }
    return launch_direct(in, dt_in, out, dt_out, numel, P, op, cfg, dP, FuseRequest{});
}

int launch_dequantize_add_minmax(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P,
                                 const LaunchCfg& cfg, const QuantParams* dP, const MinMaxScratch& scratch, const ReduceOut& ro) {
    pq_assert(numel > 0, "dequantize-ADD + min/max of an empty tensor");
    FuseRequest fr;
    fr.kind = FUSE_MINMAX;
    fr.scratch = &scratch;
    fr.reduce = &ro;
    if (const int n = launch_direct(in, dt_in, out, dt_out, numel, P, OP_ADD, cfg, dP, fr)) return n;
    // buffers that cannot be vector-aligned (or more tiles than partial slots): the two passes, same results
    const int n = launch_direct(in, dt_in, out, dt_out, numel, P, OP_ADD, cfg, dP, FuseRequest{});
    return n + launch_minmax(out, dt_out, numel, scratch, ro, cfg, true);
}

int launch_dequantize_forward(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P,
                              const LaunchCfg& cfg, const QuantParams* dP, void* fwd, const DeviceMeta* fwd_meta_src,
                              DeviceMeta* fwd_meta_dst) {
    pq_assert(numel > 0, "dequantize + forward of an empty tensor");
    FuseRequest fr;
    fr.kind = FUSE_FORWARD;
    fr.fwd = fwd;
    fr.fwd_meta_src = fwd_meta_src;
    fr.fwd_meta_dst = fwd_meta_dst;
    if (const int n = launch_direct(in, dt_in, out, dt_out, numel, P, OP_SET, cfg, dP, fr)) return n;
    PQ_CUDA_CHECK(cudaMemcpyAsync(fwd, in, packed_bytes(dt_in, static_cast<size_t>(numel)), cudaMemcpyDeviceToDevice, cfg.stream));
    if (fwd_meta_dst) PQ_CUDA_CHECK(cudaMemcpyAsync(fwd_meta_dst, fwd_meta_src, sizeof(DeviceMeta), cudaMemcpyDeviceToDevice, cfg.stream));
    return launch_direct(in, dt_in, out, dt_out, numel, P, OP_SET, cfg, dP, FuseRequest{});
}

}  // namespace pq
