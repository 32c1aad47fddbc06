// reduce_sum.cu -- acc += dequantize(in_0) + dequantize(in_1) + ... in ONE pass, with min/max + parameters of the sums (sm_100a).
//
// The reduce step of a quantized all-reduce on an NVSwitch box (piquant.distributed, algorithm "direct"): every rank has
// received one packed chunk [parameter block | payload] from every other rank and owns the float chunk they add up to.
// W - 1 successive piquant_dequantize(..., ADD) calls (the reference's recipe, README.md:29) would read and write the float
// accumulator W - 1 times -- 9 B/element each at 8 bits -- where one pass needs 8 + (W - 1) bytes per element in total.
// This kernel is that one pass: the accumulator item is loaded once, every source's packed item is loaded with its own
// vector load (all of them in flight together), the sources are folded IN ORDER in registers with exactly the
// per-element arithmetic of the ADD store op (dequant_f32 / dequant_bf16_pre of pq_device.cuh, a bf16 accumulator is
// rounded to bf16 after every source), so the result is bit-identical to the W - 1 separate ADD launches -- and to the
// reference's dequantize(ADD) applied W - 1 times, which is what the parity tests replay on the CPU.  The sums are
// stored once (L2::evict_last: the quantize pass that follows reads them again) and their min/max go through the
// pq_reduce.cuh tail into the parameter block of that quantize.
//
// Algorithmic bytes per element: 2 * sizeof(float type) + n_src * bits / 8  (u8, f32, 7 sources: 15 B/element,
// against 7 * 9 = 63 B/element for the separate launches).
#include "dequantize_common.cuh"
#include "pq_reduce.cuh"

namespace pq {

struct SumArgs {
    const uint8_t*     in[kMaxSumSources];    // first packed byte of every source
    const QuantParams* dP[kMaxSumSources];    // its parameters (inside a DeviceMeta block), produced on a device
    int                n_src;
    char*              out;                   // accumulator, updated in place
    int64_t            numel;
    int64_t            head_bytes;            // packed bytes in front of the vectorised region (same for every source)
    int64_t            n_items;               // full 64-byte accumulator items
    ReduceTail         tail;
};

// CTAs per SM the register allocation aims at: 64 registers (4 CTAs) hold an f32 accumulator item + 8 packed items without
// spilling; the bf16 cells carry twice the packed words per item (80 registers, 3 CTAs; 8-bit sources 128 registers, 2 CTAs).
constexpr int sum_min_blocks(int bits, int out_dt) { return out_dt == DT_F32 ? 4 : (bits == 8 ? 2 : 3); }

template <int BITS, int OUT_DT, bool A32, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) dequant_sum_kernel(const SumArgs a) {
    constexpr int PER = 8 / BITS;
    constexpr int V = OUT_DT == DT_F32 ? 16 : 32;       // elements per item (64 accumulator bytes)
    constexpr int OSZ = OUT_DT == DT_F32 ? 4 : 2;
    constexpr int IB = V * BITS / 8;                    // packed bytes per item and source: 4..32
    constexpr int NWI = IB / 4;
    constexpr int NWO = 16;

    __shared__ DequantArgs s_src[kMaxSumSources];     // per source: parameters + what dequant_words derives from them (fast flag, 2^23 + zp)
    __shared__ int s_bad;

    char* out = a.out + a.head_bytes * PER * OSZ;
    const int64_t n_tiles = (a.n_items + kThreads - 1) / kThreads;
    pdl_launch_dependents();
    pdl_wait();

    // the first tile's loads go out before anything else: accumulator item + one packed item per source
    uint32_t acc[NWO];
    uint32_t wi[kMaxSumSources][NWI];
    int64_t tile = blockIdx.x;
    auto issue_loads = [&](int64_t item) {
        load_words_rmw<NWO, A32>(out + item * 64, acc);
#pragma unroll
        for (int s = 0; s < kMaxSumSources; ++s)
            if (s < a.n_src) load_words<NWI, A32>(a.in[s] + a.head_bytes + item * IB, wi[s]);
    };
    if (tile < n_tiles) {
        const int64_t item = tile * kThreads + threadIdx.x;
        if (item < a.n_items) issue_loads(item);
    }
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < a.n_src) {
        if (device_params_failed(a.dP[threadIdx.x])) {
            s_bad = 1;
        } else {
            DequantArgs& d = s_src[threadIdx.x];
            d.in = a.in[threadIdx.x];
            d.out = a.out;
            d.numel = a.numel;
            d.P = *a.dP[threadIdx.x];
            set_dequant_fast(d, BITS, OUT_DT);
        }
    }
    __syncthreads();
    if (s_bad) {
        // a flagged source: nothing is accumulated, and the flag is passed on to the block this launch was to produce
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (a.tail.meta_out) a.tail.meta_out->error = 1;
            if (a.tail.meta_out2) { a.tail.meta_out2->error = 1; __threadfence_system(); }
        }
        return;
    }

    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    uint32_t pmn = 0x7f807f80u, pmx = 0xff80ff80u;      // packed bf16x2 accumulators

    for (; tile < n_tiles; tile += gridDim.x) {
        const int64_t item = tile * kThreads + threadIdx.x;
        if (item < a.n_items) {
            if (tile != blockIdx.x) issue_loads(item);
#pragma unroll
            for (int s = 0; s < kMaxSumSources; ++s) {
                if (s < a.n_src) {
                    // one ADD of source s onto the accumulator words: the byte-permute form of dequantize_common.cuh where the zero
                    // point allows it (3 instructions per element), the general per-element steps otherwise -- bit-identical
                    uint32_t next[NWO];
                    dequant_words<BITS, OUT_DT, OP_ADD, NWI, NWO>(wi[s], acc, s_src[s], next);
#pragma unroll
                    for (int k = 0; k < NWO; ++k) acc[k] = next[k];     // (a bf16 accumulator is rounded here, after every source)
                }
            }
#pragma unroll
            for (int k = 0; k < NWO; ++k) {
                if constexpr (OUT_DT == DT_F32) {
                    mn = fminf(mn, __uint_as_float(acc[k]));
                    mx = fmaxf(mx, __uint_as_float(acc[k]));
                } else {
                    pmn = min_bf16x2(pmn, acc[k]);
                    pmx = max_bf16x2(pmx, acc[k]);
                }
            }
            store_words<NWO, A32, true>(out + item * 64, acc);
        }
    }

    if (blockIdx.x == gridDim.x - 1) {
        // ragged head / tail bytes: source after source through the byte-granular ADD (keeps the reference's u2->f32 tail quirk)
        const int64_t total = (a.numel + PER - 1) / PER;
        auto ragged = [&](int64_t b) {
            for (int s = 0; s < a.n_src; ++s) dequant_one_byte<BITS, OUT_DT, OP_ADD>(s_src[s], b);
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
        };
        for (int64_t b = threadIdx.x; b < a.head_bytes; b += kThreads) ragged(b);
        for (int64_t b = a.head_bytes + a.n_items * IB + threadIdx.x; b < total; b += kThreads) ragged(b);
    }
    if constexpr (OUT_DT == DT_BF16) {
        mn = fminf(mn, fminf(bf16_lo(pmn), bf16_hi(pmn)));
        mx = fmaxf(mx, fmaxf(bf16_lo(pmx), bf16_hi(pmx)));
    }
    cta_reduce_tail(mn, mx, a.tail);
}

using SumKernel = void (*)(const SumArgs);

template <int BITS, int OUT_DT>
static int launch_sum_cell(const void* const* ins, const QuantParams* const* dPs, int n_src, void* out, int64_t numel, const LaunchCfg& cfg,
                           const MinMaxScratch& scratch, const ReduceOut& ro) {
    constexpr int PER = 8 / BITS;
    constexpr int V = OUT_DT == DT_F32 ? 16 : 32;
    constexpr int OSZ = OUT_DT == DT_F32 ? 4 : 2;
    constexpr int IB = V * BITS / 8;
    SumArgs a{};
    for (int s = 0; s < n_src; ++s) {
        a.in[s] = static_cast<const uint8_t*>(ins[s]);
        a.dP[s] = dPs[s];
        // one head serves every source only when they share the byte phase of the first one
        if ((reinterpret_cast<uintptr_t>(ins[s]) & 31u) != (reinterpret_cast<uintptr_t>(ins[0]) & 31u)) return 0;
    }
    a.n_src = n_src;
    a.out = static_cast<char*>(out);
    a.numel = numel;
    const int64_t full_bytes = numel / PER;
    bool vec = false, a32 = false;
    for (int pass = 0; pass < 2 && !vec; ++pass) {      // same head search as dequantize.cu: 32-byte aligned streams, else 16
        const uintptr_t oalign = pass == 0 ? 32 : 16;
        const uintptr_t ialign = (pass == 0 || IB < 16) ? IB : 16;
        for (int64_t h = 0; h < 64 && h <= full_bytes; ++h) {
            const uintptr_t o = reinterpret_cast<uintptr_t>(out) + static_cast<uintptr_t>(h) * PER * OSZ;
            const uintptr_t i = reinterpret_cast<uintptr_t>(ins[0]) + static_cast<uintptr_t>(h);
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
    if (!vec) return 0;
    a.tail = make_reduce_tail(scratch, ro);
    // Several tiles per CTA (grid stride): the ticketed tail -- two barriers, an atomic, a fence -- and the parameter staging are paid
    // once per CTA, and with one 45 KB tile per CTA they cost as much as the tile (measured, 7 x u8 -> f32, 33.5 M elements:
    // 1 / 2 / 4 tiles per CTA = 91.7 / 83.4 / 81.8 us, profiles/r2_sum_kernel_sweep.txt); short tensors keep one tile per CTA
    // so that every SM still gets its share.
    constexpr int MINB = sum_min_blocks(BITS, OUT_DT);
    const int64_t n_tiles = (a.n_items + kThreads - 1) / kThreads;
    int64_t per_cta = n_tiles / (static_cast<int64_t>(cfg.sm_count) * MINB * 2);
    per_cta = per_cta < 1 ? 1 : (per_cta > 4 ? 4 : per_cta);
    int64_t grid = (n_tiles + per_cta - 1) / per_cta;
    if (grid > scratch.max_blocks) grid = scratch.max_blocks;      // one partial per CTA; further tiles by grid stride
    const SumKernel fn = a32 ? dequant_sum_kernel<BITS, OUT_DT, true, MINB> : dequant_sum_kernel<BITS, OUT_DT, false, MINB>;
    launch_kernel(fn, static_cast<unsigned>(grid), kThreads, 0, cfg.stream, a);
    PQ_CUDA_CHECK(cudaGetLastError());
    return 1;
}

int launch_dequantize_sum_minmax(const void* const* ins, const QuantParams* const* dPs, int n_src, int dt_in, void* out, int dt_out,
                                 int64_t numel, const LaunchCfg& cfg, const MinMaxScratch& scratch, const ReduceOut& ro) {
    pq_assert(numel > 0, "sum of empty tensors");
    pq_assert(n_src >= 1 && n_src <= kMaxSumSources, "between 1 and %d sources per launch (got %d)", kMaxSumSources, n_src);
    // one source: that is the fused dequantize-ADD + min/max stream kernel (two items per thread, 58 registers)
    if (n_src == 1) return launch_dequantize_add_minmax(ins[0], dt_in, out, dt_out, numel, make_params(1.0f, 0, 0.0f, dt_in), cfg, dPs[0], scratch, ro);
    int n = 0;
    if (dt_out == DT_F32) {
        switch (dt_in) {
            case DT_U8: n = launch_sum_cell<8, DT_F32>(ins, dPs, n_src, out, numel, cfg, scratch, ro); break;
            case DT_U4: n = launch_sum_cell<4, DT_F32>(ins, dPs, n_src, out, numel, cfg, scratch, ro); break;
            default:    n = launch_sum_cell<2, DT_F32>(ins, dPs, n_src, out, numel, cfg, scratch, ro); break;
        }
    } else {
        switch (dt_in) {
            case DT_U8: n = launch_sum_cell<8, DT_BF16>(ins, dPs, n_src, out, numel, cfg, scratch, ro); break;
            case DT_U4: n = launch_sum_cell<4, DT_BF16>(ins, dPs, n_src, out, numel, cfg, scratch, ro); break;
            default:    n = launch_sum_cell<2, DT_BF16>(ins, dPs, n_src, out, numel, cfg, scratch, ro); break;
        }
    }
    if (n) return n;
    // buffers that cannot share one vector alignment: the separate launches, same results
    const QuantParams unit = make_params(1.0f, 0, 0.0f, dt_in);
    for (int s = 0; s + 1 < n_src; ++s) n += launch_dequantize(ins[s], dt_in, out, dt_out, numel, unit, OP_ADD, cfg, dPs[s]);
    return n + launch_dequantize_add_minmax(ins[n_src - 1], dt_in, out, dt_out, numel, unit, cfg, dPs[n_src - 1], scratch, ro);
}

}  // namespace pq
