// minmax.cu -- whole-tensor {min, max} of an f32 / bf16 stream in ONE launch (sm_100a).
//
// Replaces find_min_max_f32 / find_min_max_bf16 (src/kernels/kernels_specialized.inl:1418-1607) and
// the per-thread partial gather of compute_quant_config (src/piquant.cpp:222-244).
//
// Read-only HBM stream, 4 (f32) or 2 (bf16) bytes per element: every thread keeps 4 x LDG.256 in
// flight, folds them with FMNMX (f32) or packed HMNMX2.BF16 (two bf16 lanes per instruction),
// then warp shuffles -> one {min,max} per CTA -> the last CTA to finish (atomic ticket) folds the
// per-CTA partials and publishes the result; no second kernel, no host-side gather.
// min/max over non-NaN floats is associative and exact, so the result is bit-identical to the
// reference for any reduction order.  NaNs never win a comparison (the reference's scalar loops use
// `<` / `>`); accumulators start at +-inf and are clamped to +-FLT_MAX at the end, which equals the
// reference's +-FLT_MAX start for every input.
#include "pq_reduce.cuh"

namespace pq {

struct MinMaxArgs {
    const char* x;
    int64_t     numel;
    int64_t     head;       // elements in front of the 32-byte aligned region
    int64_t     n_items;    // 32-byte items in the aligned region
    int64_t     tiles_per_cta;  // each CTA folds this many consecutive tiles
    ReduceTail  tail;           // partials / ticket scratch, where the result goes, optional parameters and rank exchange
};

template <int IN_DT, bool KEEP>
__global__ void __launch_bounds__(kThreads) minmax_kernel(const MinMaxArgs a) {
    constexpr int ISZ = IN_DT == DT_F32 ? 4 : 2;
    constexpr int EPI = 32 / ISZ;     // elements per 32-byte item
    constexpr int U = 4;
    constexpr int64_t TILE = static_cast<int64_t>(kThreads) * U;
    const char* base = a.x + a.head * ISZ;
    const int64_t n_tiles = (a.n_items + TILE - 1) / TILE;

    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    uint32_t pmn = 0x7f807f80u, pmx = 0xff80ff80u;      // packed bf16x2 accumulators (+inf,+inf) / (-inf,-inf)
    pdl_launch_dependents();
    pdl_wait();                                          // also orders this launch after the previous one's use of the scratch

    // CTA b owns tiles [b*C, (b+1)*C): many more CTAs than SMs, dealt by the hardware scheduler, so no SM waits for
    // a slower one (static round-robin over a persistent grid measured ~4 % lower on reads)
    const int64_t tile_begin = static_cast<int64_t>(blockIdx.x) * a.tiles_per_cta;
    const int64_t tile_end = tile_begin + a.tiles_per_cta < n_tiles ? tile_begin + a.tiles_per_cta : n_tiles;
    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
        const int64_t first = tile * TILE + threadIdx.x;
        uint32_t w[U][8];
        if (tile * TILE + TILE <= a.n_items) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if constexpr (KEEP) ldg_keep(base + (first + static_cast<int64_t>(u) * kThreads) * 32, w[u]);
                else ldg_stream(base + (first + static_cast<int64_t>(u) * kThreads) * 32, w[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t item = first + static_cast<int64_t>(u) * kThreads;
                if (item < a.n_items) { if constexpr (KEEP) ldg_keep(base + item * 32, w[u]); else ldg_stream(base + item * 32, w[u]); }
                else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) w[u][k] = IN_DT == DT_F32 ? 0x7fc00000u : 0x7fc07fc0u;   // NaN: never wins
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (IN_DT == DT_F32) {
                    const float v = __uint_as_float(w[u][k]);
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                } else {
                    pmn = min_bf16x2(pmn, w[u][k]);
                    pmx = max_bf16x2(pmx, w[u][k]);
                }
            }
        }
    }
    if constexpr (IN_DT == DT_BF16) {
        mn = fminf(bf16_lo(pmn), bf16_hi(pmn));
        mx = fmaxf(bf16_lo(pmx), bf16_hi(pmx));
    }
    // ragged head / tail elements: last CTA, scalar loads
    if (blockIdx.x == gridDim.x - 1) {
        auto fold = [&](int64_t e) {
            float v;
            if constexpr (IN_DT == DT_F32) v = __ldg(reinterpret_cast<const float*>(a.x) + e);
            else v = bf16_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(a.x) + e));
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        };
        for (int64_t e = threadIdx.x; e < a.head; e += kThreads) fold(e);
        for (int64_t e = a.head + a.n_items * EPI + threadIdx.x; e < a.numel; e += kThreads) fold(e);
    }

    cta_reduce_tail(mn, mx, a.tail);
}

ReduceTail make_reduce_tail(const MinMaxScratch& scratch, const ReduceOut& ro) {
    ReduceTail t{};
    t.partials = scratch.partials;
    t.ticket = scratch.ticket;
    t.result = ro.result;
    t.mapped_result = ro.mapped_result;
    t.meta_out = ro.meta_out;
    t.meta_out2 = ro.meta_out2;
    t.meta_mapped = ro.meta_mapped;
    if (ro.meta_out) {
        pq_assert(dtype_is_quant(ro.dt_quant), "type %s is not a quantization type", dtype_name(ro.dt_quant));
        t.q_bits = dtype_bits(ro.dt_quant);
        t.q_signed = dtype_is_signed_quant(ro.dt_quant) ? 1 : 0;
        t.q_sign_xor = dtype_sign_xor(ro.dt_quant);
    }
    if (ro.px) t.px = *ro.px;
    return t;
}

int launch_minmax(const void* x, int dt, int64_t numel, const MinMaxScratch& scratch, const ReduceOut& ro,
                  const LaunchCfg& cfg, bool keep_in_l2) {
    const int isz = dt == DT_F32 ? 4 : 2;
    const int epi = 32 / isz;
    MinMaxArgs a;
    a.x = static_cast<const char*>(x);
    a.numel = numel;
    a.tail = make_reduce_tail(scratch, ro);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
    int64_t head = static_cast<int64_t>(((32 - (addr & 31u)) & 31u) / isz);   // natural alignment of x is assumed
    if (head > numel) head = numel;
    a.head = head;
    a.n_items = (numel - head) / epi;
    auto fn = dt == DT_F32 ? (keep_in_l2 ? minmax_kernel<DT_F32, true> : minmax_kernel<DT_F32, false>)
                           : (keep_in_l2 ? minmax_kernel<DT_BF16, true> : minmax_kernel<DT_BF16, false>);
    const int64_t tile = static_cast<int64_t>(kThreads) * 4;
    const int64_t n_tiles = (a.n_items + tile - 1) / tile;
    a.tiles_per_cta = (n_tiles + scratch.max_blocks - 1) / scratch.max_blocks;      // one partial per CTA must fit the scratch
    if (a.tiles_per_cta < 1) a.tiles_per_cta = 1;
    int64_t grid = (n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
    if (grid < 1) grid = 1;
    launch_kernel(fn, static_cast<unsigned>(grid), kThreads, 0, cfg.stream, a);
    PQ_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// ---------------------------------------------------------------------------------------------
// (scale, zero_point) on the device
// ---------------------------------------------------------------------------------------------

__global__ void params_kernel(const float* minmax4, int bits, int is_signed, uint32_t sign_xor, DeviceMeta* out, DeviceMeta* mapped_out) {
    pdl_launch_dependents();
    pdl_wait();
    const DeviceMeta m = meta_from_minmax(minmax4[2], minmax4[3], bits, is_signed, sign_xor);
    *out = m;
    if (mapped_out) {
        *mapped_out = m;
        __threadfence_system();
    }
}

// after an all-reduce of {-min, max} (NCCL transport): bring {min, max} in line with the combined pair
__global__ void minmax_publish_kernel(float* r4) {
    pdl_launch_dependents();
    pdl_wait();
    r4[0] = -r4[2];
    r4[1] = r4[3];
}

int launch_minmax_publish(float* minmax4, const LaunchCfg& cfg) {
    launch_kernel(minmax_publish_kernel, 1u, 1u, 0, cfg.stream, minmax4);
    PQ_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// One thread spins until a 4-byte arrival flag (written by a peer's copy engine, possibly through the NVSwitch multicast
// address) is up, lowers it again, and ends: the kernels queued behind it on the stream see what was copied before the flag.
__global__ void wait_flag_kernel(unsigned* flag) {
    pdl_launch_dependents();
    pdl_wait();
    unsigned v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == 0u) __nanosleep(64);
    } while (v == 0u);
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(0u) : "memory");
}

int launch_wait_flag(unsigned* flag, const LaunchCfg& cfg) {
    launch_kernel(wait_flag_kernel, 1u, 1u, 0, cfg.stream, flag);
    PQ_CUDA_CHECK(cudaGetLastError());
    return 1;
}

int launch_params(const float* minmax4, int dt_quant, DeviceMeta* out, DeviceMeta* mapped_out, const LaunchCfg& cfg) {
    launch_kernel(params_kernel, 1u, 1u, 0, cfg.stream, minmax4, dtype_bits(dt_quant), dtype_is_signed_quant(dt_quant) ? 1 : 0,
                  dtype_sign_xor(dt_quant), out, mapped_out);
    PQ_CUDA_CHECK(cudaGetLastError());
    return 1;
}

}  // namespace pq
