// quantize_common.cuh -- pieces shared by the direct (quantize.cu) and TMA (quantize_tma.cu) quantize kernels.
#pragma once

#include "pq_kernels.h"

namespace pq {

struct QuantArgs {
    const char* in;          // first element
    uint8_t*    out;         // first packed byte
    int64_t     numel;
    int64_t     head_bytes;  // output bytes in front of the vectorised region
    int64_t     n_items;     // full 16-byte output items in the vectorised region
    QuantParams P;
    const QuantParams* dP;   // not null: the parameters were produced on the device (params_kernel) and are read from there
    unsigned long long* sched;   // TMA kernels: {next tile, finished CTAs} (LaunchCfg::sched)
    // direct stream kernel only, precomputed on the host so that the per-thread prologue stays short (2 vectors per thread
    // leave little to amortise it over): the vectorised region and how many CTAs own a full tile
    const char* in_body;     // in + head_bytes * elements-per-byte * sizeof(input element)
    uint8_t*    out_body;    // out + head_bytes
    int64_t     n_vecs;      // 32-byte input vectors in the region
    uint32_t    n_full_tiles;
    uint32_t    reverse;     // != 0: CTA b takes tile gridDim.x - 1 - b (LaunchCfg::reverse)
    // bf16 -> 2-bit threshold kernel (quantize.cu): q(x) = #{k : x >= thr[k]} while |x| <= thr_xlim
    uint32_t    thr[3];      // bf16 threshold k in both halves of the word
    uint32_t    thr_xlim;    // largest |x| (bf16 bits) for which the three compares are the exact result
    PhiloxKey   sr_key;      // STEP_SRPE: key of this call
    int64_t     sr_base;     // STEP_SRPE: index of element 0 of this launch in the caller's tensor (multiple of 8; host-pointer chunks)
};

// STEP_SRPE: the 4 Philox words that hold the random bits of elements [8g, 8g + 8) of the caller's tensor
__device__ __forceinline__ void srpe_words(const QuantArgs& a, int64_t g, uint32_t (&r)[4]) {
    philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(static_cast<uint64_t>(g) >> 32), 0u, 0u, a.sr_key, r);
}

// Parameters computed by an earlier kernel on the stream replace the by-value ones (the per-call stochastic
// threshold always comes from the host).  Call after pdl_wait().
// Returns false when the block is flagged as failed: the kernel then returns without touching its output.
__device__ __forceinline__ bool load_device_params(QuantArgs& a) {
    if (a.dP) {
        if (device_params_failed(a.dP)) return false;
        const float xi = a.P.xi;
        a.P = *a.dP;
        a.P.xi = xi;
    }
    return true;
}

template <int IN_DT>
__device__ __forceinline__ float load_elem(const char* in, int64_t e) {
    if constexpr (IN_DT == DT_F32) return __ldg(reinterpret_cast<const float*>(in) + e);
    else return bf16_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(in) + e));
}

// One packed output byte from up to 8/BITS elements; elements past numel leave zero bits
// (reference quantize.inl:67-70, :90-98).
template <int IN_DT, int BITS, int STEP>
__device__ __forceinline__ void quant_one_byte(const QuantArgs& a, int64_t b) {
    constexpr int PER = 8 / BITS;
    constexpr int QMAX = (1 << BITS) - 1;
    uint32_t byte = 0, present = 0;
    [[maybe_unused]] uint32_t rnd[4];
    if constexpr (STEP == STEP_SRPE) srpe_words(a, (a.sr_base + b * PER) >> 3, rnd);      // a byte never straddles a group of 8
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int64_t e = b * PER + k;
        if (e < a.numel) {
            uint32_t q;
            if constexpr (STEP == STEP_SRPE) {
                const int j = static_cast<int>((a.sr_base + e) & 7);
                q = static_cast<uint32_t>(quant_step_srpe(load_elem<IN_DT>(a.in, e), a.P, QMAX, srpe_one_plus_u(rnd[j >> 1], j & 1)));
            } else {
                q = static_cast<uint32_t>(quant_step<STEP>(load_elem<IN_DT>(a.in, e), a.P, QMAX));
            }
            byte |= q << (k * BITS);
            present |= static_cast<uint32_t>(QMAX) << (k * BITS);
        }
    }
    a.out[b] = static_cast<uint8_t>(byte ^ (a.P.sign_xor & present));     // signed dtypes: sign bit of the fields that exist
}

}  // namespace pq
