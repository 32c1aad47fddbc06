// dequantize_common.cuh -- pieces shared by the direct (dequantize.cu) and TMA (dequantize_tma.cu) dequantize kernels.
#pragma once

#include "pq_kernels.h"

namespace pq {

struct DequantArgs {
    const uint8_t* in;          // first packed byte
    char*          out;         // first output element
    int64_t        numel;
    int64_t        head_bytes;  // packed input bytes in front of the vectorised region
    int64_t        n_items;     // direct kernel: full 64-byte output items; TMA kernel: units of 16 packed input bytes
    QuantParams    P;
    float          magic_zp;    // 2^23 + zp32 (exact) when fast != 0
    int32_t        fast;        // |zp| <= 2^22: the byte-permute conversion below is exact
    const QuantParams* dP;      // not null: parameters produced on the device (params_kernel), read from there
    unsigned long long* sched;  // TMA kernels: {next tile, finished CTAs} (LaunchCfg::sched)
};

__host__ __device__ inline void set_dequant_fast(DequantArgs& a, int bits, int out_dt) {
    const QuantParams& P = a.P;
    a.fast = (P.zp32 <= (1 << 22) && P.zp32 >= -(1 << 22) && !(bits == 2 && out_dt == DT_F32 && P.bigzp)) ? 1 : 0;
    a.magic_zp = 8388608.0f + static_cast<float>(a.fast ? P.zp32 : 0);
}

// Parameters computed by an earlier kernel on the stream replace the by-value ones.  Call after pdl_wait().
template <int BITS, int OUT_DT>
__device__ __forceinline__ bool load_device_params(DequantArgs& a) {
    if (a.dP) {
        if (device_params_failed(a.dP)) return false;      // flagged block: no work (see pq_device.cuh)
        a.P = *a.dP;
        set_dequant_fast(a, BITS, OUT_DT);
    }
    return true;
}

// All elements of one packed input byte (elements past numel are skipped).
template <int BITS, int OUT_DT, int OP>
__device__ __forceinline__ void dequant_one_byte(const DequantArgs& a, int64_t b) {
    constexpr int PER = 8 / BITS;
    constexpr uint32_t QMAX = (1u << BITS) - 1u;
    const uint32_t byte = a.in[b] ^ (a.P.sign_xor & 0xffu);     // signed dtypes: two's complement -> offset binary
    // reference quirk kept: the 1-3 element tail of the generic u2->f32 kernel always SETs, even
    // for ADD (src/kernels/dequantize.inl:72-86)
    const int64_t set_from = (BITS == 2 && OUT_DT == DT_F32) ? a.numel - (a.numel & 3) : a.numel;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int64_t e = b * PER + k;
        if (e >= a.numel) break;
        const uint32_t q = (byte >> (k * BITS)) & QMAX;
        if constexpr (OUT_DT == DT_F32) {
            float* o = reinterpret_cast<float*>(a.out) + e;
            if (OP == OP_ADD && e < set_from) *o = dequant_f32<BITS, OP_ADD>(q, *o, a.P);
            else *o = dequant_f32<BITS, OP_SET>(q, 0.0f, a.P);
        } else {
            uint16_t* o = reinterpret_cast<uint16_t*>(a.out) + e;
            const float prev = OP == OP_ADD ? bf16_bits_to_f32(*o) : 0.0f;
            *o = f32_to_bf16_bits(dequant_bf16_pre<BITS, OP>(q, prev, a.P));
        }
    }
}

// float bits of 2^23 + (byte B of x): one PRMT, no int->float conversion
template <int B>
__device__ __forceinline__ float magic_byte(uint32_t x) {
    return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540u + B));
}

// Dequantize NWO output words (NWO*4 bytes: NWO f32 or 2*NWO bf16 elements) from the packed words w[].
// prev[] holds the accumulator words for ADD (ignored for SET).
//
// Fast form (a.fast): a quantized field is turned into the float 2^23 + q by ONE byte-permute into the
// mantissa of 0x4B000000, and float(q - zp) = (2^23 + q) - (2^23 + zp) is exact, so the results are
// bit-identical to the reference's  float(int32(q) - zp) * scale  and  fma(float(q), scale, -float(zp)*scale).
// Nibbles / 2-bit fields are first spread to one field per byte with a shift+mask per word.
template <int BITS, int OUT_DT, int OP, int NWI, int NWO>
__device__ __forceinline__ void dequant_words(const uint32_t (&w_in)[NWI], const uint32_t (&prev)[NWO], const DequantArgs& a,
                                              uint32_t (&o)[NWO]) {
    uint32_t w[NWI];
#pragma unroll
    for (int i = 0; i < NWI; ++i) w[i] = w_in[i] ^ a.P.sign_xor;     // signed dtypes: two's complement -> offset binary
    constexpr int EV = OUT_DT == DT_F32 ? NWO : 2 * NWO;
    constexpr int FPB = 8 / BITS;                       // fields per byte
    constexpr int EPW = 32 / BITS;                      // elements per packed word
    constexpr uint32_t QMAX = (1u << BITS) - 1u;
    float v[EV];
    if (a.fast) {
        constexpr uint32_t MASK = BITS == 8 ? 0xffffffffu : (BITS == 4 ? 0x0f0f0f0fu : 0x03030303u);
        uint32_t s[FPB][NWI];                           // s[k][word]: field k of every byte, one field per byte
#pragma unroll
        for (int k = 0; k < FPB; ++k)
#pragma unroll
            for (int i = 0; i < NWI; ++i) s[k][i] = (w[i] >> (k * BITS)) & MASK;
#pragma unroll
        for (int e = 0; e < EV; ++e) {
            const int word = e / EPW, byte = (e % EPW) / FPB, field = e % FPB;
            float m;
            switch (byte) {
                case 0: m = magic_byte<0>(s[field][word]); break;
                case 1: m = magic_byte<1>(s[field][word]); break;
                case 2: m = magic_byte<2>(s[field][word]); break;
                default: m = magic_byte<3>(s[field][word]); break;
            }
            float pv = 0.0f;
            if constexpr (OP == OP_ADD) {
                if constexpr (OUT_DT == DT_F32) pv = __uint_as_float(prev[e]);
                else pv = (e & 1) ? bf16_hi(prev[e >> 1]) : bf16_lo(prev[e >> 1]);
            }
            if constexpr (OUT_DT == DT_BF16 && BITS != 8) {
                // fma(float(q), scale, -float(zp)*scale) (+ prev)      kernels_specialized.inl:1236-1262, :1361-1370
                const float f = __fmaf_rn(__fsub_rn(m, 8388608.0f), a.P.scale, a.P.bias);
                v[e] = OP == OP_ADD ? __fadd_rn(f, pv) : f;
            } else {
                const float d = __fsub_rn(m, a.magic_zp);           // == float(q - zp), exact
                v[e] = OP == OP_ADD ? __fmaf_rn(d, a.P.scale, pv) : __fmul_rn(d, a.P.scale);
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < EV; ++e) {
            const uint32_t q = (w[(e * BITS) / 32] >> ((e * BITS) % 32)) & QMAX;
            if constexpr (OUT_DT == DT_F32) {
                v[e] = dequant_f32<BITS, OP>(q, OP == OP_ADD ? __uint_as_float(prev[e]) : 0.0f, a.P);
            } else {
                const float pv = OP == OP_ADD ? ((e & 1) ? bf16_hi(prev[e >> 1]) : bf16_lo(prev[e >> 1])) : 0.0f;
                v[e] = dequant_bf16_pre<BITS, OP>(q, pv, a.P);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NWO; ++i) {
        if constexpr (OUT_DT == DT_F32) o[i] = __float_as_uint(v[i]);
        else o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    }
}

}  // namespace pq
