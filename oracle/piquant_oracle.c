/*
 * piquant_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see piquant_oracle.h).
 *
 * Scalar C11 restatement of pi-quant's per-element arithmetic.  No SIMD, no threads: every
 * function states, lane by lane, what the reference's AVX-512 bodies and scalar heads/tails
 * compute.  Compile with -ffp-contract=off: every rounding below is intentional.
 * Citations are file:line under /root/reference.
 */
#include "piquant_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* helpers                                                                                    */
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* bfp16_t -> fp32_t: piquant.hpp:95 */
float orc_bf16_to_f32(uint16_t b) { return bits_f32((uint32_t)b << 16); }

/* fp32_t -> bfp16_t: piquant.hpp:86-90 (round-to-nearest-even, NaN forced quiet).  The SIMD
 * bodies' cvt_ps_to_bf16 (kernels_specialized.inl:15-32) rounds identically for every non-NaN. */
uint16_t orc_f32_to_bf16(float x) {
    uint32_t u = f32_bits(x);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 64u);
    return (uint16_t)((u + (0x7fffu + ((u >> 16) & 1u))) >> 16);
}

/* x86 CVTTPS2DQ / CVTTSS2SI: truncate; NaN and out-of-range give the "integer indefinite" value.
 * (_mm512_cvttps_epi32 at kernels_specialized.inl:70; static_cast<int32_t>(float) in the tails.) */
static inline int32_t x86_cvtt_i32(float a) {
    if (a >= -2147483648.0f && a < 2147483648.0f) return (int32_t)a;
    return INT32_MIN;
}
static inline int64_t x86_cvtt_i64(float a) {
    if (a >= -9223372036854775808.0f && a < 9223372036854775808.0f) return (int64_t)a;
    return INT64_MIN;
}
static inline int64_t x86_cvttsd_i64(double a) {
    if (a >= -9223372036854775808.0 && a < 9223372036854775808.0) return (int64_t)a;
    return INT64_MIN;
}
/* two's-complement wrap, as vpaddd / add do */
static inline int32_t wrap_add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_sub32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int64_t wrap_add64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t wrap_sub64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }

static inline int is_quant(int dt) { return dt == ORC_UINT2 || dt == ORC_UINT4 || dt == ORC_UINT8; }
/* signed extension types (piquant_oracle.h): defined through the unsigned functions below */
static inline int is_signed_quant(int dt) { return dt == ORC_INT2 || dt == ORC_INT4 || dt == ORC_INT8; }
static inline int unsigned_view(int dt) { return dt == ORC_INT2 ? ORC_UINT2 : dt == ORC_INT4 ? ORC_UINT4 : dt == ORC_INT8 ? ORC_UINT8 : dt; }
static inline int is_float(int dt) { return dt == ORC_F32 || dt == ORC_BF16; }
static inline int bits_of(int dt) {
    switch (dt) { case ORC_F32: return 32; case ORC_BF16: return 16; case ORC_UINT2: return 2;
                  case ORC_UINT4: return 4; case ORC_UINT8: return 8; default: return 0; }
}
static inline int64_t qmax_of(int dt) { return (1ll << bits_of(dt)) - 1; } /* dtype_limits, piquant.hpp:175-186 */

/* intN <-> uintN: flip the sign bit of the fields of elements [0, numel) of a packed buffer (offset binary <-> two's complement) */
static void flip_sign_bits(uint8_t* q, int bits, int64_t numel) {
    const int per = 8 / bits;
    for (int64_t e = 0; e < numel; ++e) q[e / per] ^= (uint8_t)(1u << ((int)(e % per) * bits + bits - 1));
}

size_t orc_packed_bytes(int dtype, size_t numel) {           /* piquant_internal.hpp:41-44 */
    dtype = unsigned_view(dtype);
    size_t per_byte = 8u / (size_t)bits_of(dtype);
    return (numel + per_byte - 1) / per_byte;
}
size_t orc_storage_bytes(int dtype, size_t numel) {
    dtype = unsigned_view(dtype);
    return is_quant(dtype) ? orc_packed_bytes(dtype, numel) : numel * (size_t)(bits_of(dtype) / 8);
}

static inline float load_fp(const void* p, int dt, int64_t i) {
    return dt == ORC_F32 ? ((const float*)p)[i] : orc_bf16_to_f32(((const uint16_t*)p)[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* quantize: per-element steps                                                                */
/* ------------------------------------------------------------------------------------------ */

/* One lane of the SIMD bodies, e.g. kernels_specialized.inl:62-77 (f32->u8), :344-359 (f32->u4),
 * :686-694 (bf16->u2): p = x*inv; a = p + (p >= 0 ? .5 : -.5) [ordered compare: NaN -> -.5];
 * t = cvtt(a); q = t + zp (wrapping); clamp to [0, qmax]. */
int32_t orc_quant_step_body(float x, float inv_scale, int32_t zp32, int32_t qmax) {
    float p = x * inv_scale;
    float a = p + ((p >= 0.0f) ? 0.5f : -0.5f);
    int32_t q = wrap_add32(x86_cvtt_i32(a), zp32);
    return q < 0 ? 0 : (q > qmax ? qmax : q);
}

/* Scalar heads/tails of the specialised kernels, e.g. kernels_specialized.inl:52-56,178-182,
 * 468-472, 704-710: std::round, int32 arithmetic. */
int32_t orc_quant_step_tail32(float x, float inv_scale, int32_t zp32, int32_t qmax) {
    float r = roundf(x * inv_scale);
    int32_t q = wrap_add32(x86_cvtt_i32(r), zp32);
    return q < 0 ? 0 : (q > qmax ? qmax : q);
}

/* quant_step_scalar_nearest, quantize.inl:21-26: std::round, int64 arithmetic. */
int64_t orc_quant_step_scalar_nearest(float x, float inv_scale, int64_t zp, int64_t qmax) {
    float r = roundf(x * inv_scale);
    int64_t q = wrap_add64(x86_cvtt_i64(r), zp);
    return q < 0 ? 0 : (q > qmax ? qmax : q);
}

/* quant_step_scalar_stochastic, quantize.inl:8-19.  xi is ONE threshold per call
 * (piquant.cpp:199-201), not per element. */
int64_t orc_quant_step_scalar_stochastic(float x, float inv_scale, int64_t zp, int64_t qmax, float xi) {
    float rnd = x * inv_scale;
    float tr = truncf(rnd);
    float dec = fabsf(rnd - tr);
    float adj = (xi < dec) ? 1.0f : 0.0f;
    if (rnd < 0.0f) adj = -1.0f * adj;
    rnd = tr + adj;
    int64_t q = wrap_add64(x86_cvtt_i64(rnd), zp);
    return q < 0 ? 0 : (q > qmax ? qmax : q);
}

/* Which formula a quantize cell uses for one element. */
enum { STEP_BODY, STEP_TAIL32, STEP_SCALAR64 };

static inline uint8_t quant_elem(float x, float inv, int64_t zp, int64_t qmax, int mode, float xi, int step) {
    if (mode == ORC_STOCHASTIC) return (uint8_t)orc_quant_step_scalar_stochastic(x, inv, zp, qmax, xi);
    switch (step) {
        case STEP_BODY:   return (uint8_t)orc_quant_step_body(x, inv, (int32_t)zp, (int32_t)qmax);
        case STEP_TAIL32: return (uint8_t)orc_quant_step_tail32(x, inv, (int32_t)zp, (int32_t)qmax);
        default:          return (uint8_t)orc_quant_step_scalar_nearest(x, inv, zp, qmax);
    }
}

/* Does the reference have a hand-vectorised nearest kernel for this cell? (quantize.inl:110-130).
 * f32->u2 has none and falls through to the generic int64 scalar step. */
static inline int has_simd_quant(int dt_in, int dt_out) {
    return !(dt_in == ORC_F32 && dt_out == ORC_UINT2);
}

/* Elements per AVX-512 main-loop iteration of each specialised quantize kernel. */
static inline int64_t quant_body_width(int dt_out) { return dt_out == ORC_UINT8 ? 64 : 16; }

/*
 * One kernel invocation over `n` elements starting at element `e0` of (in, out); `out` is the
 * packed byte stream of the whole tensor and e0 is a multiple of the pack width.
 * step_of(i) decides body/tail per element.
 */
static void quant_range(const void* in, int dt_in, uint8_t* out, int dt_out, int64_t e0, int64_t n,
                        float scale, int64_t zp, int mode, float xi, int semantics) {
    const float inv = 1.0f / scale;                      /* kernels_specialized.inl:42, quantize.inl:134 */
    const int64_t qmax = qmax_of(dt_out);
    const int bits = bits_of(dt_out);
    const int per = 8 / bits;
    const int simd = (mode == ORC_NEAREST) && has_simd_quant(dt_in, dt_out);

    /* [body_lo, body_hi) in partition-local element indices uses the SIMD-lane formula */
    int64_t body_lo = 0, body_hi = 0;
    if (simd) {
        if (semantics == ORC_SEM_BODY) { body_lo = 0; body_hi = n; }
        else {
            int64_t i = 0;
            if (dt_in == ORC_F32 && dt_out == ORC_UINT8) /* scalar head until o is 16 B aligned, :52-56 */
                while (i < n && (((uintptr_t)(out + e0 + i)) & 15u) != 0) ++i;
            const int64_t w = quant_body_width(dt_out);
            body_lo = i;
            while (i + (w - 1) < n) i += w;
            body_hi = i;
        }
    }
    uint8_t* o = out + (e0 * bits) / 8;
    for (int64_t i = 0; i < n; i += per) {
        uint8_t byte = 0;
        for (int k = 0; k < per && i + k < n; ++k) {
            const int64_t e = i + k;
            int step = simd ? ((e >= body_lo && e < body_hi) ? STEP_BODY : STEP_TAIL32) : STEP_SCALAR64;
            uint8_t q = quant_elem(load_fp(in, dt_in, e0 + e), inv, zp, qmax, mode, xi, step);
            byte |= (uint8_t)((q & (uint8_t)qmax) << (k * bits)); /* low element in low bits: quantize.inl:36-50 */
        }
        o[i / per] = byte;                               /* missing tail elements stay 0: quantize.inl:67-70,90-98 */
    }
}

/* job_entry's partition of [0,n) over tc threads, piquant.cpp:132-158 */
static int partition(int64_t n, int64_t t, int64_t tc, int64_t pack_elems, int64_t* begin, int64_t* count) {
    int64_t raw_begin = n * t / tc, raw_end = n * (t + 1) / tc;
    int64_t b = pack_elems == 1 ? raw_begin : raw_begin - raw_begin % pack_elems;
    int64_t e = (t + 1 == tc || pack_elems == 1) ? raw_end : raw_end - raw_end % pack_elems;
    if (b >= e) return 0;
    *begin = b; *count = e - b;
    return 1;
}

int orc_quantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel,
                 float scale, int64_t zero_point, int round_mode, float rnd_threshold,
                 int semantics, int nthreads) {
    if (is_signed_quant(dt_out)) {      /* extension: intN = uintN with zero point + 2^(N-1), sign bits flipped */
        const int bits = bits_of(unsigned_view(dt_out));
        int rc = orc_quantize(in, dt_in, out, unsigned_view(dt_out), numel, scale, wrap_add64(zero_point, 1ll << (bits - 1)),
                              round_mode, rnd_threshold, semantics, nthreads);
        if (rc == 0) flip_sign_bits((uint8_t*)out, bits, numel);
        return rc;
    }
    if (!is_float(dt_in) || !is_quant(dt_out)) return -1;          /* piquant.cpp:288-289 */
    if (semantics == ORC_SEM_BODY || nthreads < 1) {
        quant_range(in, dt_in, (uint8_t*)out, dt_out, 0, numel, scale, zero_point, round_mode, rnd_threshold, ORC_SEM_BODY);
        return 0;
    }
    const int64_t pack = 8 / bits_of(dt_out);
    for (int64_t t = 0; t < nthreads; ++t) {
        int64_t b, c;
        if (partition(numel, t, nthreads, pack, &b, &c))
            quant_range(in, dt_in, (uint8_t*)out, dt_out, b, c, scale, zero_point, round_mode, rnd_threshold, ORC_SEM_REF);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* extension: per-element stochastic rounding (spec in piquant_oracle.h)                       */
/* ------------------------------------------------------------------------------------------ */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* one element: k = the 16 random bits of tensor index j, q = clamp(floor(p + (k + 1/2) / 65536) + zp, 0, qmax) */
static int64_t sr_quant_value(float x, float inv, int64_t zero_point, int64_t qmax, const uint32_t key[2], int64_t j) {
    const uint32_t ctr[4] = {(uint32_t)((uint64_t)j >> 3), (uint32_t)((uint64_t)j >> 35), 0u, 0u};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    const uint32_t k = (r[(j & 7) >> 1] >> (16 * (int)(j & 1))) & 0xffffu;
    const double u = ((double)k + 0.5) / 65536.0;
    const float p = x * inv;
    int64_t t;
    if (fabsf(p) < 8388608.0f) t = (int64_t)floor((double)p + u);   /* exact: < 53 significant bits */
    else t = x86_cvtt_i64(p);
    int64_t q = wrap_add64(t, zero_point);
    return q < 0 ? 0 : (q > qmax ? qmax : q);
}

int orc_quantize_sr(const void* in, int dt_in, void* out, int dt_out, int64_t numel,
                    float scale, int64_t zero_point, uint64_t key, int64_t base) {
    if (is_signed_quant(dt_out)) {
        const int bits = bits_of(unsigned_view(dt_out));
        int rc = orc_quantize_sr(in, dt_in, out, unsigned_view(dt_out), numel, scale, wrap_add64(zero_point, 1ll << (bits - 1)), key, base);
        if (rc == 0) flip_sign_bits((uint8_t*)out, bits, numel);
        return rc;
    }
    if (!is_float(dt_in) || !is_quant(dt_out)) return -1;
    const int bits = bits_of(dt_out), per = 8 / bits;
    const int64_t qmax = qmax_of(dt_out);
    const float inv = 1.0f / scale;
    const uint32_t k2[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    uint8_t* o = (uint8_t*)out;
    memset(o, 0, orc_packed_bytes(dt_out, (size_t)numel));
    for (int64_t i = 0; i < numel; ++i) {
        const int64_t q = sr_quant_value(load_fp(in, dt_in, i), inv, zero_point, qmax, k2, base + i);
        o[i / per] |= (uint8_t)(q << ((int)(i % per) * bits));
    }
    return 0;
}

/* requantize with the per-element rule: the quantized value as above, then the generic dequant_step of orc_requantize */
int orc_requantize_sr(const void* in, int dt_inout, void* out, int dt_quant, int64_t numel,
                      float scale, int64_t zero_point, uint64_t key, int64_t base, int reduce_op, int fma_add) {
    if (is_signed_quant(dt_quant))
        return orc_requantize_sr(in, dt_inout, out, unsigned_view(dt_quant), numel, scale,
                                 wrap_add64(zero_point, 1ll << (bits_of(unsigned_view(dt_quant)) - 1)), key, base, reduce_op, fma_add);
    if (!is_float(dt_inout) || !is_quant(dt_quant)) return -1;
    const float inv = 1.0f / scale;
    const int64_t qmax = qmax_of(dt_quant);
    const uint32_t k2[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    for (int64_t i = 0; i < numel; ++i) {
        const int64_t q = sr_quant_value(load_fp(in, dt_inout, i), inv, zero_point, qmax, k2, base + i);
        const int64_t d = wrap_sub64(q, zero_point);
        if (dt_inout == ORC_F32) {
            float* o = (float*)out + i;
            if (reduce_op == ORC_ADD) *o = fma_add ? fmaf((float)d, scale, *o) : *o + (float)d * scale;
            else *o = (float)d * scale;
        } else {
            uint16_t* o = (uint16_t*)out + i;
            float a = orc_bf16_to_f32(orc_f32_to_bf16((float)d));
            float sc = orc_bf16_to_f32(orc_f32_to_bf16(scale));
            uint16_t r = orc_f32_to_bf16(a * sc);
            if (reduce_op == ORC_ADD) r = orc_f32_to_bf16(orc_bf16_to_f32(*o) + orc_bf16_to_f32(r));
            *o = r;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* dequantize                                                                                 */
/* ------------------------------------------------------------------------------------------ */

static int g_fma_contract = 1;
/* GCC contracts `mul + add(o)` into one fma in every FMA-enabled translation unit of the
 * reference (-ffp-contract=fast is the GNU-mode default); 1 reproduces that build, 0 the
 * source-level two-rounding semantics. */
void orc_set_fma_contract(int on) { g_fma_contract = on; }

static inline float mul_add(float a, float b, float c) {
    return g_fma_contract ? fmaf(a, b, c) : (a * b) + c;
}

static inline uint8_t unpack(const uint8_t* x, int bits, int64_t e) {
    const int per = 8 / bits;
    return (uint8_t)((x[e / per] >> ((e % per) * bits)) & ((1u << bits) - 1u));
}

/* Elements per AVX-512 main-loop iteration of the specialised dequantize kernels; 0 = none. */
static inline int64_t dequant_body_width(int dt_in, int dt_out) {
    if (dt_in == ORC_UINT8) return 64;                              /* :741, :946 */
    if (dt_in == ORC_UINT4) return 128;                             /* :1026, :1230 */
    if (dt_in == ORC_UINT2 && dt_out == ORC_BF16) return 256;       /* :1379 */
    return 0;                                                       /* u2->f32: generic, dequantize.inl:42-87 */
}

static void dequant_range(const uint8_t* in, int dt_in, void* out, int dt_out, int64_t e0, int64_t n,
                          float scale, int64_t zp, int op, int semantics) {
    const int bits = bits_of(dt_in);
    const int32_t zp32 = (int32_t)zp;                               /* dequantize.inl:97,101,105,109,113 */
    int64_t w = dequant_body_width(dt_in, dt_out);
    int64_t body_hi = 0;
    if (w) {
        if (semantics == ORC_SEM_BODY) body_hi = n;
        else { int64_t i = 0; while (i + (w - 1) < n) i += w; body_hi = i; }
    }
    const float bias = -(float)zp32 * scale;                        /* :1204, :1318 */
    for (int64_t i = 0; i < n; ++i) {
        const int64_t e = e0 + i;
        const int32_t q = unpack(in, bits, e);
        const int in_body = i < body_hi;
        if (dt_out == ORC_F32) {
            float* o = (float*)out + e;
            if (w) {    /* u8/u4 -> f32: body :748-761 / :1030-1052 and tails :920-924 / :1168-1187 agree */
                float d = (float)wrap_sub32(q, zp32);
                *o = (op == ORC_ADD) ? mul_add(d, scale, *o) : d * scale;
            } else {    /* generic dequant_step, dequantize.inl:8-11: int64 difference */
                float d = (float)wrap_sub64((int64_t)q, zp);
                /* reference bug kept: the 1-3 element tail of dequant_uint2 always SETs,
                 * dequantize.inl:72-86 */
                int tail = (i >= n - (n & 3));
                *o = (op == ORC_ADD && !tail) ? mul_add(d, scale, *o) : d * scale;
            }
        } else {
            uint16_t* o = (uint16_t*)out + e;
            if (in_body) {
                float f;
                if (dt_in == ORC_UINT8) {                            /* :946-966 */
                    float d = (float)wrap_sub32(q, zp32);
                    f = (op == ORC_ADD) ? mul_add(d, scale, orc_bf16_to_f32(*o)) : d * scale;
                } else {                                             /* u4 :1236-1262, u2 :1361-1370 */
                    f = fmaf((float)q, scale, bias);
                    if (op == ORC_ADD) f = f + orc_bf16_to_f32(*o);
                }
                *o = orc_f32_to_bf16(f);
            } else {
                /* scalar tails: dq in f32, then bfp16_t arithmetic (piquant.hpp:97-103) */
                float dq;
                if (dt_in == ORC_UINT2) dq = ((float)q - (float)zp32) * scale;        /* :1387-1389 */
                else dq = (float)wrap_sub32(q, zp32) * scale;                          /* :978, :1290-1292 */
                uint16_t r = orc_f32_to_bf16(dq);
                if (op == ORC_ADD) r = orc_f32_to_bf16(orc_bf16_to_f32(*o) + orc_bf16_to_f32(r));
                *o = r;
            }
        }
    }
}

int orc_dequantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel,
                   float scale, int64_t zero_point, int reduce_op, int semantics, int nthreads) {
    if (is_signed_quant(dt_in)) {       /* extension: flip the sign bits of a copy, dequantize as uintN with zero point + 2^(N-1) */
        const int bits = bits_of(unsigned_view(dt_in));
        const size_t nb = orc_packed_bytes(dt_in, (size_t)numel);
        uint8_t* tmp = (uint8_t*)malloc(nb ? nb : 1);
        if (!tmp) return -1;
        memcpy(tmp, in, nb);
        flip_sign_bits(tmp, bits, numel);
        int rc = orc_dequantize(tmp, unsigned_view(dt_in), out, dt_out, numel, scale, wrap_add64(zero_point, 1ll << (bits - 1)),
                                reduce_op, semantics, nthreads);
        free(tmp);
        return rc;
    }
    if (!is_quant(dt_in) || !is_float(dt_out)) return -1;           /* piquant.cpp:321-322 */
    if (semantics == ORC_SEM_BODY || nthreads < 1) {
        dequant_range((const uint8_t*)in, dt_in, out, dt_out, 0, numel, scale, zero_point, reduce_op, ORC_SEM_BODY);
        return 0;
    }
    const int64_t pack = 8 / bits_of(dt_in);
    for (int64_t t = 0; t < nthreads; ++t) {
        int64_t b, c;
        if (partition(numel, t, nthreads, pack, &b, &c))
            dequant_range((const uint8_t*)in, dt_in, out, dt_out, b, c, scale, zero_point, reduce_op, ORC_SEM_REF);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* requantize (quantize -> dequantize, unpacked), kernels.inl:30-52                           */
/* ------------------------------------------------------------------------------------------ */

int orc_requantize(const void* in, int dt_inout, void* out, int dt_quant, int64_t numel,
                   float scale, int64_t zero_point, int round_mode, float rnd_threshold,
                   int reduce_op, int fma_add) {
    if (is_signed_quant(dt_quant))      /* extension: clamp(q + zp, -2^(N-1), 2^(N-1)-1) - zp == clamp(q + zp', 0, 2^N-1) - zp' with zp' = zp + 2^(N-1) */
        return orc_requantize(in, dt_inout, out, unsigned_view(dt_quant), numel, scale,
                              wrap_add64(zero_point, 1ll << (bits_of(unsigned_view(dt_quant)) - 1)), round_mode, rnd_threshold, reduce_op, fma_add);
    if (!is_float(dt_inout) || !is_quant(dt_quant)) return -1;      /* piquant.cpp:353-354 */
    const float inv = 1.0f / scale;
    const int64_t qmax = qmax_of(dt_quant);
    for (int64_t i = 0; i < numel; ++i) {
        float x = load_fp(in, dt_inout, i);
        int64_t q = (round_mode == ORC_STOCHASTIC)
            ? orc_quant_step_scalar_stochastic(x, inv, zero_point, qmax, rnd_threshold)
            : orc_quant_step_scalar_nearest(x, inv, zero_point, qmax);
        int64_t d = wrap_sub64(q, zero_point);
        if (dt_inout == ORC_F32) {                                  /* dequant_step<.., fp32_t>, dequantize.inl:8-11 */
            float* o = (float*)out + i;
            if (reduce_op == ORC_ADD) *o = fma_add ? fmaf((float)d, scale, *o) : *o + (float)d * scale;
            else *o = (float)d * scale;
        } else {                                                    /* bfp16_t arithmetic: both operands rounded to bf16 first */
            uint16_t* o = (uint16_t*)out + i;
            float a = orc_bf16_to_f32(orc_f32_to_bf16((float)d));
            float s = orc_bf16_to_f32(orc_f32_to_bf16(scale));
            uint16_t r = orc_f32_to_bf16(a * s);
            if (reduce_op == ORC_ADD) r = orc_f32_to_bf16(orc_bf16_to_f32(*o) + orc_bf16_to_f32(r));
            *o = r;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* min/max and quantization parameters                                                        */
/* ------------------------------------------------------------------------------------------ */

void orc_minmax_f32(const float* x, int64_t n, float out[2]) {      /* kernels_specialized.inl:1418-1516 */
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
    }
    out[0] = mn; out[1] = mx;
}

void orc_minmax_bf16(const uint16_t* x, int64_t n, float out[2]) {  /* :1518-1607 */
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int64_t i = 0; i < n; ++i) {
        float v = orc_bf16_to_f32(x[i]);
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    out[0] = mn; out[1] = mx;
}

int orc_params_from_minmax(double r_min, double r_max, int dt_quant, float* scale, int64_t* zero_point) {
    const int sgn = is_signed_quant(dt_quant);                      /* extension; the reference's is_signed branch, piquant.cpp:218,247-248 */
    dt_quant = unsigned_view(dt_quant);
    if (!is_quant(dt_quant)) return -1;                             /* compute_type_max, piquant.cpp:213-220 */
    const uint64_t type_max = (1ull << (bits_of(dt_quant) - (sgn ? 1 : 0))) - 1;
    const int64_t type_min = sgn ? -(int64_t)type_max - 1 : 0;      /* no signed dtypes at this commit: always 0 there */
    if (r_max == r_min) {                                           /* piquant.cpp:249-252 */
        *scale = 1.0f;
        /* signed: the reference's expression would wrap to INT64_MAX (uint64 + int64, logical shift) in code it never
         * reaches; the extension uses the signed midpoint (type_max + type_min) >> 1 == -1 */
        *zero_point = sgn ? -1 : (int64_t)((type_max + (uint64_t)type_min) >> 1);
        return 0;
    }
    double q_min = (double)type_min, q_max = (double)type_max;
    double s = (r_max - r_min) / (q_max - q_min);                   /* :255 */
    double zp = q_min - r_min / s;                                  /* :256 */
    zp = fmax(fmin((double)x86_cvttsd_i64(round(zp)), q_max), q_min); /* :257 */
    float sf = (float)s;
    if (isnan(sf) || !(sf >= 0.0f)) return -1;                      /* piquant.cpp:373,379 */
    *scale = sf;
    *zero_point = x86_cvttsd_i64(zp);
    return 0;
}

int orc_compute_quant_params_f32(const float* x, int64_t n, int dt_quant, float* scale, int64_t* zero_point) {
    /* empty input: r_min = DBL_MAX, r_max = -DBL_MAX -> negative scale -> abort (piquant.cpp:238-244,373) */
    if (n <= 0) return -1;
    float mm[2];
    orc_minmax_f32(x, n, mm);
    return orc_params_from_minmax((double)mm[0], (double)mm[1], dt_quant, scale, zero_point);
}

int orc_compute_quant_params_bf16(const uint16_t* x, int64_t n, int dt_quant, float* scale, int64_t* zero_point) {
    if (n <= 0) return -1;
    float mm[2];
    orc_minmax_bf16(x, n, mm);
    return orc_params_from_minmax((double)mm[0], (double)mm[1], dt_quant, scale, zero_point);
}
