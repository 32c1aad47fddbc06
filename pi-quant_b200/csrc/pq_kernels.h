// pq_kernels.h -- host-side launch interface of the sm_100a kernels (internal; the public
// boundary is include/piquant.h + include/piquant_cuda.h).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "pq_device.cuh"

namespace pq {

constexpr int kThreads = 256;   // threads per CTA of every streaming kernel

struct LaunchCfg {
    cudaStream_t stream;
    int          sm_count;   // 148 on B200; sizes the resident grids of the persistent (TMA ring, byte-granular) kernels
    int          variant;    // 0 = auto, 1 = direct LDG/STG kernels, 2 = TMA (cp.async.bulk) ring kernels
    unsigned long long* sched;   // {next tile, finished CTAs}: work counter of the persistent TMA kernels, zero between launches
    uint64_t     sr_key = 0;     // per-element stochastic rounding (mode 2): Philox key of the call
    int64_t      sr_base = 0;    // ... and the index of this launch's element 0 in the caller's tensor (host-pointer chunks)
    bool         reverse = false;   // quantize: deal the tiles from the END of the tensor -- the pass before (min/max, or the
                                    // dequantize-ADD that produced the tensor) touched the end last, so that is what L2 still holds
};

// [[noreturn]] abort with a red message on stderr -- the reference's error convention
// (src/piquant.cpp:88-98): every violated precondition, CUDA errors included, ends the process.
[[noreturn]] void panic(const char* fmt, ...);

#define PQ_CUDA_CHECK(expr)                                                                          \
    do {                                                                                             \
        cudaError_t pq_err__ = (expr);                                                               \
        if (pq_err__ != cudaSuccess)                                                                 \
            ::pq::panic("%s:%d CUDA error %s: %s <- %s", __FILE__, __LINE__, cudaGetErrorName(pq_err__), \
                        cudaGetErrorString(pq_err__), #expr);                                        \
    } while (0)

#define pq_assert(expr, msg, ...)                                                                    \
    do {                                                                                             \
        if (!(expr)) ::pq::panic("%s:%d Assertion failed: " #expr " <- " msg, __FILE__, __LINE__, ##__VA_ARGS__); \
    } while (0)

// Every kernel is launched with programmatic stream serialization (PDL): a launch may begin while the
// previous kernel on the stream drains -- its CTAs get scheduled, set up shared memory / mbarriers and then
// block in griddepcontrol.wait until the predecessor has completed and flushed.  Back-to-back calls on
// short tensors (27 M elements is ~21 us at the HBM roofline) lose no time to launch latency and prologue.
// Kernels of other libraries on the same stream are unaffected: the attribute only relaxes OUR launch, and
// griddepcontrol.wait is a full dependency on whatever ran before.
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*fn)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid, 1, 1);
    lc.blockDim = dim3(block, 1, 1);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    PQ_CUDA_CHECK(cudaLaunchKernelEx(&lc, fn, args...));
}

inline int dtype_bits(int dt) {
    switch (dt) {
        case DT_F32: return 32;
        case DT_BF16: return 16;
        case DT_U2: return 2;
        case DT_U4: return 4;
        case DT_U8: return 8;
        case DT_I2: return 2;
        case DT_I4: return 4;
        case DT_I8: return 8;
        default: return 0;
    }
}
inline bool dtype_is_quant(int dt) { return dt == DT_U2 || dt == DT_U4 || dt == DT_U8 || dt == DT_I2 || dt == DT_I4 || dt == DT_I8; }
inline bool dtype_is_signed_quant(int dt) { return dt == DT_I2 || dt == DT_I4 || dt == DT_I8; }
// the unsigned type whose kernels serve `dt` (identity for everything but the signed extension types)
inline int dtype_kernel_view(int dt) { return dt == DT_I2 ? DT_U2 : dt == DT_I4 ? DT_U4 : dt == DT_I8 ? DT_U8 : dt; }
// sign bit of every field of a packed 32-bit word, 0 for unsigned types
inline uint32_t dtype_sign_xor(int dt) { return dt == DT_I2 ? 0xAAAAAAAAu : dt == DT_I4 ? 0x88888888u : dt == DT_I8 ? 0x80808080u : 0u; }
// 2^(bits-1) for signed types: added to the zero point to reach the offset-binary view
inline int64_t dtype_zp_offset(int dt) { return dtype_is_signed_quant(dt) ? (int64_t{1} << (dtype_bits(dt) - 1)) : 0; }
inline bool dtype_is_float(int dt) { return dt == DT_F32 || dt == DT_BF16; }
inline const char* dtype_name(int dt) {
    switch (dt) {
        case DT_F32: return "f32";
        case DT_BF16: return "bf16";
        case DT_U2: return "uint2";
        case DT_U4: return "uint4";
        case DT_U8: return "uint8";
        case DT_I2: return "int2";
        case DT_I4: return "int4";
        case DT_I8: return "int8";
        default: return "?";
    }
}
// ceil(numel * bits / 8): storage bytes of a packed quantized tensor (reference src/piquant_internal.hpp:41-44)
inline size_t packed_bytes(int dt, size_t numel) {
    const size_t per = 8u / static_cast<size_t>(dtype_bits(dt));
    return (numel + per - 1) / per;
}
inline size_t storage_bytes(int dt, size_t numel) {
    return dtype_is_quant(dt) ? packed_bytes(dt, numel) : numel * static_cast<size_t>(dtype_bits(dt) / 8);
}

// dt_quant: the caller's quantized dtype; for a signed type the returned parameters are those of the unsigned
// kernel view (zero point + 2^(bits-1), wrapping; sign_xor set) -- launch with dtype_kernel_view(dt_quant).
QuantParams make_params(float scale, int64_t zero_point, float xi, int dt_quant = DT_U8);

// Device pointers (or device-accessible mapped host pointers) only.  All launches are asynchronous
// on cfg.stream.  `mode`: 0 nearest, 1 stochastic (P.xi), 2 per-element stochastic (quantize only; cfg.sr_key / sr_base).
// Returns the number of kernels launched.
int launch_quantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int mode,
                    const LaunchCfg& cfg, const QuantParams* device_params = nullptr);
int launch_dequantize(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P, int op,
                      const LaunchCfg& cfg, const QuantParams* device_params = nullptr);
// fused quantize -> dequantize, unpacked (reference src/kernels/kernels.inl:30-52)
int launch_requantize(const void* in, int dt_inout, void* out, int dt_quant, int64_t numel, const QuantParams& P,
                      int mode, int op, const LaunchCfg& cfg, const QuantParams* device_params = nullptr);

// Scratch of the ticketed single-launch reductions: one set per (device, stream) -- two streams never share a ticket.
struct MinMaxScratch {
    float2*   partials;   // [max_blocks]
    unsigned* ticket;     // last-block-done counter, self-resetting
    int       max_blocks;
};

// Parameters of one quantized tensor, produced and consumed on the device (include/piquant_cuda.h:
// piquant_cuda_meta_t has the same 64-byte layout: 16 public bytes, then the kernel-side parameters).
struct DeviceMeta {
    float       scale;
    int32_t     error;        // 1: the reference would have aborted (scale NaN or negative)
    int64_t     zero_point;
    QuantParams P;
    char        pad[64 - 16 - sizeof(QuantParams)];
};
static_assert(sizeof(QuantParams) == 40, "QuantParams must keep fitting the opaque part of piquant_cuda_meta_t");
static_assert(sizeof(DeviceMeta) == 64 && offsetof(DeviceMeta, P) == 16, "piquant_cuda_meta_t layout");

// ---- whole-tensor {-min, max} across ranks, inside the kernel -----------------------------------------------------
// Every rank owns a small mailbox in its own HBM, mapped into every other rank's address space (CUDA IPC over
// NVLink / NVSwitch).  Exchange number `seq` (the same on every rank: collectives of one context are issued in the
// same order everywhere): rank r stores its pair into slot [seq & 1][r] of EVERY rank's mailbox as two 8-byte words
// {seq : value bits} -- an aligned 8-byte store is single-copy atomic, so a reader sees either the old or the new
// word, never a torn one -- then polls its own mailbox until all `nranks` slots of this parity carry `seq`, and folds
// them.  Two parities are enough: a rank cannot finish exchange s before every peer has written its s-word, which a
// peer only does after it has finished READING exchange s - 1, so slot [s & 1] is never overwritten (by s + 2)
// while someone still waits on it.
constexpr int kMaxPeers = 16;
constexpr int kMailboxWords = 2 * kMaxPeers * 2;      // [parity][rank][{-min, max}] x u64

struct PeerExchange {
    unsigned long long* box[kMaxPeers];   // box[r]: rank r's mailbox as mapped here (box[rank] is this GPU's own)
    int      nranks;                      // <= 1: no exchange
    int      rank;
    unsigned seq;                         // >= 1
};

// What the tail writes and where.
struct ReduceTail {
    float2*     partials;        // [gridDim.x] per-CTA {min, max}
    unsigned*   ticket;          // last-CTA-done counter, self-resetting
    float*      result;          // device: {min, max, -min, max}
    float*      mapped_result;   // pinned-mapped host copy of the same, or nullptr
    DeviceMeta* meta_out;        // not null: also evaluate (scale, zero_point) for the quantized type below
    DeviceMeta* meta_out2;       // second copy (e.g. the receiver's slot on a peer GPU), or nullptr
    DeviceMeta* meta_mapped;     // pinned-mapped host copy, or nullptr
    int         q_bits;          // 2 / 4 / 8
    int         q_signed;
    uint32_t    q_sign_xor;
    PeerExchange px;
};

// Host-side description of what a reduction launch has to deliver (ReduceTail without the scratch).
struct ReduceOut {
    float*      result = nullptr;         // device {min, max, -min, max}; required
    float*      mapped_result = nullptr;  // pinned-mapped host copy, optional
    DeviceMeta* meta_out = nullptr;       // quantization parameters for dt_quant, optional
    DeviceMeta* meta_out2 = nullptr;
    DeviceMeta* meta_mapped = nullptr;
    int         dt_quant = -1;            // quantized dtype the parameters are for (may be a signed extension type)
    const PeerExchange* px = nullptr;     // whole-tensor result over the ranks of a communicator, optional
};
ReduceTail make_reduce_tail(const MinMaxScratch& scratch, const ReduceOut& ro);

// One launch: result (+ optional parameter block, + optional cross-rank exchange) of x.  min/max start from +-FLT_MAX
// and NaNs never win, like the reference (src/kernels/kernels_specialized.inl:1418-1607).
// keep_in_l2: read with L2::evict_last instead of evict_first, so that a tensor smaller than the 126 MB L2 is still
// resident when the quantize pass that follows reads it again.
int launch_minmax(const void* x, int dt, int64_t numel, const MinMaxScratch& scratch, const ReduceOut& ro,
                  const LaunchCfg& cfg, bool keep_in_l2 = false);

// Fused passes of a quantized ring reduction (dequantize.cu).  Both return the number of kernels launched and fall back
// to separate launches with identical results when the buffers cannot be vector-aligned.
//   ADD + min/max: out += dequantize(in), and `ro` receives min/max (+ parameters, + rank exchange) of the sums written.
//   SET + forward: out = dequantize(in), and the packed bytes of `in` (+ the 64-byte block fwd_meta_src) are also stored to
//   fwd (+ fwd_meta_dst) -- typically the next rank's receive slot in NVLink peer memory.
int launch_dequantize_add_minmax(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P,
                                 const LaunchCfg& cfg, const QuantParams* device_params, const MinMaxScratch& scratch, const ReduceOut& ro);
int launch_dequantize_forward(const void* in, int dt_in, void* out, int dt_out, int64_t numel, const QuantParams& P,
                              const LaunchCfg& cfg, const QuantParams* device_params, void* fwd, const DeviceMeta* fwd_meta_src,
                              DeviceMeta* fwd_meta_dst);

// out += dequantize(ins[0]) + ... + dequantize(ins[n_src-1]) in one pass, sources folded in order with the ADD store op's
// arithmetic (bit-identical to n_src separate ADD launches), min/max (+ parameters) of the sums to `ro` (reduce_sum.cu).
constexpr int kMaxSumSources = 8;
int launch_dequantize_sum_minmax(const void* const* ins, const QuantParams* const* device_params, int n_src, int dt_in, void* out, int dt_out,
                                 int64_t numel, const LaunchCfg& cfg, const MinMaxScratch& scratch, const ReduceOut& ro);

// Many small tensors, ONE launch (quantize.cu): tensors[i] = {in, out, numel, scale, zero_point}, all with the same
// (dt_in, dt_out, mode).  Returns the number of kernels launched (one per 256 tensors).
struct BatchItem {
    const void* in;
    void*       out;
    int64_t     numel;
    float       scale;
    int64_t     zero_point;
};
int launch_quantize_batch(const BatchItem* items, int count, int dt_in, int dt_out_view, int dt_out, int mode, float xi, const LaunchCfg& cfg);

// One-thread kernel: the double-precision scale / zero-point arithmetic of the reference
// (src/piquant.cpp:245-258) on the {-min, max} pair at minmax4[2..3], bit-identical to the host version,
// plus everything the kernels derive from it (1/scale, bias, range flags).  Asynchronous on cfg.stream.
// dt_quant may be a signed extension type: the public (scale, zero_point) are then the signed ones
// (reference src/piquant.cpp:245-258 with type_min = -type_max - 1) and P is the unsigned kernel view.
int launch_params(const float* minmax4, int dt_quant, DeviceMeta* out, DeviceMeta* mapped_out, const LaunchCfg& cfg);
// stream-ordered wait for a 4-byte arrival flag in device memory (set to non-zero by a copy engine of this or a peer GPU); resets it
int launch_wait_flag(unsigned* flag, const LaunchCfg& cfg);
// minmax4[0..1] = {-minmax4[2], minmax4[3]}: after {-min, max} was combined across ranks by an all-reduce
int launch_minmax_publish(float* minmax4, const LaunchCfg& cfg);

// Kernel families.  The direct LDG.256 / STG kernels are the product default (variant 0 = variant 1): with one tile per
// CTA dealt by the hardware scheduler they run at the traffic ceiling of the chip at every size and in every cell (f32->u8:
// 7.19 TB/s against 7.23 TB/s for the same read:write mix with no arithmetic at all, profiles/r1_sched_probe_static_vs_dynamic_tiles.txt).
// The TMA (cp.async.bulk + mbarrier ring) family is an opt-in variant (2): it ties the direct kernels on f32 inputs at
// >= 0.25 GB per launch and trails them everywhere else -- bf16 inputs 82-90 % vs 103-106 % at 1e9 because 768 consumer
// threads per SM convert what the direct kernel spreads over 2048, and a 4-stage ring fill per CTA dominates short launches
// (profiles/r1_sizebench_final_direct_vs_tma.txt, profiles/r2_ncu_bf16_u4_direct_vs_tma.txt).  It is kept tested to the same
// bit-exact bar, is selectable per context (piquant_cuda_set_kernel_variant), and serves by default exactly one case the
// direct kernels cannot: inputs that are 16- but not 32-byte aligned (LDG.256 needs 32, bulk copies need 16).
// There is deliberately no size- or cell-dependent switch: no measured point favours the TMA family.

}  // namespace pq
