// context.cu -- the CUDA-stream dispatcher and the C ABI of libpiquant.so.
//
// Replaces, for the B200, the reference's host runtime:
//   * extern "C" shims                      reference src/capi.cpp:19-104
//   * context / pimpl, fork-join dispatch   reference src/piquant.cpp:113-211
//   * quantization-parameter arithmetic     reference src/piquant.cpp:213-259, :371-381
//   * panic()                               reference src/piquant.cpp:88-98
// The thread pool, per-thread partitioner and CPUID kernel selection have no counterpart: the
// "threads" are CTAs dealt tile by tile by the GPU's own scheduler, the "join" is stream order, and
// the only ISA is sm_100a.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <map>
#include <mutex>
#include <random>

#include "../../include/piquant.h"
#include "../../include/piquant_cuda.h"
#include "pq_kernels.h"

static_assert(int(pq::DT_F32) == int(PIQUANT_DTYPE_F32) && int(pq::DT_BF16) == int(PIQUANT_DTYPE_BF16) && int(pq::DT_U2) == int(PIQUANT_DTYPE_UINT2) &&
              int(pq::DT_U4) == int(PIQUANT_DTYPE_UINT4) && int(pq::DT_U8) == int(PIQUANT_DTYPE_UINT8), "dtype enum ABI");
static_assert(int(pq::OP_SET) == int(PIQUANT_REDUCE_OP_SET) && int(pq::OP_ADD) == int(PIQUANT_REDUCE_OP_ADD), "reduce-op enum ABI");

namespace pq {

void panic(const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    fprintf(stderr, "\x1b[31mpiquant: %s\x1b[0m\n", buf);
    fflush(stderr);
    abort();
}

QuantParams make_params(float scale, int64_t zero_point, float xi, int dt_quant) {
    // signed extension types: the kernels work on the offset-binary view (pq_device.cuh)
    zero_point = static_cast<int64_t>(static_cast<uint64_t>(zero_point) + static_cast<uint64_t>(dtype_zp_offset(dt_quant)));
    QuantParams P;
    P.sign_xor = dtype_sign_xor(dt_quant);
    P.scale = scale;
    P.inv_scale = 1.0f / scale;                               // IEEE divide, once (kernels_specialized.inl:42)
    P.xi = xi;
    P.zp64 = zero_point;
    P.zp32 = static_cast<int32_t>(static_cast<uint32_t>(static_cast<uint64_t>(zero_point)));   // int64 -> int32 truncation (quantize.inl:112)
    volatile float nzp = -static_cast<float>(P.zp32);         // two roundings, never contracted
    P.bias = nzp * scale;
    P.bigzp = (zero_point > (1ll << 29) || zero_point < -(1ll << 29)) ? 1 : 0;
    P.spec_ok32 = (P.zp32 <= (1 << 29) && P.zp32 >= -(1 << 29)) ? 1 : 0;
    return P;
}

namespace {

// ---- NCCL, resolved at run time so that the library has no link-time dependency on it ----------
struct Nccl {
    using Comm = void*;
    struct UniqueId { char internal[128]; };
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;

    static Nccl& get() {
        static Nccl n = load();
        return n;
    }
    static Nccl load() {
        Nccl n;
        void* h = nullptr;
        const char* env = getenv("PIQUANT_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* name : names) {
            if (!name) continue;
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return n;
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(h, "ncclAllReduce"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.GetErrorString;
        return n;
    }
};
constexpr int kNcclFloat32 = 7, kNcclMax = 2;   // nccl.h: ncclFloat32, ncclMax

#define PQ_NCCL_CHECK(expr)                                                                               \
    do {                                                                                                  \
        int pq_r__ = (expr);                                                                              \
        if (pq_r__ != 0) ::pq::panic("%s:%d NCCL error: %s <- %s", __FILE__, __LINE__, Nccl::get().GetErrorString(pq_r__), #expr); \
    } while (0)

// ---- per-device resources ------------------------------------------------------------------------
constexpr int kRing = 3;                         // host-pointer pipeline depth
constexpr int kSchedSlots = 16;                  // distinct streams one context can drive concurrently
constexpr size_t kChunkElems = size_t(8) << 20;  // elements per pipeline chunk (f32: 32 MiB in flight per slot)

struct DeviceState {
    int           device = -1;
    int           sm_count = 0;
    MinMaxScratch scratch{};
    float*        d_result = nullptr;        // 4 floats
    float*        h_result = nullptr;        // pinned + mapped, 4 floats
    float*        h_result_dev = nullptr;    // device view of h_result
    DeviceMeta*   d_meta = nullptr;          // parameters produced on the device (one-shot quantize)
    DeviceMeta*   h_meta = nullptr;          // pinned + mapped copy the host reads after the sync
    DeviceMeta*   h_meta_dev = nullptr;
    // work counters of the persistent TMA kernels: one {next tile, finished CTAs} pair per stream in use
    unsigned long long* d_sched = nullptr;
    cudaStream_t  sched_stream[kSchedSlots]{};
    int           sched_used = 0;
    // host-pointer pipeline (lazily created)
    cudaStream_t  s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr;
    cudaEvent_t   ev_h2d[kRing]{}, ev_run[kRing]{}, ev_d2h[kRing]{};
    void*         d_in[kRing]{};
    void*         d_out[kRing]{};
    size_t        in_cap = 0, out_cap = 0;
    float*        h_parts = nullptr;          // host-tensor min/max: one result slot per pipeline chunk
    size_t        h_parts_cap = 0;
    bool          pipe_ready = false;
};

enum class Where { Device, HostPinned, HostPageable };

struct PtrInfo {
    Where where;
    int   device;     // owning device for Device memory, -1 otherwise
};

}  // namespace

struct Context {
    size_t                     num_threads = 0;
    cudaStream_t               stream = nullptr;
    int                        variant = 0;
    bool                       xi_fixed = false;
    float                      xi = 0.0f, last_xi = 0.0f;
    bool                       sr_key_fixed = false;      // per-element stochastic rounding (piquant_cuda.h)
    uint64_t                   sr_key = 0, last_sr_key = 0;
    std::mt19937_64            rng{std::random_device{}()};
    std::mutex                 rng_mu;                    // the draws happen before the dispatch lock is taken
    std::mutex                 mu;
    std::map<int, DeviceState> devs;
    uint64_t                   launches = 0;
    Nccl::Comm                 comm = nullptr;
    int                        host_mode = 0;    // 0 = staged pipeline for every host pointer, 1 = zero-copy kernels on pinned host memory

    ~Context() {
        if (comm && Nccl::get().ok) Nccl::get().CommDestroy(comm);
        for (auto& kv : devs) {
            DeviceState& d = kv.second;
            int prev = 0;
            if (cudaGetDevice(&prev) != cudaSuccess) break;
            cudaSetDevice(d.device);
            cudaFree(d.scratch.partials);
            cudaFree(d.scratch.ticket);
            cudaFree(d.d_result);
            cudaFreeHost(d.h_result);
            cudaFree(d.d_meta);
            cudaFree(d.d_sched);
            cudaFreeHost(d.h_meta);
            if (d.h_parts) cudaFreeHost(d.h_parts);
            if (d.pipe_ready) {
                for (int i = 0; i < kRing; ++i) {
                    cudaFree(d.d_in[i]);
                    cudaFree(d.d_out[i]);
                    cudaEventDestroy(d.ev_h2d[i]);
                    cudaEventDestroy(d.ev_run[i]);
                    cudaEventDestroy(d.ev_d2h[i]);
                }
                cudaStreamDestroy(d.s_h2d);
                cudaStreamDestroy(d.s_run);
                cudaStreamDestroy(d.s_d2h);
            }
            cudaSetDevice(prev);
        }
    }

    float draw_xi() {
        // one threshold per call, U[0,1) (reference src/piquant.cpp:199-201; its generator is thread_local, here a lock
        // keeps concurrent callers of one context off the shared one)
        std::lock_guard<std::mutex> lock(rng_mu);
        last_xi = xi_fixed ? xi : std::uniform_real_distribution<float>{0.0f, 1.0f}(rng);
        return last_xi;
    }

    uint64_t draw_sr_key() {
        // one Philox key per call: fresh random bits for every element of every call, replayable after piquant_cuda_seed
        std::lock_guard<std::mutex> lock(rng_mu);
        last_sr_key = sr_key_fixed ? sr_key : rng();
        return last_sr_key;
    }

    DeviceState& dev_state(int device) {
        auto it = devs.find(device);
        if (it != devs.end()) return it->second;
        DeviceState d;
        d.device = device;
        cudaDeviceProp prop{};
        PQ_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            panic("device %d (%s) has compute capability %d.%d; this library contains sm_100a code only", device, prop.name,
                  prop.major, prop.minor);
        d.sm_count = prop.multiProcessorCount;
        d.scratch.max_blocks = 16384;
        PQ_CUDA_CHECK(cudaMalloc(&d.scratch.partials, sizeof(float2) * d.scratch.max_blocks));
        PQ_CUDA_CHECK(cudaMalloc(&d.scratch.ticket, sizeof(unsigned)));
        PQ_CUDA_CHECK(cudaMemset(d.scratch.ticket, 0, sizeof(unsigned)));
        PQ_CUDA_CHECK(cudaMalloc(&d.d_result, 4 * sizeof(float)));
        PQ_CUDA_CHECK(cudaHostAlloc(&d.h_result, 4 * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
        PQ_CUDA_CHECK(cudaHostGetDevicePointer(&d.h_result_dev, d.h_result, 0));
        PQ_CUDA_CHECK(cudaMalloc(&d.d_meta, sizeof(DeviceMeta)));
        PQ_CUDA_CHECK(cudaMalloc(&d.d_sched, sizeof(unsigned long long) * 2 * kSchedSlots));
        PQ_CUDA_CHECK(cudaMemset(d.d_sched, 0, sizeof(unsigned long long) * 2 * kSchedSlots));
        PQ_CUDA_CHECK(cudaHostAlloc(&d.h_meta, sizeof(DeviceMeta), cudaHostAllocMapped | cudaHostAllocPortable));
        PQ_CUDA_CHECK(cudaHostGetDevicePointer(&d.h_meta_dev, d.h_meta, 0));
        PQ_CUDA_CHECK(cudaDeviceSynchronize());
        return devs.emplace(device, d).first->second;
    }

    void ensure_pipe(DeviceState& d, size_t in_bytes, size_t out_bytes) {
        if (!d.pipe_ready) {
            PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking));
            PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&d.s_run, cudaStreamNonBlocking));
            PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking));
            for (int i = 0; i < kRing; ++i) {
                PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_h2d[i], cudaEventDisableTiming));
                PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_run[i], cudaEventDisableTiming));
                PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_d2h[i], cudaEventDisableTiming));
            }
            d.pipe_ready = true;
        }
        if (in_bytes > d.in_cap) {
            for (int i = 0; i < kRing; ++i) {
                if (d.d_in[i]) PQ_CUDA_CHECK(cudaFree(d.d_in[i]));
                PQ_CUDA_CHECK(cudaMalloc(&d.d_in[i], in_bytes));
            }
            d.in_cap = in_bytes;
        }
        if (out_bytes > d.out_cap) {
            for (int i = 0; i < kRing; ++i) {
                if (d.d_out[i]) PQ_CUDA_CHECK(cudaFree(d.d_out[i]));
                PQ_CUDA_CHECK(cudaMalloc(&d.d_out[i], out_bytes));
            }
            d.out_cap = out_bytes;
        }
    }
};

namespace {

// Launch configuration for `stream`: kernels of one stream run one after the other, so each stream owns one
// self-resetting work-counter pair; two streams never share one.
LaunchCfg make_cfg(Context& c, DeviceState& d, cudaStream_t stream) {
    int slot = -1;
    for (int i = 0; i < d.sched_used; ++i)
        if (d.sched_stream[i] == stream) slot = i;
    if (slot < 0) {
        if (d.sched_used == kSchedSlots) {          // table full: quiesce the device, every pair is zero again
            PQ_CUDA_CHECK(cudaDeviceSynchronize());
            d.sched_used = 0;
        }
        slot = d.sched_used++;
        d.sched_stream[slot] = stream;
    }
    return LaunchCfg{stream, d.sm_count, c.variant, d.d_sched + 2 * slot};
}

int require_device() {
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        panic("no usable CUDA device (%s); libpiquant.so has no CPU path -- the B200 build runs on sm_100a only",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    int cur = 0;
    PQ_CUDA_CHECK(cudaGetDevice(&cur));
    return cur;
}

PtrInfo classify(const void* p) {
    cudaPointerAttributes attr{};
    const cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return {Where::HostPageable, -1};
    }
    switch (attr.type) {
        case cudaMemoryTypeDevice: return {Where::Device, attr.device};
        case cudaMemoryTypeManaged: {   // managed memory migrates on demand: run where the caller is if it has no home device
            int dev = attr.device;
            if (dev < 0) PQ_CUDA_CHECK(cudaGetDevice(&dev));
            return {Where::Device, dev};
        }
        case cudaMemoryTypeHost: return {Where::HostPinned, -1};
        default: return {Where::HostPageable, -1};
    }
}

// RAII: run on `device`, restore the caller's current device afterwards
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    DeviceGuard(int current, int device) : prev(current) {
        if (device != current) {
            PQ_CUDA_CHECK(cudaSetDevice(device));
            switched = true;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

enum class Cmd { Quant, Dequant, Requant };

struct Job {
    Cmd         cmd;
    const void* in;
    int         dt_in;
    void*       out;
    int         dt_out;      // for Requant: the quantized dtype
    size_t      numel;
    QuantParams P;
    int         mode;
    int         op;
    const QuantParams* dP = nullptr;   // parameters that live in device memory (produced by params_kernel)
    uint64_t    sr_key = 0;            // mode 2 (per-element stochastic rounding): Philox key of this call
};

// bytes of the `in` / `out` buffers for a range of `n` elements
size_t job_in_bytes(const Job& j, size_t n) { return storage_bytes(j.dt_in, n); }
size_t job_out_bytes(const Job& j, size_t n) { return j.cmd == Cmd::Requant ? storage_bytes(j.dt_in, n) : storage_bytes(j.dt_out, n); }

// e0: index of the first element of this launch in the caller's tensor (host-pointer chunks)
int launch_job(const Job& j, const void* in, void* out, size_t n, const LaunchCfg& cfg0, size_t e0 = 0) {
    LaunchCfg cfg = cfg0;
    cfg.sr_key = j.sr_key;
    cfg.sr_base = static_cast<int64_t>(e0);
    switch (j.cmd) {
        case Cmd::Quant: return launch_quantize(in, j.dt_in, out, j.dt_out, static_cast<int64_t>(n), j.P, j.mode, cfg, j.dP);
        case Cmd::Dequant: return launch_dequantize(in, j.dt_in, out, j.dt_out, static_cast<int64_t>(n), j.P, j.op, cfg, j.dP);
        default: return launch_requantize(in, j.dt_in, out, j.dt_out, static_cast<int64_t>(n), j.P, j.mode, j.op, cfg, j.dP);
    }
}

// Host-pointer path: stream the tensor through the GPU in chunks, three stages overlapped on three
// streams (H2D copy of chunk i+1 | kernel on chunk i | D2H copy of chunk i-1).  Synchronous.
void run_staged(Context& c, DeviceState& d, const Job& j, bool in_host, bool out_host) {
    const size_t chunk = kChunkElems;     // multiple of every pack width and of 128 elements
    const bool out_rmw = j.op == OP_ADD && j.cmd != Cmd::Quant;
    const size_t per_slot = j.numel < chunk ? j.numel : chunk;          // small tensors get small staging buffers
    c.ensure_pipe(d, in_host ? job_in_bytes(j, per_slot) : 0, out_host ? job_out_bytes(j, per_slot) : 0);
    // everything already queued on the context stream (producers of device-side operands) goes first
    cudaEvent_t& gate = d.ev_h2d[0];
    PQ_CUDA_CHECK(cudaEventRecord(gate, c.stream));
    PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, gate, 0));
    PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_run, gate, 0));
    const LaunchCfg cfg = make_cfg(c, d, d.s_run);
    size_t i = 0;
    for (size_t e0 = 0; e0 < j.numel; e0 += chunk, ++i) {
        const size_t n = (j.numel - e0 < chunk) ? j.numel - e0 : chunk;
        const int k = static_cast<int>(i % kRing);
        const char* src = static_cast<const char*>(j.in) + job_in_bytes(j, e0);
        char* dst = static_cast<char*>(j.out) + job_out_bytes(j, e0);
        const void* k_in = src;
        void* k_out = dst;
        if (i >= kRing) {
            // slot reuse: its previous kernel must have consumed d_in, its previous D2H must have drained d_out
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, d.ev_run[k], 0));
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, d.ev_d2h[k], 0));
        }
        if (in_host) {
            PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_in[k], src, job_in_bytes(j, n), cudaMemcpyHostToDevice, d.s_h2d));
            k_in = d.d_in[k];
        }
        if (out_host) {
            if (out_rmw) PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_out[k], dst, job_out_bytes(j, n), cudaMemcpyHostToDevice, d.s_h2d));
            k_out = d.d_out[k];
        }
        PQ_CUDA_CHECK(cudaEventRecord(d.ev_h2d[k], d.s_h2d));
        PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_run, d.ev_h2d[k], 0));
        c.launches += launch_job(j, k_in, k_out, n, cfg, e0);
        PQ_CUDA_CHECK(cudaEventRecord(d.ev_run[k], d.s_run));
        if (out_host) {
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_d2h, d.ev_run[k], 0));
            PQ_CUDA_CHECK(cudaMemcpyAsync(dst, d.d_out[k], job_out_bytes(j, n), cudaMemcpyDeviceToHost, d.s_d2h));
        }
        PQ_CUDA_CHECK(cudaEventRecord(d.ev_d2h[k], d.s_d2h));
    }
    PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_h2d));
    PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_run));
    PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_d2h));
}

void run_job(Context& c, const Job& j) {
    if (j.numel == 0) return;
    std::lock_guard<std::mutex> lock(c.mu);
    const int cur = require_device();
    const PtrInfo pi = classify(j.in), po = classify(j.out);
    int device = cur;
    if (pi.where == Where::Device) device = pi.device;
    else if (po.where == Where::Device) device = po.device;
    if (pi.where == Where::Device && po.where == Where::Device && pi.device != po.device)
        panic("input lives on device %d but output on device %d; shard-local buffers are required", pi.device, po.device);
    DeviceGuard guard(cur, device);
    DeviceState& d = c.dev_state(device);
    const bool zero_copy = c.host_mode == 1;
    const bool in_host = pi.where == Where::HostPageable || (pi.where == Where::HostPinned && !zero_copy);
    const bool out_host = po.where == Where::HostPageable || (po.where == Where::HostPinned && !zero_copy);
    if (!in_host && !out_host) {
        const void* in = j.in;
        void* out = j.out;
        if (pi.where == Where::HostPinned) PQ_CUDA_CHECK(cudaHostGetDevicePointer(const_cast<void**>(&in), const_cast<void*>(j.in), 0));
        if (po.where == Where::HostPinned) PQ_CUDA_CHECK(cudaHostGetDevicePointer(&out, j.out, 0));
        const LaunchCfg cfg = make_cfg(c, d, c.stream);
        c.launches += launch_job(j, in, out, j.numel, cfg);
        // pinned host operands: keep the reference's synchronous contract
        if (pi.where == Where::HostPinned || po.where == Where::HostPinned) PQ_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return;
    }
    run_staged(c, d, j, in_host, out_host);
}

int64_t x86_cvttsd_i64(double a) {
    return (a >= -9223372036854775808.0 && a < 9223372036854775808.0) ? static_cast<int64_t>(a) : std::numeric_limits<int64_t>::min();
}

// compute_quant_config after the gather (reference src/piquant.cpp:245-258), same double arithmetic
void params_from_minmax(double r_min, double r_max, int dt_quant, float* scale, int64_t* zero_point) {
    pq_assert(dtype_is_quant(dt_quant), "type %s is not a quantization type", dtype_name(dt_quant));
    // compute_type_max / type_min (reference src/piquant.cpp:212-220, :246-248): a signed type loses one bit of range
    const bool is_signed = dtype_is_signed_quant(dt_quant);
    const uint64_t type_max = (1ull << (dtype_bits(dt_quant) - (is_signed ? 1 : 0))) - 1;
    const int64_t type_min = is_signed ? -static_cast<int64_t>(type_max) - 1 : 0;
    float s;
    int64_t z;
    if (r_max == r_min) {
        s = 1.0f;
        // unsigned: (type_max + type_min) >> 1 as in the reference.  For a signed type the reference's expression would wrap
        // (uint64 + int64 -> uint64, logical shift: INT64_MAX) -- dead code there; here the signed midpoint, -1.
        z = is_signed ? -1 : static_cast<int64_t>((type_max + static_cast<uint64_t>(type_min)) >> 1);
    } else {
        const double q_min = static_cast<double>(type_min), q_max = static_cast<double>(type_max);
        const double sd = (r_max - r_min) / (q_max - q_min);
        double zp = q_min - r_min / sd;
        zp = std::fmax(std::fmin(static_cast<double>(x86_cvttsd_i64(std::round(zp))), q_max), q_min);
        s = static_cast<float>(sd);
        z = x86_cvttsd_i64(zp);
    }
    pq_assert(!std::isnan(s) && s >= 0.0f, "scale must be positive");      // reference src/piquant.cpp:373
    *scale = s;
    *zero_point = z;
}

void compute_params(Context& c, const void* x, int dt_in, size_t n, int dt_quant, float* out_scale, int64_t* out_zp) {
    pq_assert(dtype_is_quant(dt_quant), "type %s is not a quantization type", dtype_name(dt_quant));
    std::lock_guard<std::mutex> lock(c.mu);
    const int cur = require_device();
    const PtrInfo pi = classify(x);
    const int device = pi.where == Where::Device ? pi.device : cur;
    DeviceGuard guard(cur, device);
    DeviceState& d = c.dev_state(device);
    float mn = std::numeric_limits<float>::max(), mx = std::numeric_limits<float>::lowest();
    // an empty shard contributes {+FLT_MAX, -FLT_MAX}; an empty whole tensor ends in a negative scale -> abort below,
    // exactly what the reference does (reference src/piquant.cpp:238-244, :373)
    if (n > 0) {
        if (pi.where == Where::Device || (pi.where == Where::HostPinned && c.host_mode == 1)) {
            const void* xp = x;
            if (pi.where == Where::HostPinned) PQ_CUDA_CHECK(cudaHostGetDevicePointer(const_cast<void**>(&xp), const_cast<void*>(x), 0));
            const LaunchCfg cfg = make_cfg(c, d, c.stream);
            c.launches += launch_minmax(xp, dt_in, static_cast<int64_t>(n), d.scratch, d.d_result, c.comm ? nullptr : d.h_result_dev, cfg);
        } else {
            // host tensor: chunks through the ring, partial results folded on the host
            const size_t chunk = kChunkElems * 2;
            const size_t isz = static_cast<size_t>(dtype_bits(dt_in) / 8);
            c.ensure_pipe(d, (n < chunk ? n : chunk) * isz, 0);
            const LaunchCfg cfg = make_cfg(c, d, d.s_run);
            const size_t n_chunks = (n + chunk - 1) / chunk;
            if (n_chunks > d.h_parts_cap) {                   // per-chunk {min,max,-min,max}, pinned + mapped, kept for the next call
                if (d.h_parts) PQ_CUDA_CHECK(cudaFreeHost(d.h_parts));
                d.h_parts_cap = n_chunks < 256 ? 256 : n_chunks;
                PQ_CUDA_CHECK(cudaHostAlloc(&d.h_parts, d.h_parts_cap * 4 * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
            }
            float* h_parts = d.h_parts;
            float* h_parts_dev = nullptr;
            PQ_CUDA_CHECK(cudaHostGetDevicePointer(&h_parts_dev, h_parts, 0));
            size_t i = 0;
            for (size_t e0 = 0; e0 < n; e0 += chunk, ++i) {
                const size_t m = (n - e0 < chunk) ? n - e0 : chunk;
                const int k = static_cast<int>(i % kRing);
                if (i >= kRing) PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, d.ev_run[k], 0));
                PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_in[k], static_cast<const char*>(x) + e0 * isz, m * isz, cudaMemcpyHostToDevice, d.s_h2d));
                PQ_CUDA_CHECK(cudaEventRecord(d.ev_h2d[k], d.s_h2d));
                PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_run, d.ev_h2d[k], 0));
                c.launches += launch_minmax(d.d_in[k], dt_in, static_cast<int64_t>(m), d.scratch, d.d_result, h_parts_dev + 4 * i, cfg);
                PQ_CUDA_CHECK(cudaEventRecord(d.ev_run[k], d.s_run));
            }
            PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_run));
            for (size_t q = 0; q < n_chunks; ++q) {
                mn = std::fmin(mn, h_parts[4 * q]);
                mx = std::fmax(mx, h_parts[4 * q + 1]);
            }
            if (c.comm) {
                const float r[4] = {mn, mx, -mn, mx};
                PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_result, r, sizeof(r), cudaMemcpyHostToDevice, c.stream));
            } else {
                d.h_result[0] = mn;
                d.h_result[1] = mx;
            }
        }
    } else {
        const float r[4] = {mn, mx, -mn, mx};
        if (c.comm) PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_result, r, sizeof(r), cudaMemcpyHostToDevice, c.stream));
        d.h_result[0] = mn;
        d.h_result[1] = mx;
    }
    if (c.comm) {
        // the one exchange step of a sharded tensor: max over ranks of {-min, max}
        PQ_NCCL_CHECK(Nccl::get().AllReduce(d.d_result + 2, d.d_result + 2, 2, kNcclFloat32, kNcclMax, c.comm, c.stream));
        PQ_CUDA_CHECK(cudaMemcpyAsync(d.h_result, d.d_result, 4 * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        PQ_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        mn = -d.h_result[2];
        mx = d.h_result[3];
    } else {
        PQ_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        mn = d.h_result[0];
        mx = d.h_result[1];
    }
    params_from_minmax(static_cast<double>(mn), static_cast<double>(mx), dt_quant, out_scale, out_zp);
}

Context* as_ctx(piquant_context_t* p) {
    pq_assert(p != nullptr, "context must not be NULL");
    return reinterpret_cast<Context*>(p);
}

void check_float_ptr(const void* p, int dt, const char* what) {
    pq_assert(p != nullptr, "%s pointer must not be NULL", what);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    pq_assert(a % static_cast<uintptr_t>(dtype_bits(dt) / 8) == 0, "%s pointer %p is not aligned for %s", what, p, dtype_name(dt));
}

}  // namespace
}  // namespace pq

using namespace pq;

// -------------------------------------------------------------------------------------------------
// piquant.h
// -------------------------------------------------------------------------------------------------

extern "C" piquant_context_t* piquant_context_create(size_t num_threads) {
    Context* c = new Context();      // no CUDA call here: the Python package creates a context at import time
    c->num_threads = num_threads;
    if (const char* v = getenv("PIQUANT_CUDA_VARIANT")) c->variant = atoi(v);
    if (const char* v = getenv("PIQUANT_CUDA_HOST_MODE")) c->host_mode = (strcmp(v, "zerocopy") == 0 || strcmp(v, "1") == 0) ? 1 : 0;
    return reinterpret_cast<piquant_context_t*>(c);
}

extern "C" void piquant_context_destroy(piquant_context_t* ctx) { delete reinterpret_cast<Context*>(ctx); }

extern "C" void piquant_quantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                 piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                 piquant_round_mode_t mode) {
    Context* c = as_ctx(ctx);
    // reference src/piquant.cpp:288-289
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    pq_assert(mode == PIQUANT_NEAREST || mode == PIQUANT_STOCHASTIC || mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT,
              "invalid round mode %d", static_cast<int>(mode));
    if (numel == 0) return;
    check_float_ptr(in, dtype_in, "input");
    pq_assert(out != nullptr, "output pointer must not be NULL");
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    Job j{Cmd::Quant, in, dtype_in, out, dtype_kernel_view(dtype_out), numel, make_params(scale, zero_point, xi, dtype_out), static_cast<int>(mode), OP_SET};
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) j.sr_key = c->draw_sr_key();
    run_job(*c, j);
}

extern "C" void piquant_dequantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                   piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                   piquant_reduce_op_t op) {
    Context* c = as_ctx(ctx);
    // reference src/piquant.cpp:321-322
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    pq_assert(op == PIQUANT_REDUCE_OP_SET || op == PIQUANT_REDUCE_OP_ADD, "invalid reduce op %d", static_cast<int>(op));
    if (numel == 0) return;
    pq_assert(in != nullptr, "input pointer must not be NULL");
    check_float_ptr(out, dtype_out, "output");
    Job j{Cmd::Dequant, in, dtype_kernel_view(dtype_in), out, dtype_out, numel, make_params(scale, zero_point, 0.0f, dtype_in), 0, static_cast<int>(op)};
    run_job(*c, j);
}

extern "C" void piquant_compute_quant_params_float32(piquant_context_t* ctx, const float* x, size_t n,
                                                     piquant_dtype_t target_quant_dtype, float* out_scale,
                                                     int64_t* out_zero_point) {
    if (n) check_float_ptr(x, DT_F32, "input");
    compute_params(*as_ctx(ctx), x, DT_F32, n, target_quant_dtype, out_scale, out_zero_point);
}

extern "C" void piquant_compute_quant_params_bfloat16(piquant_context_t* ctx, const uint16_t* x, size_t n,
                                                      piquant_dtype_t target_quant_dtype, float* out_scale,
                                                      int64_t* out_zero_point) {
    if (n) check_float_ptr(x, DT_BF16, "input");
    compute_params(*as_ctx(ctx), x, DT_BF16, n, target_quant_dtype, out_scale, out_zero_point);
}

// -------------------------------------------------------------------------------------------------
// piquant_cuda.h
// -------------------------------------------------------------------------------------------------

extern "C" void piquant_cuda_set_stream(piquant_context_t* ctx, void* cuda_stream) {
    as_ctx(ctx)->stream = static_cast<cudaStream_t>(cuda_stream);
}
extern "C" void* piquant_cuda_get_stream(piquant_context_t* ctx) { return as_ctx(ctx)->stream; }

extern "C" void piquant_cuda_synchronize(piquant_context_t* ctx) {
    require_device();
    PQ_CUDA_CHECK(cudaStreamSynchronize(as_ctx(ctx)->stream));
}

extern "C" void piquant_cuda_set_kernel_variant(piquant_context_t* ctx, int variant) {
    pq_assert(variant >= 0 && variant <= 2, "kernel variant must be 0 (auto), 1 (direct) or 2 (tma), got %d", variant);
    as_ctx(ctx)->variant = variant;
}

extern "C" uint64_t piquant_cuda_kernel_launches(piquant_context_t* ctx) { return as_ctx(ctx)->launches; }

extern "C" int piquant_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void piquant_cuda_set_stochastic_threshold(piquant_context_t* ctx, float xi) {
    Context* c = as_ctx(ctx);
    if (xi < 0.0f) {
        c->xi_fixed = false;
    } else {
        pq_assert(xi < 1.0f, "stochastic threshold must be in [0, 1), got %f", static_cast<double>(xi));
        c->xi_fixed = true;
        c->xi = xi;
    }
}

extern "C" void piquant_cuda_seed(piquant_context_t* ctx, uint64_t seed) {
    Context* c = as_ctx(ctx);
    std::lock_guard<std::mutex> lock(c->rng_mu);
    c->rng.seed(seed);
}

extern "C" float piquant_cuda_last_stochastic_threshold(piquant_context_t* ctx) { return as_ctx(ctx)->last_xi; }

extern "C" void piquant_cuda_set_sr_key(piquant_context_t* ctx, uint64_t key) {
    Context* c = as_ctx(ctx);
    c->sr_key_fixed = true;
    c->sr_key = key;
}
extern "C" void piquant_cuda_clear_sr_key(piquant_context_t* ctx) { as_ctx(ctx)->sr_key_fixed = false; }
extern "C" uint64_t piquant_cuda_last_sr_key(piquant_context_t* ctx) { return as_ctx(ctx)->last_sr_key; }

extern "C" void piquant_cuda_requantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in_out, void* out,
                                        piquant_dtype_t quant_dtype, size_t numel, float scale, int64_t zero_point,
                                        piquant_round_mode_t mode, piquant_reduce_op_t op) {
    Context* c = as_ctx(ctx);
    // reference src/piquant.cpp:353-354
    pq_assert(dtype_is_float(dtype_in_out), "input dtype must be a dequantized type");
    pq_assert(dtype_is_quant(quant_dtype), "quant dtype must be a quantized type");
    pq_assert(mode == PIQUANT_NEAREST || mode == PIQUANT_STOCHASTIC || mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT,
              "invalid round mode %d", static_cast<int>(mode));
    if (numel == 0) return;
    check_float_ptr(in, dtype_in_out, "input");
    check_float_ptr(out, dtype_in_out, "output");
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    Job j{Cmd::Requant, in, dtype_in_out, out, dtype_kernel_view(quant_dtype), numel, make_params(scale, zero_point, xi, quant_dtype), static_cast<int>(mode), static_cast<int>(op)};
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) j.sr_key = c->draw_sr_key();
    run_job(*c, j);
}

extern "C" void piquant_cuda_minmax_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n, float* out4) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype), "min/max input must be f32 or bf16");
    pq_assert(n > 0, "min/max of an empty tensor");
    std::lock_guard<std::mutex> lock(c->mu);
    const int cur = require_device();
    const PtrInfo pi = classify(x), po = classify(out4);
    pq_assert(pi.where == Where::Device && po.where == Where::Device, "piquant_cuda_minmax_async needs device pointers");
    DeviceGuard guard(cur, pi.device);
    DeviceState& d = c->dev_state(pi.device);
    const LaunchCfg cfg = make_cfg(*c, d, c->stream);
    c->launches += launch_minmax(x, dtype, static_cast<int64_t>(n), d.scratch, out4, nullptr, cfg);
}

extern "C" void piquant_cuda_params_from_minmax(float min, float max, piquant_dtype_t target_quant_dtype, float* out_scale,
                                                int64_t* out_zero_point) {
    params_from_minmax(static_cast<double>(min), static_cast<double>(max), target_quant_dtype, out_scale, out_zero_point);
}

extern "C" int piquant_cuda_nccl_unique_id(void* out128) {
    Nccl& n = Nccl::get();
    if (!n.ok) return -1;
    Nccl::UniqueId id;
    if (n.GetUniqueId(&id) != 0) return -1;
    memcpy(out128, id.internal, sizeof(id.internal));
    return 0;
}

extern "C" void piquant_cuda_comm_init_rank(piquant_context_t* ctx, const void* unique_id128, int nranks, int rank) {
    Context* c = as_ctx(ctx);
    Nccl& n = Nccl::get();
    pq_assert(n.ok, "libnccl.so.2 could not be loaded (set PIQUANT_NCCL_LIB)");
    pq_assert(c->comm == nullptr, "context already has a communicator");
    require_device();
    Nccl::UniqueId id;
    memcpy(id.internal, unique_id128, sizeof(id.internal));
    PQ_NCCL_CHECK(n.CommInitRank(&c->comm, nranks, id, rank));
}

extern "C" void piquant_cuda_comm_destroy(piquant_context_t* ctx) {
    Context* c = as_ctx(ctx);
    if (c->comm) {
        PQ_NCCL_CHECK(Nccl::get().CommDestroy(c->comm));
        c->comm = nullptr;
    }
}

// -------------------------------------------------------------------------------------------------
// device-resident parameters: min/max -> (scale, zero_point) -> quantize without a host round trip
// -------------------------------------------------------------------------------------------------

namespace pq {
namespace {

static_assert(sizeof(piquant_cuda_meta_t) == sizeof(DeviceMeta), "public meta block == kernel-side meta block");

struct DeviceCall {          // common prologue of the *_async entry points: everything must be device memory on ONE device
    int cur;
    int device;
};

DeviceCall require_device_ptrs(const void* a, const void* b, const void* c, const char* who) {
    const int cur = require_device();
    int device = -1;
    // The kernel runs on the device that owns the FIRST buffer (the tensor being read).  The other buffers only have
    // to be device memory: they may live on a peer GPU whose memory is mapped here (NVLink P2P / symmetric memory) --
    // that is how a ring step quantizes straight into its neighbour's receive buffer.
    for (const void* p : {a, b, c}) {
        if (!p) continue;
        const PtrInfo pi = classify(p);
        pq_assert(pi.where == Where::Device, "%s needs CUDA device pointers", who);
        if (device < 0) device = pi.device;
    }
    return {cur, device};
}

// min/max of x -> {-min, max} (all-reduced over the communicator, if any) -> DeviceMeta at d_meta; asynchronous.
// small_tensor_bytes: > 0 asks the min/max pass to leave x in L2 for the pass that follows.
void compute_meta_async(Context& c, DeviceState& d, const void* x, int dt_in, size_t n, int dt_quant, DeviceMeta* d_meta,
                        DeviceMeta* mapped, bool keep_in_l2) {
    const LaunchCfg cfg = make_cfg(c, d, c.stream);
    if (n > 0) {
        c.launches += launch_minmax(x, dt_in, static_cast<int64_t>(n), d.scratch, d.d_result, nullptr, cfg, keep_in_l2);
    } else {
        const float fmax = std::numeric_limits<float>::max();
        const float r[4] = {fmax, -fmax, -fmax, -fmax};          // an empty shard is the identity of the reduction
        PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_result, r, sizeof(r), cudaMemcpyHostToDevice, c.stream));
    }
    if (c.comm) PQ_NCCL_CHECK(Nccl::get().AllReduce(d.d_result + 2, d.d_result + 2, 2, kNcclFloat32, kNcclMax, c.comm, c.stream));
    c.launches += launch_params(d.d_result, dt_quant, d_meta, mapped, cfg);
}

}  // namespace
}  // namespace pq

extern "C" void piquant_cuda_compute_meta_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                piquant_dtype_t target_quant_dtype, piquant_cuda_meta_t* d_meta) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype), "input dtype (%s) must be a dequantized type", dtype_name(dtype));
    pq_assert(dtype_is_quant(target_quant_dtype), "type %s is not a quantization type", dtype_name(target_quant_dtype));
    std::lock_guard<std::mutex> lock(c->mu);
    const DeviceCall dc = require_device_ptrs(n ? x : nullptr, d_meta, nullptr, "piquant_cuda_compute_meta_async");
    DeviceGuard guard(dc.cur, dc.device);
    DeviceState& d = c->dev_state(dc.device);
    compute_meta_async(*c, d, x, dtype, n, target_quant_dtype, reinterpret_cast<DeviceMeta*>(d_meta), nullptr,
                       n * static_cast<size_t>(dtype_bits(dtype) / 8) <= (size_t(96) << 20));
}

extern "C" void piquant_cuda_quantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                 piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                                 const piquant_cuda_meta_t* d_meta) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    if (numel == 0) return;
    std::lock_guard<std::mutex> lock(c->mu);
    const DeviceCall dc = require_device_ptrs(in, out, d_meta, "piquant_cuda_quantize_meta_async");
    DeviceGuard guard(dc.cur, dc.device);
    DeviceState& d = c->dev_state(dc.device);
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    LaunchCfg cfg = make_cfg(*c, d, c->stream);
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) cfg.sr_key = c->draw_sr_key();
    c->launches += launch_quantize(in, dtype_in, out, dtype_kernel_view(dtype_out), static_cast<int64_t>(numel), make_params(1.0f, 0, xi, dtype_out), static_cast<int>(mode),
                                   cfg, &reinterpret_cast<const DeviceMeta*>(d_meta)->P);
}

extern "C" void piquant_cuda_dequantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                   piquant_dtype_t dtype_out, size_t numel, piquant_reduce_op_t op,
                                                   const piquant_cuda_meta_t* d_meta) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    if (numel == 0) return;
    std::lock_guard<std::mutex> lock(c->mu);
    const DeviceCall dc = require_device_ptrs(in, out, d_meta, "piquant_cuda_dequantize_meta_async");
    DeviceGuard guard(dc.cur, dc.device);
    DeviceState& d = c->dev_state(dc.device);
    const LaunchCfg cfg = make_cfg(*c, d, c->stream);
    c->launches += launch_dequantize(in, dtype_kernel_view(dtype_in), out, dtype_out, static_cast<int64_t>(numel), make_params(1.0f, 0, 0.0f, dtype_in), static_cast<int>(op),
                                     cfg, &reinterpret_cast<const DeviceMeta*>(d_meta)->P);
}

extern "C" void piquant_cuda_quantize_auto(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                           piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode, float* out_scale,
                                           int64_t* out_zero_point) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    pq_assert(numel > 0, "scale must be positive");          // the reference aborts on an empty tensor (src/piquant.cpp:373)
    {
        const PtrInfo pi = (require_device(), classify(in)), po = classify(out);
        if (pi.where != Where::Device || po.where != Where::Device) {   // host tensors: the two-step path
            if (dtype_in == PIQUANT_DTYPE_F32) piquant_compute_quant_params_float32(ctx, static_cast<const float*>(in), numel, dtype_out, out_scale, out_zero_point);
            else piquant_compute_quant_params_bfloat16(ctx, static_cast<const uint16_t*>(in), numel, dtype_out, out_scale, out_zero_point);
            piquant_quantize(ctx, in, dtype_in, out, dtype_out, numel, *out_scale, *out_zero_point, mode);
            return;
        }
    }
    std::lock_guard<std::mutex> lock(c->mu);
    const DeviceCall dc = require_device_ptrs(in, out, nullptr, "piquant_cuda_quantize_auto");
    DeviceGuard guard(dc.cur, dc.device);
    DeviceState& d = c->dev_state(dc.device);
    const size_t in_bytes = numel * static_cast<size_t>(dtype_bits(dtype_in) / 8);
    compute_meta_async(*c, d, in, dtype_in, numel, dtype_out, d.d_meta, d.h_meta_dev, in_bytes <= (size_t(96) << 20));
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    LaunchCfg cfg = make_cfg(*c, d, c->stream);
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) cfg.sr_key = c->draw_sr_key();
    c->launches += launch_quantize(in, dtype_in, out, dtype_kernel_view(dtype_out), static_cast<int64_t>(numel), make_params(1.0f, 0, xi, dtype_out), static_cast<int>(mode),
                                   cfg, &d.d_meta->P);
    PQ_CUDA_CHECK(cudaStreamSynchronize(c->stream));         // the ONE host sync of the whole sequence
    pq_assert(d.h_meta->error == 0, "scale must be positive");
    *out_scale = d.h_meta->scale;
    *out_zero_point = d.h_meta->zero_point;
}
