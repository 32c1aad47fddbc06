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
//
// Concurrency model (the reference's context methods are const and callable from several threads,
// reference src/piquant.cpp:194-211):
//   * everything a launch sequence needs that must not be shared -- reduction scratch + ticket, result / parameter
//     blocks, the work counters of the persistent kernels -- lives in a StreamSlot, one per (device, stream) in use;
//     a slot's mutex is held while a sequence (min/max -> quantize ...) is enqueued, so sequences of two threads on
//     one stream cannot interleave, and two streams never touch the same scratch;
//   * the host-pointer pipeline (staging buffers, copy streams) is per device and has its own mutex: a multi-second
//     host transfer blocks neither device-pointer calls nor other devices;
//   * Context::mu only guards the tables themselves and is never held across a launch, a copy or a synchronisation.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <atomic>
#include <condition_variable>
#include <emmintrin.h>
#include <initializer_list>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

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
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
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
        n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.AllGather && n.GetErrorString;
        return n;
    }
};
constexpr int kNcclInt8 = 0, kNcclInt32 = 2, kNcclFloat32 = 7, kNcclMax = 2, kNcclMin = 3;   // nccl.h: ncclInt8, ncclInt32, ncclFloat32, ncclMax, ncclMin

#define PQ_NCCL_CHECK(expr)                                                                               \
    do {                                                                                                  \
        int pq_r__ = (expr);                                                                              \
        if (pq_r__ != 0) ::pq::panic("%s:%d NCCL error: %s <- %s", __FILE__, __LINE__, Nccl::get().GetErrorString(pq_r__), #expr); \
    } while (0)

// ---- host copy workers ---------------------------------------------------------------------------
// Pageable host tensors (what the reference's callers pass: CPU torch tensors, reference python/src/piquant/torch.py:87,117)
// cannot be DMA-ed directly; cudaMemcpyAsync stages them through the driver's own small pinned buffer on the calling
// thread, which neither overlaps nor reaches the link rate.  The library therefore moves them itself: a few host threads
// copy each chunk between the caller's memory and a pinned bounce buffer (non-temporal stores, so the copy does not
// read-for-ownership the destination), and the copy engines take it from there.  This is where the `num_threads` of
// piquant_context_create ends up: it caps the number of copy workers.
void stream_copy(void* dst, const void* src, size_t n) {
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15u)) & 15u;
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head; s += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 0), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 1);
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 2), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 3);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 0, a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 1, b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 2, c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 3, e);
        s += 64; d += 64;
    }
    _mm_sfence();
    memcpy(d, s, n - blocks * 64);
}

class CopyPool {
public:
    explicit CopyPool(int workers) {
        for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    // dst <- src, split into 1 MiB pieces dealt to the workers and the calling thread; returns when all of it is there
    void copy(void* dst, const void* src, size_t bytes) {
        constexpr size_t kPiece = size_t(1) << 20;
        const size_t pieces = (bytes + kPiece - 1) / kPiece;
        if (pieces <= 1 || threads_.empty()) {
            stream_copy(dst, src, bytes);
            return;
        }
        // one distributed copy at a time: calls on different devices of one context share the pool (they are bound by the same
        // host memory system, so taking turns costs nothing)
        std::lock_guard<std::mutex> turn(call_mu_);
        {
            std::lock_guard<std::mutex> lock(mu_);
            dst_ = static_cast<char*>(dst);
            src_ = static_cast<const char*>(src);
            bytes_ = bytes;
            pieces_ = pieces;
            next_.store(0, std::memory_order_relaxed);
            done_ = 0;
            ++generation_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lock(mu_);
        cv_done_.wait(lock, [this] { return done_ == pieces_; });
    }

private:
    void work() {
        constexpr size_t kPiece = size_t(1) << 20;
        size_t mine = 0;
        for (;;) {
            const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= pieces_) break;
            const size_t off = i * kPiece;
            stream_copy(dst_ + off, src_ + off, bytes_ - off < kPiece ? bytes_ - off : kPiece);
            ++mine;
        }
        if (mine) {
            std::lock_guard<std::mutex> lock(mu_);
            done_ += mine;
            if (done_ == pieces_) cv_done_.notify_all();
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            work();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex call_mu_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    bool stop_ = false;
    uint64_t generation_ = 0;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, pieces_ = 0, done_ = 0;
    std::atomic<size_t> next_{0};
};

// ---- per-device, per-stream resources ------------------------------------------------------------
constexpr int kRing = 3;                         // host-pointer pipeline depth
constexpr int kSlots = 16;                       // distinct streams one context can drive concurrently on one device
constexpr size_t kChunkElemsMax = size_t(8) << 20;   // elements per pipeline chunk (f32: 32 MiB in flight per ring entry)
constexpr int kReduceBlocks = 32768;             // per-CTA partials a ticketed reduction may use
constexpr size_t kBounceMinBytes = size_t(8) << 20;  // pageable tensors smaller than this go through plain cudaMemcpyAsync

struct StreamSlot {
    std::mutex    mu;                        // held while a launch sequence is enqueued on `stream`
    cudaStream_t  stream = nullptr;
    bool          used = false;              // bound to `stream`
    MinMaxScratch scratch{};
    float*        d_result = nullptr;        // {min, max, -min, max}
    DeviceMeta*   d_meta = nullptr;          // parameters produced on the device (one-shot quantize)
    unsigned long long* d_sched = nullptr;   // {next tile, finished CTAs} of the persistent TMA kernels
    float*        h_result = nullptr;        // host view ...
    float*        h_result_dev = nullptr;    // ... and device view of the mapped result
    DeviceMeta*   h_meta = nullptr;
    DeviceMeta*   h_meta_dev = nullptr;
};

struct DeviceState {
    int           device = -1;
    int           sm_count = 0;
    StreamSlot    slots[kSlots];
    char*         d_slots = nullptr;         // ONE device allocation carved into the slots: partials | ticket | result | meta | work counters
    char*         h_slots = nullptr;         // ONE pinned + mapped allocation: result | meta per slot
    // host-pointer pipeline (lazily created, sized once)
    std::mutex    pipe_mu;                   // one staged transfer at a time per device
    bool          pipe_ready = false, bounce_ready = false;
    cudaStream_t  s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr;
    cudaEvent_t   ev_gate = nullptr, ev_h2d[kRing]{}, ev_run[kRing]{}, ev_d2h[kRing]{};
    char*         d_in[kRing]{};             // kChunkElemsMax * 4 bytes each
    char*         d_out[kRing]{};
    char*         b_in[kRing]{};             // pinned bounce buffers for pageable tensors, same sizes
    char*         b_out[kRing]{};
    float*        h_parts = nullptr;         // host-tensor min/max: one mapped result block per pipeline chunk
    size_t        h_parts_cap = 0;
};

// whole-tensor parameters of a sharded tensor: the communicator and, when the ranks can map each other's memory, the
// mailboxes of the in-kernel exchange (pq_reduce.cuh)
struct Comm {
    Nccl::Comm    nccl = nullptr;
    int           nranks = 0, rank = 0, device = -1;
    int           transport = 0;             // requested: 0 auto, 1 NCCL all-reduce, 2 peer-memory mailboxes
    bool          p2p = false;               // mailboxes are mapped
    unsigned long long* box[kMaxPeers]{};    // box[r]: rank r's mailbox as mapped into this process
    unsigned      seq = 0;                   // exchanges issued so far
    cudaEvent_t   ev_last = nullptr;         // exchanges issued from different streams are chained through this event
    cudaStream_t  last_stream = nullptr;
    bool          have_last = false;
    std::mutex    mu;
    bool uses_p2p() const { return p2p && transport != 1; }
};

enum class Where { Device, HostPinned, HostPageable };

struct PtrInfo {
    Where where;
    int   device;     // owning device for Device memory, -1 otherwise
};

}  // namespace

struct Context {
    size_t                     num_threads = 0;
    std::atomic<cudaStream_t>  stream{nullptr};
    std::atomic<int>           variant{0};
    bool                       xi_fixed = false;
    float                      xi = 0.0f, last_xi = 0.0f;
    bool                       sr_key_fixed = false;      // per-element stochastic rounding (piquant_cuda.h)
    uint64_t                   sr_key = 0, last_sr_key = 0;
    std::mt19937_64            rng{std::random_device{}()};
    std::mutex                 rng_mu;                    // the draws happen before any other lock is taken
    std::mutex                 mu;                        // guards `devs` and the slot tables; never held across CUDA work
    std::map<int, DeviceState> devs;
    std::atomic<uint64_t>      launches{0};
    Comm                       comm;
    int                        host_mode = 0;    // 0 = staged pipeline for every host pointer, 1 = zero-copy kernels on pinned host memory
    size_t                     chunk_elems = kChunkElemsMax;
    std::unique_ptr<CopyPool>  copiers;
    std::once_flag             copiers_once;

    ~Context() {
        int prev = 0;
        const bool have_dev = cudaGetDevice(&prev) == cudaSuccess;
        comm_teardown();
        for (auto& kv : devs) {
            DeviceState& d = kv.second;
            if (!have_dev) break;
            cudaSetDevice(d.device);
            cudaFree(d.d_slots);
            cudaFreeHost(d.h_slots);
            if (d.h_parts) cudaFreeHost(d.h_parts);
            if (d.pipe_ready) {
                for (int i = 0; i < kRing; ++i) {
                    cudaFree(d.d_in[i]);
                    cudaFree(d.d_out[i]);
                    if (d.b_in[i]) cudaFreeHost(d.b_in[i]);
                    if (d.b_out[i]) cudaFreeHost(d.b_out[i]);
                    cudaEventDestroy(d.ev_h2d[i]);
                    cudaEventDestroy(d.ev_run[i]);
                    cudaEventDestroy(d.ev_d2h[i]);
                }
                cudaEventDestroy(d.ev_gate);
                cudaStreamDestroy(d.s_h2d);
                cudaStreamDestroy(d.s_run);
                cudaStreamDestroy(d.s_d2h);
            }
        }
        if (have_dev) cudaSetDevice(prev);
    }

    void comm_teardown() {
        if (!comm.nccl) return;
        int prev = 0;
        if (cudaGetDevice(&prev) == cudaSuccess && comm.device >= 0) {
            cudaSetDevice(comm.device);
            for (int r = 0; r < comm.nranks && comm.p2p; ++r) {
                if (!comm.box[r]) continue;
                if (r == comm.rank) cudaFree(comm.box[r]);
                else cudaIpcCloseMemHandle(comm.box[r]);
                comm.box[r] = nullptr;
            }
            if (comm.ev_last) cudaEventDestroy(comm.ev_last);
            cudaSetDevice(prev);
        }
        if (Nccl::get().ok) Nccl::get().CommDestroy(comm.nccl);
        comm.nccl = nullptr;
        comm.p2p = false;
        comm.ev_last = nullptr;
        comm.have_last = false;
        comm.seq = 0;
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

    CopyPool& copy_pool() {
        std::call_once(copiers_once, [this] {
            const unsigned hw = std::thread::hardware_concurrency();
            size_t want = num_threads ? num_threads : 1;
            if (hw && want > hw) want = hw;
            // Measured on a 16-core B200 host (tools/pageable_probe.py, profiles/r2_pageable_probe.txt): 1 / 2 / 4 / 6 / 8 / 12 / 16
            // workers move 2.3 / 4.3 / 7.1 / 8.2 / 9.5 / 10.8 / 10.7 Gelem/s of f32 -> u8 (pinned buffers: 13.4), so up to 12 pay.
            // One process per GPU is the usual layout: each takes its share of the cores, at least 2.  The launcher says how many
            // share the box (torchrun: LOCAL_WORLD_SIZE; Open MPI / Slurm equivalents); without one this is the only process.
            long local_world = 1;
            for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"}) {
                if (const char* e = getenv(name)) {
                    const long v = strtol(e, nullptr, 10);
                    if (v >= 1 && v <= 1024) { local_world = v; break; }
                }
            }
            size_t cap = hw ? hw / static_cast<unsigned>(local_world) : 6;
            if (cap < 2) cap = 2;
            if (cap > 12) cap = 12;
            if (const char* e = getenv("PIQUANT_COPY_THREADS")) {      // explicit override
                const long v = strtol(e, nullptr, 10);
                if (v >= 1 && v <= 256) cap = static_cast<size_t>(v);
            }
            if (want > cap) want = cap;
            copiers.reset(new CopyPool(static_cast<int>(want) - 1));     // the calling thread copies too
        });
        return *copiers;
    }

    // call with `mu` held and `device` current
    DeviceState& dev_state(int device) {
        auto it = devs.find(device);
        if (it != devs.end()) return it->second;
        cudaDeviceProp prop{};
        PQ_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            panic("device %d (%s) has compute capability %d.%d; this library contains sm_100a code only", device, prop.name,
                  prop.major, prop.minor);
        DeviceState& d = devs.try_emplace(device).first->second;
        d.device = device;
        d.sm_count = prop.multiProcessorCount;
        carve_slots(d);
        return d;
    }

    // All slots are allocated with the device state, in two allocations: binding a slot to a new stream later needs no
    // CUDA call at all (a first call on a fresh stream may sit inside a stream capture, where cudaMalloc is not allowed).
    static void carve_slots(DeviceState& d);
};

void Context::carve_slots(DeviceState& d) {
    const size_t o_ticket = (sizeof(float2) * kReduceBlocks + 255) / 256 * 256;
    const size_t o_result = o_ticket + 256, o_meta = o_result + 256, o_sched = o_meta + 256, per_slot = o_sched + 256;
    PQ_CUDA_CHECK(cudaMalloc(&d.d_slots, per_slot * kSlots));
    PQ_CUDA_CHECK(cudaMemset(d.d_slots, 0, per_slot * kSlots));
    PQ_CUDA_CHECK(cudaHostAlloc(&d.h_slots, 256 * kSlots, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(d.h_slots, 0, 256 * kSlots);
    char* h_dev = nullptr;
    PQ_CUDA_CHECK(cudaHostGetDevicePointer(&h_dev, d.h_slots, 0));
    for (int i = 0; i < kSlots; ++i) {
        StreamSlot& s = d.slots[i];
        char* base = d.d_slots + per_slot * static_cast<size_t>(i);
        s.scratch.partials = reinterpret_cast<float2*>(base);
        s.scratch.ticket = reinterpret_cast<unsigned*>(base + o_ticket);
        s.scratch.max_blocks = kReduceBlocks;
        s.d_result = reinterpret_cast<float*>(base + o_result);
        s.d_meta = reinterpret_cast<DeviceMeta*>(base + o_meta);
        s.d_sched = reinterpret_cast<unsigned long long*>(base + o_sched);
        s.h_result = reinterpret_cast<float*>(d.h_slots + 256 * i);
        s.h_result_dev = reinterpret_cast<float*>(h_dev + 256 * i);
        s.h_meta = reinterpret_cast<DeviceMeta*>(d.h_slots + 256 * i + 64);
        s.h_meta_dev = reinterpret_cast<DeviceMeta*>(h_dev + 256 * i + 64);
    }
    PQ_CUDA_CHECK(cudaDeviceSynchronize());       // the memset has landed before any stream uses the blocks
}

namespace {

// A launch sequence owns its stream's slot from acquire() until the Lease goes out of scope.
struct Lease {
    DeviceState* d = nullptr;
    StreamSlot*  s = nullptr;
    std::unique_lock<std::mutex> lock;
    LaunchCfg cfg(const Context& c) const { return LaunchCfg{s->stream, d->sm_count, c.variant.load(std::memory_order_relaxed), s->d_sched}; }
};

// `device` must be current.
Lease acquire(Context& c, int device, cudaStream_t stream) {
    for (;;) {
        DeviceState* d;
        StreamSlot* s = nullptr;
        {
            std::lock_guard<std::mutex> lock(c.mu);
            d = &c.dev_state(device);
            StreamSlot* free_slot = nullptr;
            for (StreamSlot& k : d->slots) {
                if (k.used && k.stream == stream) { s = &k; break; }
                if (!k.used && !free_slot) free_slot = &k;
            }
            if (!s && !free_slot) {
                // table full: wait for every sequence in flight, quiesce the device (tickets and work counters are zero
                // again, nothing reads the blocks) and hand the slots out afresh
                for (StreamSlot& k : d->slots) k.mu.lock();
                PQ_CUDA_CHECK(cudaDeviceSynchronize());
                for (StreamSlot& k : d->slots) k.used = false;
                for (StreamSlot& k : d->slots) k.mu.unlock();
                free_slot = &d->slots[0];
            }
            if (!s) {
                s = free_slot;
                s->stream = stream;
                s->used = true;
            }
        }
        std::unique_lock<std::mutex> lk(s->mu);
        if (!s->used || s->stream != stream) continue;        // recycled between the two locks: look again
        Lease l;
        l.d = d;
        l.s = s;
        l.lock = std::move(lk);
        return l;
    }
}

int require_device() {
    static std::atomic<int> count{-1};
    if (count.load(std::memory_order_relaxed) <= 0) {
        int n = 0;
        const cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            panic("no usable CUDA device (%s); libpiquant.so has no CPU path -- the B200 build runs on sm_100a only",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        count.store(n, std::memory_order_relaxed);
    }
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

// Where and on which stream a call runs.  device >= 0: the caller vouches that every pointer is device-accessible memory
// of that device (no classification, no driver query); -1: find out from the pointers.
struct Site {
    int          device;
    cudaStream_t stream;
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
    const QuantParams* dP = nullptr;   // parameters that live in device memory
    uint64_t    sr_key = 0;            // mode 2 (per-element stochastic rounding): Philox key of this call
    bool        reverse = false;
};

// bytes of the `in` / `out` buffers for a range of `n` elements
size_t job_in_bytes(const Job& j, size_t n) { return storage_bytes(j.dt_in, n); }
size_t job_out_bytes(const Job& j, size_t n) { return j.cmd == Cmd::Requant ? storage_bytes(j.dt_in, n) : storage_bytes(j.dt_out, n); }

// e0: index of the first element of this launch in the caller's tensor (host-pointer chunks)
int launch_job(const Job& j, const void* in, void* out, size_t n, const LaunchCfg& cfg0, size_t e0 = 0) {
    LaunchCfg cfg = cfg0;
    cfg.sr_key = j.sr_key;
    cfg.sr_base = static_cast<int64_t>(e0);
    cfg.reverse = j.reverse;
    switch (j.cmd) {
        case Cmd::Quant: return launch_quantize(in, j.dt_in, out, j.dt_out, static_cast<int64_t>(n), j.P, j.mode, cfg, j.dP);
        case Cmd::Dequant: return launch_dequantize(in, j.dt_in, out, j.dt_out, static_cast<int64_t>(n), j.P, j.op, cfg, j.dP);
        default: return launch_requantize(in, j.dt_in, out, j.dt_out, static_cast<int64_t>(n), j.P, j.mode, j.op, cfg, j.dP);
    }
}

// call with d.pipe_mu held and the device current; everything is sized for the largest chunk, once
void ensure_pipe(DeviceState& d, bool bounce) {
    const size_t cap = kChunkElemsMax * 4;
    if (!d.pipe_ready) {
        PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking));
        PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&d.s_run, cudaStreamNonBlocking));
        PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking));
        PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_gate, cudaEventDisableTiming));
        for (int i = 0; i < kRing; ++i) {
            PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_h2d[i], cudaEventDisableTiming));
            PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_run[i], cudaEventDisableTiming));
            PQ_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_d2h[i], cudaEventDisableTiming));
            PQ_CUDA_CHECK(cudaMalloc(&d.d_in[i], cap));
            PQ_CUDA_CHECK(cudaMalloc(&d.d_out[i], cap));
        }
        d.pipe_ready = true;
    }
    if (bounce && !d.bounce_ready) {
        for (int i = 0; i < kRing; ++i) {
            PQ_CUDA_CHECK(cudaHostAlloc(&d.b_in[i], cap, cudaHostAllocPortable));
            PQ_CUDA_CHECK(cudaHostAlloc(&d.b_out[i], cap, cudaHostAllocPortable));
        }
        d.bounce_ready = true;
    }
}

// everything already queued on the caller's stream (producers of device-side operands) goes first -- if that stream
// belongs to this device at all (a stream left bound by a call on another GPU does not, and cannot be recorded here)
void gate_on_caller_stream(DeviceState& d, cudaStream_t stream) {
    int sdev = -1;
    if (cudaStreamGetDevice(stream, &sdev) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (sdev != d.device) return;
    PQ_CUDA_CHECK(cudaEventRecord(d.ev_gate, stream));
    PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, d.ev_gate, 0));
    PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_run, d.ev_gate, 0));
}

// Host-pointer path: stream the tensor through the GPU in chunks, the stages overlapped on three streams
// (H2D copy of chunk i+1 | kernel on chunk i | D2H copy of chunk i-1); pageable tensors add a host-side stage at either
// end (copy workers <-> pinned bounce buffers).  Synchronous, like every call of the reference.
void run_staged(Context& c, DeviceState& d, const Job& j, Where in_where, Where out_where, cudaStream_t caller_stream) {
    const bool in_host = in_where != Where::Device, out_host = out_where != Where::Device;
    const size_t chunk = c.chunk_elems;   // multiple of every pack width and of 128 elements
    const bool out_rmw = j.op == OP_ADD && j.cmd != Cmd::Quant;
    const bool big = job_in_bytes(j, j.numel) + job_out_bytes(j, j.numel) >= kBounceMinBytes;
    const bool in_bounce = in_where == Where::HostPageable && big, out_bounce = out_where == Where::HostPageable && big;
    std::lock_guard<std::mutex> pipe_lock(d.pipe_mu);
    ensure_pipe(d, in_bounce || out_bounce);
    CopyPool* pool = (in_bounce || out_bounce) ? &c.copy_pool() : nullptr;
    gate_on_caller_stream(d, caller_stream);
    Lease lease = acquire(c, d.device, d.s_run);
    const LaunchCfg cfg = lease.cfg(c);
    const size_t n_chunks = (j.numel + chunk - 1) / chunk;
    auto chunk_len = [&](size_t i) { return (j.numel - i * chunk < chunk) ? j.numel - i * chunk : chunk; };
    auto drain = [&](size_t i) {          // chunk i's results: pinned bounce buffer -> the caller's pageable memory
        const int k = static_cast<int>(i % kRing);
        PQ_CUDA_CHECK(cudaEventSynchronize(d.ev_d2h[k]));
        pool->copy(static_cast<char*>(j.out) + job_out_bytes(j, i * chunk), d.b_out[k], job_out_bytes(j, chunk_len(i)));
    };
    for (size_t i = 0; i < n_chunks; ++i) {
        const size_t e0 = i * chunk, n = chunk_len(i);
        const int k = static_cast<int>(i % kRing);
        const char* src = static_cast<const char*>(j.in) + job_in_bytes(j, e0);
        char* dst = static_cast<char*>(j.out) + job_out_bytes(j, e0);
        const void* k_in = src;
        void* k_out = dst;
        if (in_bounce || (out_bounce && out_rmw)) {
            // the bounce buffers of this ring entry are free once its previous H2D copies have left them
            if (i >= kRing) PQ_CUDA_CHECK(cudaEventSynchronize(d.ev_h2d[k]));
            if (in_bounce) pool->copy(d.b_in[k], src, job_in_bytes(j, n));
            if (out_bounce && out_rmw) pool->copy(d.b_out[k], dst, job_out_bytes(j, n));
        }
        if (i >= kRing) {
            // ring entry reuse: its previous kernel must have consumed d_in, its previous D2H must have drained d_out
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, d.ev_run[k], 0));
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_h2d, d.ev_d2h[k], 0));
        }
        if (in_host) {
            PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_in[k], in_bounce ? d.b_in[k] : src, job_in_bytes(j, n), cudaMemcpyHostToDevice, d.s_h2d));
            k_in = d.d_in[k];
        }
        if (out_host) {
            if (out_rmw) PQ_CUDA_CHECK(cudaMemcpyAsync(d.d_out[k], out_bounce ? d.b_out[k] : dst, job_out_bytes(j, n), cudaMemcpyHostToDevice, d.s_h2d));
            k_out = d.d_out[k];
        }
        PQ_CUDA_CHECK(cudaEventRecord(d.ev_h2d[k], d.s_h2d));
        PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_run, d.ev_h2d[k], 0));
        c.launches += launch_job(j, k_in, k_out, n, cfg, e0);
        PQ_CUDA_CHECK(cudaEventRecord(d.ev_run[k], d.s_run));
        if (out_host) {
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d.s_d2h, d.ev_run[k], 0));
            PQ_CUDA_CHECK(cudaMemcpyAsync(out_bounce ? d.b_out[k] : dst, d.d_out[k], job_out_bytes(j, n), cudaMemcpyDeviceToHost, d.s_d2h));
        }
        PQ_CUDA_CHECK(cudaEventRecord(d.ev_d2h[k], d.s_d2h));
        // keep kRing - 1 chunks in flight behind this one; the oldest is copied out while the GPU works on the others
        if (out_bounce && i + 1 >= kRing) drain(i + 1 - kRing);
    }
    if (out_bounce)
        for (size_t i = n_chunks > kRing - 1 ? n_chunks - (kRing - 1) : 0; i < n_chunks; ++i) drain(i);
    PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_h2d));
    PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_run));
    PQ_CUDA_CHECK(cudaStreamSynchronize(d.s_d2h));
}

void run_job(Context& c, const Job& j, const Site& site) {
    if (j.numel == 0) return;
    const int cur = require_device();
    if (site.device >= 0) {                   // the caller knows where its tensors live
        DeviceGuard guard(cur, site.device);
        Lease lease = acquire(c, site.device, site.stream);
        c.launches += launch_job(j, j.in, j.out, j.numel, lease.cfg(c));
        return;
    }
    const PtrInfo pi = classify(j.in), po = classify(j.out);
    int device = cur;
    if (pi.where == Where::Device) device = pi.device;
    else if (po.where == Where::Device) device = po.device;
    if (pi.where == Where::Device && po.where == Where::Device && pi.device != po.device)
        panic("input lives on device %d but output on device %d; shard-local buffers are required", pi.device, po.device);
    DeviceGuard guard(cur, device);
    const bool zero_copy = c.host_mode == 1;
    const bool in_host = pi.where == Where::HostPageable || (pi.where == Where::HostPinned && !zero_copy);
    const bool out_host = po.where == Where::HostPageable || (po.where == Where::HostPinned && !zero_copy);
    if (!in_host && !out_host) {
        const void* in = j.in;
        void* out = j.out;
        if (pi.where == Where::HostPinned) PQ_CUDA_CHECK(cudaHostGetDevicePointer(const_cast<void**>(&in), const_cast<void*>(j.in), 0));
        if (po.where == Where::HostPinned) PQ_CUDA_CHECK(cudaHostGetDevicePointer(&out, j.out, 0));
        {
            Lease lease = acquire(c, device, site.stream);
            c.launches += launch_job(j, in, out, j.numel, lease.cfg(c));
        }
        // pinned host operands: keep the reference's synchronous contract
        if (pi.where == Where::HostPinned || po.where == Where::HostPinned) PQ_CUDA_CHECK(cudaStreamSynchronize(site.stream));
        return;
    }
    DeviceState* d;
    {
        std::lock_guard<std::mutex> lock(c.mu);
        d = &c.dev_state(device);
    }
    run_staged(c, *d, j, in_host ? pi.where : Where::Device, out_host ? po.where : Where::Device, site.stream);
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


// ---- the one exchange step of a sharded tensor -----------------------------------------------------
// Fills `px` for the next exchange on `stream` and returns a pointer to it, or nullptr when this reduction is local (no
// communicator, or the NCCL transport, whose all-reduce the caller then enqueues).  Call with the lease of `stream` held.
const PeerExchange* begin_exchange(Context& c, cudaStream_t stream, bool local_only, PeerExchange& px) {
    Comm& cm = c.comm;
    if (local_only || !cm.nccl || !cm.uses_p2p()) return nullptr;
    std::lock_guard<std::mutex> lock(cm.mu);
    for (int r = 0; r < cm.nranks; ++r) px.box[r] = cm.box[r];
    px.nranks = cm.nranks;
    px.rank = cm.rank;
    px.seq = ++cm.seq;
    // two exchanges of one rank must not be in flight at once (pq_reduce.cuh): chain launches that come from different streams
    if (cm.have_last && cm.last_stream != stream) PQ_CUDA_CHECK(cudaStreamWaitEvent(stream, cm.ev_last, 0));
    return &px;
}
void end_exchange(Context& c, cudaStream_t stream, const PeerExchange* px) {
    if (!px) return;
    Comm& cm = c.comm;
    std::lock_guard<std::mutex> lock(cm.mu);
    PQ_CUDA_CHECK(cudaEventRecord(cm.ev_last, stream));
    cm.last_stream = stream;
    cm.have_last = true;
}
bool wants_nccl_allreduce(const Context& c, bool local_only) { return !local_only && c.comm.nccl && !c.comm.uses_p2p(); }

// min/max of x[0, n) on the lease's stream -> s.d_result (+ optionally the mapped host copy, the parameter blocks), whole-
// tensor over the communicator unless local_only.  Device memory only.  n == 0: the identity of the reduction.
void reduce_on_device(Context& c, Lease& lease, const void* x, int dt_in, size_t n, ReduceOut ro, bool local_only, bool keep_in_l2) {
    StreamSlot& s = *lease.s;
    const LaunchCfg cfg = lease.cfg(c);
    if (wants_nccl_allreduce(c, local_only)) {
        // kernel -> ncclAllReduce(max) of {-min, max} -> (parameters | copy to the host view)
        ReduceOut first;
        first.result = s.d_result;
        c.launches += launch_minmax(x, dt_in, static_cast<int64_t>(n), s.scratch, first, cfg, keep_in_l2);
        PQ_NCCL_CHECK(Nccl::get().AllReduce(s.d_result + 2, s.d_result + 2, 2, kNcclFloat32, kNcclMax, c.comm.nccl, cfg.stream));
        if (ro.result != s.d_result) c.launches += launch_minmax_publish(s.d_result, cfg);      // a caller's block gets all four values
        if (ro.meta_out) {
            c.launches += launch_params(s.d_result, ro.dt_quant, ro.meta_out, ro.meta_mapped, cfg);
            if (ro.meta_out2) PQ_CUDA_CHECK(cudaMemcpyAsync(ro.meta_out2, ro.meta_out, sizeof(DeviceMeta), cudaMemcpyDeviceToDevice, cfg.stream));
        }
        if (ro.result != s.d_result) PQ_CUDA_CHECK(cudaMemcpyAsync(ro.result, s.d_result, 4 * sizeof(float), cudaMemcpyDeviceToDevice, cfg.stream));
        if (ro.mapped_result) PQ_CUDA_CHECK(cudaMemcpyAsync(s.h_result, s.d_result, 4 * sizeof(float), cudaMemcpyDeviceToHost, cfg.stream));
        return;
    }
    PeerExchange px{};
    ro.px = begin_exchange(c, cfg.stream, local_only, px);
    c.launches += launch_minmax(x, dt_in, static_cast<int64_t>(n), s.scratch, ro, cfg, keep_in_l2);
    end_exchange(c, cfg.stream, ro.px);
}

void compute_params(Context& c, const void* x, int dt_in, size_t n, int dt_quant, float* out_scale, int64_t* out_zp, const Site& site) {
    pq_assert(dtype_is_quant(dt_quant), "type %s is not a quantization type", dtype_name(dt_quant));
    const int cur = require_device();
    PtrInfo pi{Where::Device, site.device};
    if (site.device < 0) pi = n ? classify(x) : PtrInfo{Where::Device, cur};
    const int device = pi.where == Where::Device ? pi.device : cur;
    DeviceGuard guard(cur, device);
    float mn, mx;
    // an empty shard contributes {+FLT_MAX, -FLT_MAX}; an empty whole tensor ends in a negative scale -> abort below,
    // exactly what the reference does (reference src/piquant.cpp:238-244, :373)
    if (n == 0 || pi.where == Where::Device || (pi.where == Where::HostPinned && c.host_mode == 1)) {
        const void* xp = x;
        if (n && pi.where == Where::HostPinned) PQ_CUDA_CHECK(cudaHostGetDevicePointer(const_cast<void**>(&xp), const_cast<void*>(x), 0));
        Lease lease = acquire(c, device, site.stream);
        StreamSlot& s = *lease.s;
        ReduceOut ro;
        ro.result = s.d_result;
        ro.mapped_result = s.h_result_dev;
        reduce_on_device(c, lease, xp, dt_in, n, ro, false, false);
        PQ_CUDA_CHECK(cudaStreamSynchronize(site.stream));        // the slot stays leased: h_result is this call's until it is read
        mn = -s.h_result[2];                                      // {-min, max}: the pair every transport of the exchange combines
        mx = s.h_result[3];
    } else {
        // host tensor: chunks through the ring, one mapped result block per chunk, folded on the host
        DeviceState* d;
        {
            std::lock_guard<std::mutex> lock(c.mu);
            d = &c.dev_state(device);
        }
        const size_t chunk = c.chunk_elems;
        const size_t isz = static_cast<size_t>(dtype_bits(dt_in) / 8);
        const bool bounce = pi.where == Where::HostPageable && n * isz >= kBounceMinBytes;
        std::lock_guard<std::mutex> pipe_lock(d->pipe_mu);
        ensure_pipe(*d, bounce);
        CopyPool* pool = bounce ? &c.copy_pool() : nullptr;
        Lease lease = acquire(c, device, d->s_run);
        StreamSlot& s = *lease.s;
        const LaunchCfg cfg = lease.cfg(c);
        const size_t n_chunks = (n + chunk - 1) / chunk;
        if (n_chunks > d->h_parts_cap) {                   // kept for the next call
            if (d->h_parts) PQ_CUDA_CHECK(cudaFreeHost(d->h_parts));
            d->h_parts_cap = n_chunks < 256 ? 256 : n_chunks;
            PQ_CUDA_CHECK(cudaHostAlloc(&d->h_parts, d->h_parts_cap * 4 * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
        }
        float* h_parts = d->h_parts;
        float* h_parts_dev = nullptr;
        PQ_CUDA_CHECK(cudaHostGetDevicePointer(&h_parts_dev, h_parts, 0));
        for (size_t i = 0; i < n_chunks; ++i) {
            const size_t e0 = i * chunk, m = (n - e0 < chunk) ? n - e0 : chunk;
            const int k = static_cast<int>(i % kRing);
            const char* src = static_cast<const char*>(x) + e0 * isz;
            if (bounce) {
                if (i >= kRing) PQ_CUDA_CHECK(cudaEventSynchronize(d->ev_h2d[k]));
                pool->copy(d->b_in[k], src, m * isz);
                src = d->b_in[k];
            }
            if (i >= kRing) PQ_CUDA_CHECK(cudaStreamWaitEvent(d->s_h2d, d->ev_run[k], 0));
            PQ_CUDA_CHECK(cudaMemcpyAsync(d->d_in[k], src, m * isz, cudaMemcpyHostToDevice, d->s_h2d));
            PQ_CUDA_CHECK(cudaEventRecord(d->ev_h2d[k], d->s_h2d));
            PQ_CUDA_CHECK(cudaStreamWaitEvent(d->s_run, d->ev_h2d[k], 0));
            ReduceOut ro;
            ro.result = s.d_result;
            ro.mapped_result = h_parts_dev + 4 * i;
            c.launches += launch_minmax(d->d_in[k], dt_in, static_cast<int64_t>(m), s.scratch, ro, cfg);
            PQ_CUDA_CHECK(cudaEventRecord(d->ev_run[k], d->s_run));
        }
        PQ_CUDA_CHECK(cudaStreamSynchronize(d->s_run));
        mn = std::numeric_limits<float>::max();
        mx = std::numeric_limits<float>::lowest();
        for (size_t q = 0; q < n_chunks; ++q) {
            mn = std::fmin(mn, h_parts[4 * q]);
            mx = std::fmax(mx, h_parts[4 * q + 1]);
        }
        if (c.comm.nccl) {
            // combine with the other ranks: the folded pair goes back to the device as a 2-element tensor (min = mn, max = mx
            // since mn <= mx) and through the same exchange as a device-resident shard
            const float pair[2] = {mn, mx};
            PQ_CUDA_CHECK(cudaMemcpyAsync(d->d_in[0], pair, sizeof(pair), cudaMemcpyHostToDevice, d->s_run));
            ReduceOut ro;
            ro.result = s.d_result;
            ro.mapped_result = s.h_result_dev;
            reduce_on_device(c, lease, d->d_in[0], DT_F32, 2, ro, false, false);
            PQ_CUDA_CHECK(cudaStreamSynchronize(d->s_run));
            mn = -s.h_result[2];
            mx = s.h_result[3];
        }
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

void check_round_mode(int mode) {
    pq_assert(mode == PIQUANT_NEAREST || mode == PIQUANT_STOCHASTIC || mode == static_cast<int>(PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT),
              "invalid round mode %d", mode);
}
void check_reduce_op(int op) { pq_assert(op == PIQUANT_REDUCE_OP_SET || op == PIQUANT_REDUCE_OP_ADD, "invalid reduce op %d", op); }

Site ctx_site(Context& c) { return Site{-1, c.stream.load(std::memory_order_relaxed)}; }
Site call_site(int device, void* stream) {
    pq_assert(device >= -1, "device must be a CUDA device index or PIQUANT_CUDA_DEVICE_AUTO, got %d", device);
    return Site{device, static_cast<cudaStream_t>(stream)};
}

void do_quantize(Context* c, const void* in, piquant_dtype_t dtype_in, void* out, piquant_dtype_t dtype_out, size_t numel, float scale,
                 int64_t zero_point, piquant_round_mode_t mode, const Site& site) {
    // reference src/piquant.cpp:288-289
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    check_round_mode(static_cast<int>(mode));
    if (numel == 0) return;
    check_float_ptr(in, dtype_in, "input");
    pq_assert(out != nullptr, "output pointer must not be NULL");
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    Job j{Cmd::Quant, in, dtype_in, out, dtype_kernel_view(dtype_out), numel, make_params(scale, zero_point, xi, dtype_out), static_cast<int>(mode), OP_SET};
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) j.sr_key = c->draw_sr_key();
    run_job(*c, j, site);
}

void do_dequantize(Context* c, const void* in, piquant_dtype_t dtype_in, void* out, piquant_dtype_t dtype_out, size_t numel, float scale,
                   int64_t zero_point, piquant_reduce_op_t op, const Site& site) {
    // reference src/piquant.cpp:321-322
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    check_reduce_op(static_cast<int>(op));
    if (numel == 0) return;
    pq_assert(in != nullptr, "input pointer must not be NULL");
    check_float_ptr(out, dtype_out, "output");
    Job j{Cmd::Dequant, in, dtype_kernel_view(dtype_in), out, dtype_out, numel, make_params(scale, zero_point, 0.0f, dtype_in), 0, static_cast<int>(op)};
    run_job(*c, j, site);
}

void do_requantize(Context* c, const void* in, piquant_dtype_t dtype_in_out, void* out, piquant_dtype_t quant_dtype, size_t numel, float scale,
                   int64_t zero_point, piquant_round_mode_t mode, piquant_reduce_op_t op, const Site& site) {
    // reference src/piquant.cpp:353-354
    pq_assert(dtype_is_float(dtype_in_out), "input dtype must be a dequantized type");
    pq_assert(dtype_is_quant(quant_dtype), "quant dtype must be a quantized type");
    check_round_mode(static_cast<int>(mode));
    check_reduce_op(static_cast<int>(op));
    if (numel == 0) return;
    check_float_ptr(in, dtype_in_out, "input");
    check_float_ptr(out, dtype_in_out, "output");
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    Job j{Cmd::Requant, in, dtype_in_out, out, dtype_kernel_view(quant_dtype), numel, make_params(scale, zero_point, xi, quant_dtype), static_cast<int>(mode), static_cast<int>(op)};
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) j.sr_key = c->draw_sr_key();
    run_job(*c, j, site);
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
    if (const char* v = getenv("PIQUANT_CUDA_CHUNK_MIB")) {      // elements per pipeline chunk, as MiB of f32 (experiments; default 32)
        size_t mib = static_cast<size_t>(atoi(v));
        if (mib < 1) mib = 1;
        if (mib > 32) mib = 32;
        c->chunk_elems = mib << 18;
    }
    return reinterpret_cast<piquant_context_t*>(c);
}

extern "C" void piquant_context_destroy(piquant_context_t* ctx) { delete reinterpret_cast<Context*>(ctx); }

extern "C" void piquant_quantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                 piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                 piquant_round_mode_t mode) {
    Context* c = as_ctx(ctx);
    do_quantize(c, in, dtype_in, out, dtype_out, numel, scale, zero_point, mode, ctx_site(*c));
}

extern "C" void piquant_dequantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                   piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                   piquant_reduce_op_t op) {
    Context* c = as_ctx(ctx);
    do_dequantize(c, in, dtype_in, out, dtype_out, numel, scale, zero_point, op, ctx_site(*c));
}

extern "C" void piquant_compute_quant_params_float32(piquant_context_t* ctx, const float* x, size_t n,
                                                     piquant_dtype_t target_quant_dtype, float* out_scale,
                                                     int64_t* out_zero_point) {
    Context* c = as_ctx(ctx);
    if (n) check_float_ptr(x, DT_F32, "input");
    compute_params(*c, x, DT_F32, n, target_quant_dtype, out_scale, out_zero_point, ctx_site(*c));
}

extern "C" void piquant_compute_quant_params_bfloat16(piquant_context_t* ctx, const uint16_t* x, size_t n,
                                                      piquant_dtype_t target_quant_dtype, float* out_scale,
                                                      int64_t* out_zero_point) {
    Context* c = as_ctx(ctx);
    if (n) check_float_ptr(x, DT_BF16, "input");
    compute_params(*c, x, DT_BF16, n, target_quant_dtype, out_scale, out_zero_point, ctx_site(*c));
}

// -------------------------------------------------------------------------------------------------
// piquant_cuda.h
// -------------------------------------------------------------------------------------------------

extern "C" void piquant_cuda_set_stream(piquant_context_t* ctx, void* cuda_stream) {
    as_ctx(ctx)->stream.store(static_cast<cudaStream_t>(cuda_stream));
}
extern "C" void* piquant_cuda_get_stream(piquant_context_t* ctx) { return as_ctx(ctx)->stream.load(); }

extern "C" void piquant_cuda_synchronize(piquant_context_t* ctx) {
    require_device();
    PQ_CUDA_CHECK(cudaStreamSynchronize(as_ctx(ctx)->stream.load()));
}

extern "C" void piquant_cuda_set_kernel_variant(piquant_context_t* ctx, int variant) {
    pq_assert(variant >= 0 && variant <= 2, "kernel variant must be 0 (auto), 1 (direct) or 2 (tma), got %d", variant);
    as_ctx(ctx)->variant = variant;
}

extern "C" uint64_t piquant_cuda_kernel_launches(piquant_context_t* ctx) { return as_ctx(ctx)->launches.load(); }

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
    std::lock_guard<std::mutex> lock(c->rng_mu);
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
    std::lock_guard<std::mutex> lock(c->rng_mu);
    c->sr_key_fixed = true;
    c->sr_key = key;
}
extern "C" void piquant_cuda_clear_sr_key(piquant_context_t* ctx) {
    Context* c = as_ctx(ctx);
    std::lock_guard<std::mutex> lock(c->rng_mu);
    c->sr_key_fixed = false;
}
extern "C" uint64_t piquant_cuda_last_sr_key(piquant_context_t* ctx) { return as_ctx(ctx)->last_sr_key; }

extern "C" void piquant_cuda_requantize(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in_out, void* out,
                                        piquant_dtype_t quant_dtype, size_t numel, float scale, int64_t zero_point,
                                        piquant_round_mode_t mode, piquant_reduce_op_t op) {
    Context* c = as_ctx(ctx);
    do_requantize(c, in, dtype_in_out, out, quant_dtype, numel, scale, zero_point, mode, op, ctx_site(*c));
}

extern "C" void piquant_cuda_params_from_minmax(float min, float max, piquant_dtype_t target_quant_dtype, float* out_scale,
                                                int64_t* out_zero_point) {
    params_from_minmax(static_cast<double>(min), static_cast<double>(max), target_quant_dtype, out_scale, out_zero_point);
}

// ---- communicator --------------------------------------------------------------------------------

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
    pq_assert(c->comm.nccl == nullptr, "context already has a communicator");
    pq_assert(nranks >= 1 && rank >= 0 && rank < nranks, "invalid rank %d of %d", rank, nranks);
    Comm& cm = c->comm;
    cm.device = require_device();
    Nccl::UniqueId id;
    memcpy(id.internal, unique_id128, sizeof(id.internal));
    PQ_NCCL_CHECK(n.CommInitRank(&cm.nccl, nranks, id, rank));
    cm.nranks = nranks;
    cm.rank = rank;
    cm.seq = 0;
    cm.have_last = false;
    cm.p2p = false;
    PQ_CUDA_CHECK(cudaEventCreateWithFlags(&cm.ev_last, cudaEventDisableTiming));
    if (const char* v = getenv("PIQUANT_CUDA_COMM_TRANSPORT")) cm.transport = strcmp(v, "nccl") == 0 ? 1 : (strcmp(v, "p2p") == 0 ? 2 : 0);
    // Mailboxes for the in-kernel exchange: every rank allocates one, the CUDA IPC handles travel through ONE ncclAllGather,
    // every rank maps the others'.  All ranks must agree on the outcome (a lone rank falling back to NCCL would hang the
    // rest), hence the final MIN all-reduce of the success flag.  Anything that fails leaves the NCCL transport in place.
    int ok = nranks <= kMaxPeers ? 1 : 0;
    unsigned long long* own = nullptr;
    cudaIpcMemHandle_t mine{};
    if (ok && cudaMalloc(&own, kMailboxWords * sizeof(unsigned long long)) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(own, 0, kMailboxWords * sizeof(unsigned long long)) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, own) != cudaSuccess) ok = 0;
    cudaGetLastError();
    char* d_handles = nullptr;
    PQ_CUDA_CHECK(cudaMalloc(&d_handles, static_cast<size_t>(nranks + 1) * sizeof(cudaIpcMemHandle_t) + 16));
    char* d_mine = d_handles + static_cast<size_t>(nranks) * sizeof(cudaIpcMemHandle_t);
    PQ_CUDA_CHECK(cudaMemcpy(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    cudaStream_t st = nullptr;
    PQ_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    PQ_NCCL_CHECK(n.AllGather(d_mine, d_handles, sizeof(cudaIpcMemHandle_t), kNcclInt8, cm.nccl, st));
    PQ_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<cudaIpcMemHandle_t> all(static_cast<size_t>(nranks));
    PQ_CUDA_CHECK(cudaMemcpy(all.data(), d_handles, all.size() * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost));
    int* d_flag = reinterpret_cast<int*>(d_mine + sizeof(cudaIpcMemHandle_t));
    for (int r = 0; r < nranks && ok; ++r) {
        if (r == rank) { cm.box[r] = own; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[static_cast<size_t>(r)], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        cm.box[r] = static_cast<unsigned long long*>(p);
    }
    PQ_CUDA_CHECK(cudaMemcpy(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice));
    PQ_NCCL_CHECK(n.AllReduce(d_flag, d_flag, 1, kNcclInt32, kNcclMin, cm.nccl, st));
    PQ_CUDA_CHECK(cudaStreamSynchronize(st));
    PQ_CUDA_CHECK(cudaMemcpy(&ok, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    PQ_CUDA_CHECK(cudaStreamDestroy(st));
    PQ_CUDA_CHECK(cudaFree(d_handles));
    if (ok) {
        cm.p2p = true;
    } else {
        for (int r = 0; r < nranks; ++r) {
            if (cm.box[r] && r != rank) cudaIpcCloseMemHandle(cm.box[r]);
            cm.box[r] = nullptr;
        }
        if (own) cudaFree(own);
        cudaGetLastError();
    }
    pq_assert(cm.transport != 2 || cm.p2p, "the peer-memory transport was requested but the ranks cannot map each other's memory");
}

extern "C" void piquant_cuda_comm_destroy(piquant_context_t* ctx) { as_ctx(ctx)->comm_teardown(); }

extern "C" void piquant_cuda_comm_set_transport(piquant_context_t* ctx, int transport) {
    Context* c = as_ctx(ctx);
    pq_assert(transport >= 0 && transport <= 2, "transport must be 0 (auto), 1 (nccl) or 2 (peer memory), got %d", transport);
    pq_assert(transport != 2 || !c->comm.nccl || c->comm.p2p, "the peer-memory transport is not available on this communicator");
    c->comm.transport = transport;
}

extern "C" int piquant_cuda_comm_transport(piquant_context_t* ctx) {
    Context* c = as_ctx(ctx);
    if (!c->comm.nccl) return 0;
    return c->comm.uses_p2p() ? 2 : 1;
}

// -------------------------------------------------------------------------------------------------
// explicit device + stream per call; device-resident parameters; fused ring passes; batches
// -------------------------------------------------------------------------------------------------

namespace pq {
namespace {

static_assert(sizeof(piquant_cuda_meta_t) == sizeof(DeviceMeta), "public meta block == kernel-side meta block");
static_assert(sizeof(piquant_cuda_batch_item_t) == sizeof(BatchItem) && offsetof(piquant_cuda_batch_item_t, zero_point) == offsetof(BatchItem, zero_point),
              "public batch item == kernel-side batch item");

// The device a call with device memory only runs on: the site's, or the owner of the FIRST buffer (the tensor being
// read).  The other buffers only have to be device memory: they may live on a peer GPU whose memory is mapped here
// (NVLink P2P / symmetric memory) -- that is how a ring step quantizes straight into its neighbour's receive buffer.
int device_of(const Site& site, std::initializer_list<const void*> ptrs, const char* who) {
    if (site.device >= 0) return site.device;
    int device = -1;
    for (const void* p : ptrs) {
        if (!p) continue;
        const PtrInfo pi = classify(p);
        pq_assert(pi.where == Where::Device, "%s needs CUDA device pointers", who);
        if (device < 0) device = pi.device;
    }
    pq_assert(device >= 0, "%s: no device pointer to run on", who);
    return device;
}

bool fits_l2(size_t numel, int dt) { return numel * static_cast<size_t>(dtype_bits(dt) / 8) <= (size_t(96) << 20); }

}  // namespace
}  // namespace pq

extern "C" void piquant_cuda_quantize_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                                piquant_round_mode_t mode, int device, void* stream) {
    do_quantize(as_ctx(ctx), in, dtype_in, out, dtype_out, numel, scale, zero_point, mode, call_site(device, stream));
}

extern "C" void piquant_cuda_dequantize_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                  piquant_dtype_t dtype_out, size_t numel, float scale, int64_t zero_point,
                                                  piquant_reduce_op_t op, int device, void* stream) {
    do_dequantize(as_ctx(ctx), in, dtype_in, out, dtype_out, numel, scale, zero_point, op, call_site(device, stream));
}

extern "C" void piquant_cuda_requantize_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in_out, void* out,
                                                  piquant_dtype_t quant_dtype, size_t numel, float scale, int64_t zero_point,
                                                  piquant_round_mode_t mode, piquant_reduce_op_t op, int device, void* stream) {
    do_requantize(as_ctx(ctx), in, dtype_in_out, out, quant_dtype, numel, scale, zero_point, mode, op, call_site(device, stream));
}

extern "C" void piquant_cuda_compute_quant_params_on_stream(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                            piquant_dtype_t target_quant_dtype, float* out_scale, int64_t* out_zero_point,
                                                            int device, void* stream) {
    pq_assert(dtype_is_float(dtype), "input dtype (%s) must be a dequantized type", dtype_name(dtype));
    if (n) check_float_ptr(x, dtype, "input");
    compute_params(*as_ctx(ctx), x, dtype, n, target_quant_dtype, out_scale, out_zero_point, call_site(device, stream));
}

extern "C" void piquant_cuda_minmax_on_stream(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n, float* out4,
                                              unsigned flags, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype), "min/max input must be f32 or bf16");
    pq_assert(n > 0, "min/max of an empty tensor");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {x, out4}, "piquant_cuda_minmax");
    DeviceGuard guard(cur, dev);
    Lease lease = acquire(*c, dev, site.stream);
    ReduceOut ro;
    ro.result = out4;
    reduce_on_device(*c, lease, x, dtype, n, ro, (flags & PIQUANT_CUDA_FLAG_LOCAL) != 0, (flags & PIQUANT_CUDA_FLAG_KEEP_IN_L2) != 0);
}

extern "C" void piquant_cuda_minmax_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n, float* out4) {
    // shard-local by definition: this is the building block callers combine themselves (piquant.distributed)
    piquant_cuda_minmax_on_stream(ctx, x, dtype, n, out4, PIQUANT_CUDA_FLAG_LOCAL, PIQUANT_CUDA_DEVICE_AUTO, as_ctx(ctx)->stream.load());
}

extern "C" void piquant_cuda_compute_meta_on_stream(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                    piquant_dtype_t target_quant_dtype, piquant_cuda_meta_t* d_meta, unsigned flags,
                                                    int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype), "input dtype (%s) must be a dequantized type", dtype_name(dtype));
    pq_assert(dtype_is_quant(target_quant_dtype), "type %s is not a quantization type", dtype_name(target_quant_dtype));
    pq_assert(d_meta != nullptr, "parameter block must not be NULL");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {n ? x : nullptr, d_meta}, "piquant_cuda_compute_meta");
    DeviceGuard guard(cur, dev);
    Lease lease = acquire(*c, dev, site.stream);
    ReduceOut ro;
    ro.result = lease.s->d_result;
    ro.meta_out = reinterpret_cast<DeviceMeta*>(d_meta);
    ro.dt_quant = target_quant_dtype;
    reduce_on_device(*c, lease, x, dtype, n, ro, (flags & PIQUANT_CUDA_FLAG_LOCAL) != 0,
                     (flags & PIQUANT_CUDA_FLAG_KEEP_IN_L2) != 0 || fits_l2(n, dtype));
}

extern "C" void piquant_cuda_compute_meta_async(piquant_context_t* ctx, const void* x, piquant_dtype_t dtype, size_t n,
                                                piquant_dtype_t target_quant_dtype, piquant_cuda_meta_t* d_meta) {
    piquant_cuda_compute_meta_on_stream(ctx, x, dtype, n, target_quant_dtype, d_meta, 0, PIQUANT_CUDA_DEVICE_AUTO, as_ctx(ctx)->stream.load());
}

extern "C" void piquant_cuda_quantize_meta_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                     piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                                     const piquant_cuda_meta_t* d_meta, unsigned flags, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    check_round_mode(static_cast<int>(mode));
    if (numel == 0) return;
    check_float_ptr(in, dtype_in, "input");
    pq_assert(out != nullptr && d_meta != nullptr, "output and parameter block must not be NULL");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {in, out, d_meta}, "piquant_cuda_quantize_meta");
    DeviceGuard guard(cur, dev);
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    Job j{Cmd::Quant, in, dtype_in, out, dtype_kernel_view(dtype_out), numel, make_params(1.0f, 0, xi, dtype_out), static_cast<int>(mode), OP_SET};
    j.dP = &reinterpret_cast<const DeviceMeta*>(d_meta)->P;
    j.reverse = (flags & PIQUANT_CUDA_FLAG_REVERSE) != 0;
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) j.sr_key = c->draw_sr_key();
    Lease lease = acquire(*c, dev, site.stream);
    c->launches += launch_job(j, in, out, numel, lease.cfg(*c));
}

extern "C" void piquant_cuda_quantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                 piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode,
                                                 const piquant_cuda_meta_t* d_meta) {
    piquant_cuda_quantize_meta_on_stream(ctx, in, dtype_in, out, dtype_out, numel, mode, d_meta, 0, PIQUANT_CUDA_DEVICE_AUTO, as_ctx(ctx)->stream.load());
}

extern "C" void piquant_cuda_dequantize_meta_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                       piquant_dtype_t dtype_out, size_t numel, piquant_reduce_op_t op,
                                                       const piquant_cuda_meta_t* d_meta, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    check_reduce_op(static_cast<int>(op));
    if (numel == 0) return;
    pq_assert(in != nullptr && d_meta != nullptr, "input and parameter block must not be NULL");
    check_float_ptr(out, dtype_out, "output");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {out, in, d_meta}, "piquant_cuda_dequantize_meta");
    DeviceGuard guard(cur, dev);
    Job j{Cmd::Dequant, in, dtype_kernel_view(dtype_in), out, dtype_out, numel, make_params(1.0f, 0, 0.0f, dtype_in), 0, static_cast<int>(op)};
    j.dP = &reinterpret_cast<const DeviceMeta*>(d_meta)->P;
    Lease lease = acquire(*c, dev, site.stream);
    c->launches += launch_job(j, in, out, numel, lease.cfg(*c));
}

extern "C" void piquant_cuda_dequantize_meta_async(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                   piquant_dtype_t dtype_out, size_t numel, piquant_reduce_op_t op,
                                                   const piquant_cuda_meta_t* d_meta) {
    piquant_cuda_dequantize_meta_on_stream(ctx, in, dtype_in, out, dtype_out, numel, op, d_meta, PIQUANT_CUDA_DEVICE_AUTO, as_ctx(ctx)->stream.load());
}

extern "C" void piquant_cuda_dequantize_add_minmax_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                             piquant_dtype_t dtype_out, size_t numel, const piquant_cuda_meta_t* d_meta,
                                                             piquant_dtype_t next_quant_dtype, piquant_cuda_meta_t* d_meta_next,
                                                             piquant_cuda_meta_t* d_meta_next_copy, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    pq_assert(dtype_is_quant(next_quant_dtype), "type %s is not a quantization type", dtype_name(next_quant_dtype));
    pq_assert(numel > 0, "dequantize-ADD + min/max of an empty tensor");
    pq_assert(in != nullptr && d_meta != nullptr && d_meta_next != nullptr, "input and parameter blocks must not be NULL");
    check_float_ptr(out, dtype_out, "output");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {out, in, d_meta, d_meta_next}, "piquant_cuda_dequantize_add_minmax");
    DeviceGuard guard(cur, dev);
    Lease lease = acquire(*c, dev, site.stream);
    ReduceOut ro;
    ro.result = lease.s->d_result;
    ro.meta_out = reinterpret_cast<DeviceMeta*>(d_meta_next);
    ro.meta_out2 = reinterpret_cast<DeviceMeta*>(d_meta_next_copy);
    ro.dt_quant = next_quant_dtype;
    c->launches += launch_dequantize_add_minmax(in, dtype_kernel_view(dtype_in), out, dtype_out, static_cast<int64_t>(numel), make_params(1.0f, 0, 0.0f, dtype_in),
                                                lease.cfg(*c), &reinterpret_cast<const DeviceMeta*>(d_meta)->P, lease.s->scratch, ro);
}

extern "C" void piquant_cuda_dequantize_sum_minmax_on_stream(piquant_context_t* ctx, const void* const* ins,
                                                             const piquant_cuda_meta_t* const* d_metas, size_t count,
                                                             piquant_dtype_t dtype_in, void* out, piquant_dtype_t dtype_out, size_t numel,
                                                             piquant_dtype_t next_quant_dtype, piquant_cuda_meta_t* d_meta_next,
                                                             piquant_cuda_meta_t* d_meta_next_copy, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    pq_assert(dtype_is_quant(next_quant_dtype), "type %s is not a quantization type", dtype_name(next_quant_dtype));
    pq_assert(numel > 0, "sum of empty tensors");
    pq_assert(count >= 1 && count <= PIQUANT_CUDA_MAX_SUM_SOURCES, "between 1 and %d sources per call (got %zu)", PIQUANT_CUDA_MAX_SUM_SOURCES, count);
    pq_assert(ins != nullptr && d_metas != nullptr && d_meta_next != nullptr, "source arrays and parameter block must not be NULL");
    check_float_ptr(out, dtype_out, "output");
    const void* in_ptrs[kMaxSumSources];
    const QuantParams* params[kMaxSumSources];
    for (size_t s = 0; s < count; ++s) {
        pq_assert(ins[s] != nullptr && d_metas[s] != nullptr, "source %zu: input and parameter block must not be NULL", s);
        in_ptrs[s] = ins[s];
        params[s] = &reinterpret_cast<const DeviceMeta*>(d_metas[s])->P;
    }
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {out, in_ptrs[0], d_metas[0], d_meta_next}, "piquant_cuda_dequantize_sum_minmax");
    DeviceGuard guard(cur, dev);
    Lease lease = acquire(*c, dev, site.stream);
    ReduceOut ro;
    ro.result = lease.s->d_result;
    ro.meta_out = reinterpret_cast<DeviceMeta*>(d_meta_next);
    ro.meta_out2 = reinterpret_cast<DeviceMeta*>(d_meta_next_copy);
    ro.dt_quant = next_quant_dtype;
    c->launches += launch_dequantize_sum_minmax(in_ptrs, params, static_cast<int>(count), dtype_kernel_view(dtype_in), out, dtype_out,
                                                static_cast<int64_t>(numel), lease.cfg(*c), lease.s->scratch, ro);
}

extern "C" void piquant_cuda_dequantize_forward_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                          piquant_dtype_t dtype_out, size_t numel, const piquant_cuda_meta_t* d_meta,
                                                          void* forward_to, piquant_cuda_meta_t* forward_meta_to, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_quant(dtype_in), "input dtype (%s) must be a quantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_float(dtype_out), "output dtype (%s) must be a dequantized type", dtype_name(dtype_out));
    pq_assert(numel > 0, "dequantize + forward of an empty tensor");
    pq_assert(in != nullptr && d_meta != nullptr && forward_to != nullptr, "input, parameter block and forward target must not be NULL");
    check_float_ptr(out, dtype_out, "output");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {out, in, d_meta}, "piquant_cuda_dequantize_forward");
    DeviceGuard guard(cur, dev);
    Lease lease = acquire(*c, dev, site.stream);
    c->launches += launch_dequantize_forward(in, dtype_kernel_view(dtype_in), out, dtype_out, static_cast<int64_t>(numel), make_params(1.0f, 0, 0.0f, dtype_in),
                                             lease.cfg(*c), &reinterpret_cast<const DeviceMeta*>(d_meta)->P, forward_to,
                                             reinterpret_cast<const DeviceMeta*>(d_meta), reinterpret_cast<DeviceMeta*>(forward_meta_to));
}

extern "C" void piquant_cuda_copy_on_stream(piquant_context_t* ctx, void* dst, const void* src, size_t nbytes, int device, void* stream) {
    as_ctx(ctx);
    if (nbytes == 0) return;
    pq_assert(dst != nullptr && src != nullptr, "copy pointers must not be NULL");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {src}, "piquant_cuda_copy");
    DeviceGuard guard(cur, dev);
    PQ_CUDA_CHECK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, site.stream));
}

extern "C" void piquant_cuda_wait_flag_on_stream(piquant_context_t* ctx, void* flag, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(flag != nullptr && (reinterpret_cast<uintptr_t>(flag) & 3u) == 0, "the flag must be a 4-byte aligned device address");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {flag}, "piquant_cuda_wait_flag");
    DeviceGuard guard(cur, dev);
    LaunchCfg cfg{};
    cfg.stream = site.stream;
    c->launches += launch_wait_flag(static_cast<unsigned*>(flag), cfg);
}

extern "C" void piquant_cuda_quantize_auto_on_stream(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                                     piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode, float* out_scale,
                                                     int64_t* out_zero_point, int device, void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    check_round_mode(static_cast<int>(mode));
    pq_assert(numel > 0, "scale must be positive");          // the reference aborts on an empty tensor (src/piquant.cpp:373)
    check_float_ptr(in, dtype_in, "input");
    pq_assert(out != nullptr, "output pointer must not be NULL");
    const Site site = call_site(device, stream);
    const int cur = require_device();
    int dev = site.device;
    if (dev < 0) {
        const PtrInfo pi = classify(in), po = classify(out);
        if (pi.where != Where::Device || po.where != Where::Device) {   // host tensors: the two-step path
            compute_params(*c, in, dtype_in, numel, dtype_out, out_scale, out_zero_point, site);
            do_quantize(c, in, dtype_in, out, dtype_out, numel, *out_scale, *out_zero_point, mode, site);
            return;
        }
        dev = pi.device;
    }
    DeviceGuard guard(cur, dev);
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;
    Job j{Cmd::Quant, in, dtype_in, out, dtype_kernel_view(dtype_out), numel, make_params(1.0f, 0, xi, dtype_out), static_cast<int>(mode), OP_SET};
    if (mode == PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT) j.sr_key = c->draw_sr_key();
    Lease lease = acquire(*c, dev, site.stream);
    StreamSlot& s = *lease.s;
    // launch 1: min/max, the cross-rank exchange (if any) and the parameter arithmetic, all in the reduction's tail;
    // launch 2: quantize with the parameters read from the block.  A tensor that fits L2 is read with evict_last by the
    // first pass and from its END by the second (the part of it L2 still holds), i.e. from HBM about once.
    const bool l2 = fits_l2(numel, dtype_in);
    ReduceOut ro;
    ro.result = s.d_result;
    ro.meta_out = s.d_meta;
    ro.meta_mapped = s.h_meta_dev;
    ro.dt_quant = dtype_out;
    reduce_on_device(*c, lease, in, dtype_in, numel, ro, false, l2);
    j.dP = &s.d_meta->P;
    j.reverse = l2;
    c->launches += launch_job(j, in, out, numel, lease.cfg(*c));
    PQ_CUDA_CHECK(cudaStreamSynchronize(site.stream));         // the ONE host sync of the whole sequence; the slot stays leased until h_meta is read
    pq_assert(s.h_meta->error == 0, "scale must be positive");
    *out_scale = s.h_meta->scale;
    *out_zero_point = s.h_meta->zero_point;
}

extern "C" void piquant_cuda_quantize_auto(piquant_context_t* ctx, const void* in, piquant_dtype_t dtype_in, void* out,
                                           piquant_dtype_t dtype_out, size_t numel, piquant_round_mode_t mode, float* out_scale,
                                           int64_t* out_zero_point) {
    piquant_cuda_quantize_auto_on_stream(ctx, in, dtype_in, out, dtype_out, numel, mode, out_scale, out_zero_point, PIQUANT_CUDA_DEVICE_AUTO,
                                         as_ctx(ctx)->stream.load());
}

extern "C" void piquant_cuda_quantize_batch(piquant_context_t* ctx, const piquant_cuda_batch_item_t* items, size_t count,
                                            piquant_dtype_t dtype_in, piquant_dtype_t dtype_out, piquant_round_mode_t mode, int device,
                                            void* stream) {
    Context* c = as_ctx(ctx);
    pq_assert(dtype_is_float(dtype_in), "input dtype (%s) must be a dequantized type", dtype_name(dtype_in));
    pq_assert(dtype_is_quant(dtype_out), "output dtype (%s) must be a quantized type", dtype_name(dtype_out));
    pq_assert(mode == PIQUANT_NEAREST || mode == PIQUANT_STOCHASTIC, "the batch entry point serves nearest and per-call stochastic rounding, got mode %d",
              static_cast<int>(mode));
    pq_assert(count <= (size_t(1) << 30), "too many tensors in one batch");
    if (count == 0) return;
    pq_assert(items != nullptr, "batch items must not be NULL");
    const void* first_in = nullptr;
    for (size_t i = 0; i < count; ++i) {
        if (items[i].numel == 0) continue;
        check_float_ptr(items[i].in, dtype_in, "input");
        pq_assert(items[i].out != nullptr, "output pointer must not be NULL");
        if (!first_in) first_in = items[i].in;
    }
    if (!first_in) return;
    const Site site = call_site(device, stream);
    const int cur = require_device();
    const int dev = device_of(site, {first_in}, "piquant_cuda_quantize_batch");
    DeviceGuard guard(cur, dev);
    const float xi = mode == PIQUANT_STOCHASTIC ? c->draw_xi() : 0.0f;     // one threshold for the whole batch = one call
    Lease lease = acquire(*c, dev, site.stream);
    c->launches += launch_quantize_batch(reinterpret_cast<const BatchItem*>(items), static_cast<int>(count), dtype_in, dtype_kernel_view(dtype_out), dtype_out,
                                         static_cast<int>(mode), xi, lease.cfg(*c));
}
