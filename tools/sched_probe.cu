// sched_probe.cu -- development microbenchmark: does it matter HOW tiles are dealt to CTAs?
// A 4:1 read:write stream (the traffic of f32->u8 quantize, no arithmetic) and a read-only stream, 4 GiB each:
//   static   persistent grid, tile = blockIdx + k*gridDim (what the first kernels did)
//   chunk C  non-persistent grid, CTA b owns C consecutive tiles, the hardware scheduler balances the SMs
//   atomic   persistent grid, next tile from a global atomic counter
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/sched_probe tools/sched_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

// tile = 256 threads * 4 vectors * 32 B = 32 KiB in, 8 KiB out
__device__ __forceinline__ void mix_tile(const char* in, char* out, int64_t tile) {
    const char* ib = in + tile * 32768;
    char* ob = out + tile * 8192;
    uint32_t w[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) ldg256(ib + (j * 256 + threadIdx.x) * 32, w[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t a = w[j][0] ^ w[j][2] ^ w[j][4] ^ w[j][6], b = w[j][1] ^ w[j][3] ^ w[j][5] ^ w[j][7];
        asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(ob + (j * 256 + threadIdx.x) * 8), "r"(a), "r"(b) : "memory");
    }
}
__device__ __forceinline__ uint32_t read_tile(const char* in, int64_t tile) {
    const char* ib = in + tile * 32768;
    uint32_t w[4][8], acc = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) ldg256(ib + (j * 256 + threadIdx.x) * 32, w[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = max(acc, w[j][k]);
    return acc;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) k_static(const char* in, char* out, int64_t n_tiles, uint32_t* sink) {
    uint32_t acc = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if constexpr (WRITE) mix_tile(in, out, t); else acc = max(acc, read_tile(in, t));
    }
    if (!WRITE && acc == 0xdeadbeefu) *sink = acc;
}
template <bool WRITE>
__global__ void __launch_bounds__(256) k_chunk(const char* in, char* out, int64_t n_tiles, int chunk, uint32_t* sink) {
    uint32_t acc = 0;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * chunk;
    const int64_t t1 = t0 + chunk < n_tiles ? t0 + chunk : n_tiles;
    for (int64_t t = t0; t < t1; ++t) {
        if constexpr (WRITE) mix_tile(in, out, t); else acc = max(acc, read_tile(in, t));
    }
    if (!WRITE && acc == 0xdeadbeefu) *sink = acc;
}
template <bool WRITE>
__global__ void __launch_bounds__(256) k_atomic(const char* in, char* out, int64_t n_tiles, int chunk, unsigned long long* counter, uint32_t* sink) {
    __shared__ unsigned long long s_next;
    uint32_t acc = 0;
    while (true) {
        if (threadIdx.x == 0) s_next = atomicAdd(counter, 1ull);
        __syncthreads();
        const int64_t t0 = static_cast<int64_t>(s_next) * chunk;
        __syncthreads();
        if (t0 >= n_tiles) break;
        const int64_t t1 = t0 + chunk < n_tiles ? t0 + chunk : n_tiles;
        for (int64_t t = t0; t < t1; ++t) {
            if constexpr (WRITE) mix_tile(in, out, t); else acc = max(acc, read_tile(in, t));
        }
    }
    if (!WRITE && acc == 0xdeadbeefu) *sink = acc;
}

template <typename F>
float time_ms(F&& launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int64_t in_bytes = int64_t(4) << 30, out_bytes = in_bytes / 4;
    char *in = nullptr, *out = nullptr;
    uint32_t* sink = nullptr;
    unsigned long long* counter = nullptr;
    CK(cudaMalloc(&in, in_bytes));
    CK(cudaMalloc(&out, out_bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&counter, 8));
    CK(cudaMemset(in, 1, in_bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int64_t n_tiles = in_bytes / 32768;
    for (int i = 0; i < 400; ++i) k_chunk<true><<<(unsigned)n_tiles, 256>>>(in, out, n_tiles, 1, sink);   // warm the clocks
    CK(cudaDeviceSynchronize());
    auto rep = [&](const char* name, float ms, double bytes) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
    for (int round = 0; round < 2; ++round) {
        for (int w = 1; w >= 0; --w) {
            const double bytes = w ? double(in_bytes + out_bytes) : double(in_bytes);
            const char* tag = w ? "4R:1W" : "read ";
            char nm[96];
            for (int per_sm : {4, 8}) {
                snprintf(nm, 96, "%s static persistent, %d CTA/SM", tag, per_sm);
                rep(nm, time_ms([&] { if (w) k_static<true><<<sms * per_sm, 256>>>(in, out, n_tiles, sink); else k_static<false><<<sms * per_sm, 256>>>(in, out, n_tiles, sink); }, 10), bytes);
            }
            for (int chunk : {1, 2, 4, 8, 16}) {
                snprintf(nm, 96, "%s chunk %2d (non-persistent)", tag, chunk);
                const unsigned grid = (unsigned)((n_tiles + chunk - 1) / chunk);
                rep(nm, time_ms([&] { if (w) k_chunk<true><<<grid, 256>>>(in, out, n_tiles, chunk, sink); else k_chunk<false><<<grid, 256>>>(in, out, n_tiles, chunk, sink); }, 10), bytes);
            }
            for (int chunk : {1, 4}) {
                snprintf(nm, 96, "%s atomic persistent 8 CTA/SM, chunk %d", tag, chunk);
                rep(nm, time_ms([&] { cudaMemsetAsync(counter, 0, 8); if (w) k_atomic<true><<<sms * 8, 256>>>(in, out, n_tiles, chunk, counter, sink); else k_atomic<false><<<sms * 8, 256>>>(in, out, n_tiles, chunk, counter, sink); }, 10), bytes);
            }
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
