// write_probe.cu -- development microbenchmark: how fast can one B200 WRITE 4 GB, by store flavour?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/write_probe tools/write_probe.cu && gpurun_out/write_probe
// Variants: STG.128 default / STG.256 default / STG.256 L1::no_allocate / STG.256 .cs / TMA bulk store from
// shared memory (16 KiB and 32 KiB tiles), each as a persistent grid (k CTAs per SM) and as a one-tile-per-CTA grid.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__device__ __forceinline__ void store32(char* p, uint32_t v) {
    if constexpr (MODE == 0) {
        asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
        asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p + 16), "r"(v) : "memory");
    } else if constexpr (MODE == 1) {
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    } else if constexpr (MODE == 2) {
        asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    } else {
        asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
        asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p + 16), "r"(v) : "memory");
    }
}

// each thread writes 32 B per step; warp-interleaved; tile = 256 threads * 4 * 32 B = 32 KiB
template <int MODE>
__global__ void __launch_bounds__(256) direct_kernel(char* out, int64_t n_tiles, uint32_t v) {
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        char* base = out + tile * 32768;
#pragma unroll
        for (int j = 0; j < 4; ++j) store32<MODE>(base + (j * 256 + threadIdx.x) * 32, v + (uint32_t)tile);
    }
}

// each thread writes 64 contiguous bytes (2 x 32 B) per step, like the direct dequantize kernel
template <int MODE>
__global__ void __launch_bounds__(256) direct64_kernel(char* out, int64_t n_tiles, uint32_t v) {
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        char* base = out + tile * 32768;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            store32<MODE>(base + (j * 256 + threadIdx.x) * 64, v + (uint32_t)tile);
            store32<MODE>(base + (j * 256 + threadIdx.x) * 64 + 32, v + (uint32_t)tile);
        }
    }
}

template <int TILE, int NBUF>
__global__ void __launch_bounds__(256) tma_kernel(char* out, int64_t n_tiles, uint32_t v) {
    extern __shared__ __align__(128) unsigned char smem[];
    int i = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
        unsigned char* buf = smem + (i % NBUF) * TILE;
        if (threadIdx.x == 0 && i >= NBUF) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        __syncthreads();
        uint4* b4 = reinterpret_cast<uint4*>(buf);
#pragma unroll
        for (int j = 0; j < TILE / 16 / 256; ++j) b4[j * 256 + threadIdx.x] = make_uint4(v, v + (uint32_t)tile, v, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + tile * TILE), "r"(smem_u32(buf)), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    const int64_t bytes = int64_t(4) << 30;
    char* out = nullptr;
    CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(out, 0, bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    // warm the clocks
    for (int i = 0; i < 300; ++i) direct_kernel<1><<<sms * 8, 256>>>(out, bytes / 32768, i);
    CK(cudaDeviceSynchronize());
    const int64_t n32 = bytes / 32768;
    auto rep = [&](const char* name, float ms) { printf("%-46s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
    for (int per_sm : {2, 4, 8}) {
        char nm[96];
        snprintf(nm, 96, "STG.128x2 default, persistent %d CTA/SM", per_sm);   rep(nm, time_ms([&] { direct_kernel<0><<<sms * per_sm, 256>>>(out, n32, 1); }, 10));
        snprintf(nm, 96, "STG.256 default, persistent %d CTA/SM", per_sm);     rep(nm, time_ms([&] { direct_kernel<1><<<sms * per_sm, 256>>>(out, n32, 1); }, 10));
        snprintf(nm, 96, "STG.256 L1::no_allocate, persistent %d CTA/SM", per_sm); rep(nm, time_ms([&] { direct_kernel<2><<<sms * per_sm, 256>>>(out, n32, 1); }, 10));
        snprintf(nm, 96, "STG.128x2 .cs, persistent %d CTA/SM", per_sm);       rep(nm, time_ms([&] { direct_kernel<3><<<sms * per_sm, 256>>>(out, n32, 1); }, 10));
        snprintf(nm, 96, "64B/thread STG.256 no_alloc, persistent %d CTA/SM", per_sm); rep(nm, time_ms([&] { direct64_kernel<2><<<sms * per_sm, 256>>>(out, n32, 1); }, 10));
    }
    rep("STG.256 default, one tile per CTA", time_ms([&] { direct_kernel<1><<<(unsigned)n32, 256>>>(out, n32, 1); }, 10));
    rep("STG.256 no_allocate, one tile per CTA", time_ms([&] { direct_kernel<2><<<(unsigned)n32, 256>>>(out, n32, 1); }, 10));
    CK(cudaFuncSetAttribute(tma_kernel<16384, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384));
    CK(cudaFuncSetAttribute(tma_kernel<32768, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768));
    CK(cudaFuncSetAttribute(tma_kernel<16384, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384));
    for (int per_sm : {1, 2, 3, 4}) {
        char nm[96];
        snprintf(nm, 96, "TMA bulk store 16 KiB x3 buf, %d CTA/SM", per_sm); rep(nm, time_ms([&] { tma_kernel<16384, 3><<<sms * per_sm, 256, 3 * 16384>>>(out, bytes / 16384, 1); }, 10));
        snprintf(nm, 96, "TMA bulk store 16 KiB x2 buf, %d CTA/SM", per_sm); rep(nm, time_ms([&] { tma_kernel<16384, 2><<<sms * per_sm, 256, 2 * 16384>>>(out, bytes / 16384, 1); }, 10));
        if (per_sm <= 2) { snprintf(nm, 96, "TMA bulk store 32 KiB x3 buf, %d CTA/SM", per_sm); rep(nm, time_ms([&] { tma_kernel<32768, 3><<<sms * per_sm, 256, 3 * 32768>>>(out, bytes / 32768, 1); }, 10)); }
    }
    rep("cudaMemsetAsync", time_ms([&] { cudaMemsetAsync(out, 1, bytes); }, 10));
    CK(cudaDeviceSynchronize());
    return 0;
}
