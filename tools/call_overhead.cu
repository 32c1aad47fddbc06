// call_overhead.cu -- what one call through the C ABI costs on the host, without Python in the way (development tool).
//
//   nvcc -O2 -std=c++17 -I include tools/call_overhead.cu -L pi-quant_b200/piquant -lpiquant -Xlinker -rpath=$PWD/pi-quant_b200/piquant -o tools/bin/call_overhead
//
// Times, per call, on tiny tensors (4096 elements: the GPU work is negligible, the stream never backs up because every batch of
// calls is followed by a synchronisation that is NOT timed):
//   * an empty kernel launched with <<<>>>                      -- the floor any launch pays
//   * piquant_quantize                (reference ABI: pointer classification = 2 driver queries per call)
//   * piquant_cuda_quantize_on_stream (device and stream passed with the call: no queries)
//   * piquant_dequantize / piquant_cuda_dequantize_on_stream
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

#include "piquant.h"
#include "piquant_cuda.h"

__global__ void empty_kernel() {}

template <typename F>
static double per_call_us(F&& call, int calls = 2000, int reps = 7) {
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < calls; ++i) call();
        const auto t1 = std::chrono::steady_clock::now();
        cudaDeviceSynchronize();
        const double us = std::chrono::duration<double, std::micro>(t1 - t0).count() / calls;
        if (us < best) best = us;
    }
    return best;
}

int main() {
    const size_t n = 4096;
    float* x = nullptr;
    unsigned char* q = nullptr;
    float* y = nullptr;
    cudaMalloc(&x, n * sizeof(float));
    cudaMalloc(&q, n);
    cudaMalloc(&y, n * sizeof(float));
    cudaMemset(x, 0, n * sizeof(float));
    cudaStream_t st;
    cudaStreamCreate(&st);
    piquant_context_t* ctx = piquant_context_create(1);
    piquant_quantize(ctx, x, PIQUANT_DTYPE_F32, q, PIQUANT_DTYPE_UINT8, n, 0.01f, 128, PIQUANT_NEAREST);
    cudaDeviceSynchronize();
    printf("host time per call, numel = %zu (best of 7 x 2000 calls)\n", n);
    printf("  empty kernel <<<1,32>>>                          %6.2f us\n", per_call_us([&] { empty_kernel<<<1, 32, 0, st>>>(); }));
    printf("  piquant_quantize (reference ABI)                 %6.2f us\n",
           per_call_us([&] { piquant_quantize(ctx, x, PIQUANT_DTYPE_F32, q, PIQUANT_DTYPE_UINT8, n, 0.01f, 128, PIQUANT_NEAREST); }));
    printf("  piquant_cuda_quantize_on_stream                  %6.2f us\n",
           per_call_us([&] { piquant_cuda_quantize_on_stream(ctx, x, PIQUANT_DTYPE_F32, q, PIQUANT_DTYPE_UINT8, n, 0.01f, 128, PIQUANT_NEAREST, 0, st); }));
    printf("  piquant_dequantize (reference ABI)               %6.2f us\n",
           per_call_us([&] { piquant_dequantize(ctx, q, PIQUANT_DTYPE_UINT8, y, PIQUANT_DTYPE_F32, n, 0.01f, 128, PIQUANT_REDUCE_OP_ADD); }));
    printf("  piquant_cuda_dequantize_on_stream                %6.2f us\n",
           per_call_us([&] { piquant_cuda_dequantize_on_stream(ctx, q, PIQUANT_DTYPE_UINT8, y, PIQUANT_DTYPE_F32, n, 0.01f, 128, PIQUANT_REDUCE_OP_ADD, 0, st); }));
    piquant_context_destroy(ctx);
    return 0;
}
