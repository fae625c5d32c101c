// Microbenchmark: sustained integer multiply-add throughput of one B200 (the roofline denominator for the
// Starknet-prime kernels).  Measures IMAD (32-bit), IMAD.WIDE.U32 (32x32+64) and the carry-chained
// IMAD.WIDE.U32.X form with enough independent chains per thread to saturate the fma pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu ; run: ./imad_peak
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[CH];
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = x[i] * a + b;
    uint32_t s = 0;
    for (int i = 0; i < CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_wide(uint64_t* out, uint32_t a) {
    uint64_t x[CH];
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = (uint64_t)(uint32_t)x[i] * a + x[i];
    uint64_t s = 0;
    for (int i = 0; i < CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_wide_carry(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t lo[CH], hi[CH], top[CH];
    for (int i = 0; i < CH; i++) { lo[i] = threadIdx.x + i; hi[i] = i; top[i] = 0; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                         : "+r"(lo[i]), "+r"(hi[i]), "+r"(top[i]) : "r"(a), "r"(b + i));
    uint32_t s = 0;
    for (int i = 0; i < CH; i++) s += lo[i] ^ hi[i] ^ top[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto launch) {
        launch(); cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        return best;
    };
    const double ops = (double)blocks * threads * ITERS * CH;
    float t1 = time([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 0x9E3779B9u, 12345u); });
    float t2 = time([&] { k_wide<<<blocks, threads>>>((uint64_t*)buf, 0x9E3779B9u); });
    float t3 = time([&] { k_wide_carry<<<blocks, threads>>>((uint32_t*)buf, 0x9E3779B9u, 777u); });
    printf("{\"sms\": %d, \"imad32_Tops\": %.3f, \"imad_wide_Tops\": %.3f, \"imad_wide_carry_Tops\": %.3f, "
           "\"imad32_per_clk_per_sm_at_1965MHz\": %.1f, \"imad_wide_per_clk_per_sm_at_1965MHz\": %.1f}\n",
           p.multiProcessorCount, ops / t1 / 1e9, ops / t2 / 1e9, ops / t3 / 1e9,
           ops / t1 / 1e3 / p.multiProcessorCount / 1.965e6, ops / t2 / 1e3 / p.multiProcessorCount / 1.965e6);
    return 0;
}
