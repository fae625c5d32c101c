// Microbenchmark: throughput of the 32-bit modular-multiplication idioms on one B200 (which of them the integer
// pipes execute fastest).  Eight independent chains per thread, eight CTAs of 256 threads per SM, CUDA events.
//   imad_hi      IMAD.HI.U32
//   mont         Montgomery multiplication by a constant: t = x * w (wide), m = lo(t) * np, r = hi(m * p + t), min(r, r - p)
//   shoup        Shoup multiplication by a constant: q = hi(x * w'), r = x * w - q * p (two 32-bit IMADs), min(r, r - p)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe_peak int_pipe_peak.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;
constexpr uint32_t P = 0x78000001u, NP = 0x77FFFFFFu;

__global__ void k_hi(uint32_t* out, uint32_t a) {
    uint32_t x[CH];
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 2654435761u + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = __umulhi(x[i], a) + x[i];
    uint32_t s = 0;
    for (int i = 0; i < CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mont(uint32_t* out, uint32_t w) {
    uint32_t x[CH];
    for (int i = 0; i < CH; i++) x[i] = (threadIdx.x * 2654435761u + i) % P;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const uint64_t t = (uint64_t)x[i] * w;
            const uint32_t m = (uint32_t)t * NP;
            const uint32_t r = (uint32_t)(((uint64_t)m * P + t) >> 32);
            x[i] = min(r, r - P);
        }
    uint32_t s = 0;
    for (int i = 0; i < CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shoup(uint32_t* out, uint32_t w, uint32_t wp) {
    uint32_t x[CH];
    for (int i = 0; i < CH; i++) x[i] = (threadIdx.x * 2654435761u + i) % P;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const uint32_t q = __umulhi(x[i], wp);
            const uint32_t r = x[i] * w - q * P;
            x[i] = min(r, r - P);
        }
    uint32_t s = 0;
    for (int i = 0; i < CH; i++) s += x[i];
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
    const uint32_t w = 0x1e04309eu, wp = (uint32_t)(((unsigned __int128)w << 32) / P);
    float t1 = time([&] { k_hi<<<blocks, threads>>>((uint32_t*)buf, 0x9E3779B9u); });
    float t2 = time([&] { k_mont<<<blocks, threads>>>((uint32_t*)buf, w); });
    float t3 = time([&] { k_shoup<<<blocks, threads>>>((uint32_t*)buf, w, wp); });
    printf("{\"sms\": %d, \"imad_hi_Tops\": %.3f, \"mont_mulc_Tmul\": %.3f, \"shoup_mulc_Tmul\": %.3f}\n",
           p.multiProcessorCount, ops / t1 / 1e9, ops / t2 / 1e9, ops / t3 / 1e9);
    return 0;
}
