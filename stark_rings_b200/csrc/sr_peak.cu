// Measured integer multiply-add peaks of the device the context runs on: the roofline denominator of the kernels
// that are bound by the integer pipe (Starknet prime: 252-bit Montgomery arithmetic is made of 32 x 32 + 64 -> 64
// multiply-adds, SASS IMAD.WIDE.U32[.X]).  Three dependent-chain microbenchmarks, eight independent chains per
// thread, eight CTAs of 256 threads per SM: 32-bit IMAD, IMAD.WIDE.U32, and the carry-chained IMAD.WIDE.U32.X form
// the multi-limb kernels use.  Timed with CUDA events on the context's stream, best of five.
#include <cuda_runtime.h>

#include <cstdint>

namespace sr {

constexpr int PK_ITERS = 4096, PK_CH = 8;

__global__ void peak_imad_kernel(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[PK_CH];
    for (int i = 0; i < PK_CH; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < PK_ITERS; it++)
#pragma unroll
        for (int i = 0; i < PK_CH; i++) x[i] = x[i] * a + b;
    uint32_t s = 0;
    for (int i = 0; i < PK_CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void peak_wide_kernel(uint64_t* out, uint32_t a) {
    uint64_t x[PK_CH];
    for (int i = 0; i < PK_CH; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < PK_ITERS; it++)
#pragma unroll
        for (int i = 0; i < PK_CH; i++) x[i] = (uint64_t)(uint32_t)x[i] * a + x[i];
    uint64_t s = 0;
    for (int i = 0; i < PK_CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void peak_wide_carry_kernel(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t lo[PK_CH], hi[PK_CH], top[PK_CH];
    for (int i = 0; i < PK_CH; i++) { lo[i] = threadIdx.x + i; hi[i] = i; top[i] = 0; }
    for (int it = 0; it < PK_ITERS; it++)
#pragma unroll
        for (int i = 0; i < PK_CH; i++)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                         : "+r"(lo[i]), "+r"(hi[i]), "+r"(top[i]) : "r"(a), "r"(b + i));
    uint32_t s = 0;
    for (int i = 0; i < PK_CH; i++) s += lo[i] ^ hi[i] ^ top[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// tops[0..2] = 10^12 operations per second: IMAD, IMAD.WIDE.U32, IMAD.WIDE.U32.X (+ the addc of the chain)
cudaError_t imad_peak_measure(int sms, cudaStream_t st, double* tops) {
    const int blocks = sms * 8, threads = 256;
    void* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)blocks * threads * 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double ops = (double)blocks * threads * PK_ITERS * PK_CH;
    for (int which = 0; which < 3 && e == cudaSuccess; which++) {
        float best = 1e30f;
        for (int r = 0; r < 6 && e == cudaSuccess; r++) {  // first pass = warm-up
            cudaEventRecord(e0, st);
            if (which == 0) peak_imad_kernel<<<blocks, threads, 0, st>>>((uint32_t*)buf, 0x9E3779B9u, 12345u);
            if (which == 1) peak_wide_kernel<<<blocks, threads, 0, st>>>((uint64_t*)buf, 0x9E3779B9u);
            if (which == 2) peak_wide_carry_kernel<<<blocks, threads, 0, st>>>((uint32_t*)buf, 0x9E3779B9u, 777u);
            cudaEventRecord(e1, st);
            e = cudaEventSynchronize(e1);
            float ms = 0;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        tops[which] = ops / best / 1e9;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return e;
}

}  // namespace sr
