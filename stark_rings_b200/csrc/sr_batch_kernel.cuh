// Batched element-wise kernel: one thread per ring element, one CTA per tile of T elements,
// persistent grid-stride loop over tiles.  See sr_tile.cuh for the HBM <-> shared staging.
#pragma once
#include <cuda_runtime.h>

#include "sr_launch.cuh"
#include "sr_tile.cuh"

namespace sr {

template <class R, int OP, int T, int MINB>
__global__ void __launch_bounds__(T, MINB)
batch_kernel(const u64* a, const u64* b, u64* out, size_t n) {
    extern __shared__ uint4 smem_raw[];
    u32* sA = reinterpret_cast<u32*>(smem_raw);
    u32* sB = sA + T * R::ROW;
    const size_t ntiles = (n + T - 1) / T;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t e0 = tile * T;
        const int ne = (n - e0 < (size_t)T) ? (int)(n - e0) : T;
        stage_in<R, T>(sA, a + e0 * R::WORDS64, ne);
        if (OP == OP_NTT_MUL || OP == OP_RING_MUL) stage_in<R, T>(sB, b + e0 * R::WORDS64, ne);
        __syncthreads();
        // rows >= ne of a ragged last tile hold stale data: transforming them is harmless (straight-line integer
        // code, never stored) and keeps every thread on the same path, which the phase barriers of gl_ring.cuh need
        {
            u32* rowA = sA + threadIdx.x * R::ROW;
            u32* rowB = sB + threadIdx.x * R::ROW;
            if (OP == OP_CRT) R::op_crt(rowA);
            if (OP == OP_ICRT) R::op_icrt(rowA);
            if (OP == OP_NTT_MUL) R::op_ntt_mul(rowA, rowB);
            if (OP == OP_RING_MUL) R::op_ring_mul(rowA, rowB);
        }
        __syncthreads();
        stage_out<R, T>(out + e0 * R::WORDS64, sA, ne);
        __syncthreads();
    }
}

template <class R, int OP, int T, int MINB>
cudaError_t launch_batch_op(const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    auto kern = batch_kernel<R, OP, T, MINB>;
    const bool two = (OP == OP_NTT_MUL || OP == OP_RING_MUL);
    const size_t smem = (size_t)(two ? 2 : 1) * T * R::ROW * sizeof(u32);
    static KernelCache cache;  // per instantiation, per device
    int blocks_per_sm = 0;
    cudaError_t e = cache.configure(kern, T, smem, &blocks_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (n + T - 1) / T;
    if (ntiles == 0) return cudaSuccess;
    size_t grid = (size_t)sms * blocks_per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, T, smem, st>>>(a, b, out, n);
    return cudaGetLastError();
}

}  // namespace sr
