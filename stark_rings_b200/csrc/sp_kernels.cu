// Starknet-prime batch kernels (instantiations).
#include "sp_quad.cuh"
#include "sp_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

// the 32 roots in Montgomery form (copied into shared memory by every CTA)
__constant__ sp::RootTable SP_WTAB = {SR_SP_ROOTS_MONT};

// ---- four threads per element, stage by stage over the shared-memory row (sp_quad.cuh) ------------------
template <int OP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
sp_quad_kernel(const u64* a, const u64* b, u64* out, size_t n) {
    typedef SPPolicy R;
    constexpr int T = WARPS * 32, TE = WARPS * 8;
    extern __shared__ uint4 smem_raw[];
    u32* wtab = reinterpret_cast<u32*>(smem_raw);  // 32 roots x 8 limbs
    u32* sA = wtab + 256;
    u32* sB = sA + TE * R::ROW;
    for (int i = threadIdx.x; i < 256; i += T) wtab[i] = SP_WTAB.w[i >> 3][i & 7];
    // lanes 8t .. 8t + 7 are thread t of the warp's eight elements: a quarter-warp then touches the same coefficient
    // of eight consecutive rows, which the 16-byte row padding makes conflict-free in every stage
    const int lane = threadIdx.x & 31, t = lane >> 3;
    const int el = (threadIdx.x >> 5) * 8 + (lane & 7);
    u32* rowA = sA + el * R::ROW;
    u32* rowB = sB + el * R::ROW;
    const size_t ntiles = (n + TE - 1) / TE;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t e0 = tile * TE;
        const int ne = (n - e0 < (size_t)TE) ? (int)(n - e0) : TE;
        stage_in<R, T>(sA, a + e0 * R::WORDS64, ne);
        if (OP == OP_RING_MUL || OP == OP_NTT_MUL) stage_in<R, T>(sB, b + e0 * R::WORDS64, ne);
        __syncthreads();
        // rows >= ne of a ragged last tile hold stale data: transforming them is harmless and keeps the warp converged
        if (OP == OP_CRT || OP == OP_RING_MUL) {
            int trips = (OP == OP_RING_MUL) ? 2 : 1;
            asm volatile("" : "+r"(trips));  // opaque: ONE copy of the forward transform for both operands
            const ptrdiff_t delta = rowB - rowA;
#pragma unroll 1
            for (int k = 0; k < trips; k++) {
#pragma unroll 1
                for (int s = 0; s < 4; s++) {
                    sp::quad_fwd_stage<true, OP == OP_CRT>(rowA + k * delta, wtab, s, t);  // unreduced; the CRT canonicalises its last stage
                    __syncwarp();
                }
            }
        }
        if (OP == OP_RING_MUL || OP == OP_NTT_MUL) {
            sp::quad_slots<OP == OP_RING_MUL>(rowA, rowB, t);
            __syncwarp();
        }
        if (OP == OP_ICRT || OP == OP_RING_MUL) {
#pragma unroll 1
            for (int s = 0; s < 3; s++) {
                sp::quad_inv_stage<true>(rowA, wtab, s, t);  // unreduced until the last stage
                __syncwarp();
            }
            sp::quad_inv_last<true, OP == OP_ICRT>(rowA, t);
        }
        __syncthreads();
        stage_out<R, T>(out + e0 * R::WORDS64, sA, ne);
        __syncthreads();
    }
}

template <int OP, int WARPS, int MINB>
static cudaError_t launch_sp_quad(const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    auto kern = sp_quad_kernel<OP, WARPS, MINB>;
    constexpr int TE = WARPS * 8;
    const size_t smem = 1024 + (size_t)(OP == OP_RING_MUL || OP == OP_NTT_MUL ? 2 : 1) * TE * SPPolicy::ROW * sizeof(u32);
    static KernelCache cache;  // per instantiation, per device
    int blocks_per_sm = 0;
    cudaError_t e = cache.configure(kern, WARPS * 32, smem, &blocks_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (n + TE - 1) / TE;
    if (ntiles == 0) return cudaSuccess;
    size_t grid = (size_t)sms * blocks_per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(a, b, out, n);
    return cudaGetLastError();
}

// Tuning (B200, n = 2^20, profiles/r01b_gl_ncu.md): 2 warps x 12 CTAs = 24 warps per SM (80 registers) is the fastest
// configuration; with the unreduced arithmetic of round 2 the ICRT gains 2% from 14 CTAs (72 registers; n = 2^22:
// 2.404 against 2.451 ms), the CRT loses 2% (2.144 against 2.100 ms), 16 CTAs (64 registers) are no better.
cudaError_t sp_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_sp_quad<OP_CRT, 2, 12>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_sp_quad<OP_ICRT, 2, 14>(a, b, out, n, st, sms);
    case OP_NTT_MUL: return launch_sp_quad<OP_NTT_MUL, 2, 12>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_sp_quad<OP_RING_MUL, 2, 12>(a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
