// BabyBear batch kernels (instantiations).
#include "bb_half.cuh"
#include "bb_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

// ---- fused ring multiplication, two threads per element (bb_half.cuh) -------------------------
__constant__ bb::HalfConsts BB_HALF_C[2] = {bb::half_consts(0), bb::half_consts(1)};

#ifndef SR_BB_HALF_WARPS
#define SR_BB_HALF_WARPS 2
#endif
#ifndef SR_BB_HALF_MINB
#define SR_BB_HALF_MINB 8
#endif
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
bb_ring_mul_half_kernel(const u64* a, const u64* b, u64* out, size_t n) {
    typedef BBPolicy R;
    constexpr int T = WARPS * 32, TE = WARPS * 16;  // threads / elements per tile
    extern __shared__ uint4 smem_raw[];
    u32* sA = reinterpret_cast<u32*>(smem_raw);
    u32* sB = sA + TE * R::ROW;
    const int lane = threadIdx.x & 31, h = lane >> 4;
    const int el = (threadIdx.x >> 5) * 16 + (lane & 15);
    const bb::HalfConsts K = BB_HALF_C[h];
    const size_t ntiles = (n + TE - 1) / TE;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t e0 = tile * TE;
        const int ne = (n - e0 < (size_t)TE) ? (int)(n - e0) : TE;
        stage_in<R, T>(sA, a + e0 * R::WORDS64, ne);
        stage_in<R, T>(sB, b + e0 * R::WORDS64, ne);
        __syncthreads();
        u32 A[36], B[36], Y[36];
        bb::half_crt(A, sA + el * R::ROW, K);
        bb::half_crt(B, sB + el * R::ROW, K);
        bb::half_slots(B, A, K);
        bb::half_icrt_local(B, K);
#pragma unroll
        for (int i = 0; i < 36; i++) Y[i] = __shfl_xor_sync(0xffffffffu, B[i], 16);
        bb::half_final(A, B, Y, K);
        __syncwarp();  // both halves have consumed the input row before it is overwritten
        u32* dst = sA + el * R::ROW + 36 * h;
#pragma unroll
        for (int i = 0; i < 9; i++)
            *reinterpret_cast<uint4*>(dst + 4 * i) = make_uint4(A[4 * i], A[4 * i + 1], A[4 * i + 2], A[4 * i + 3]);
        __syncthreads();
        stage_out<R, T>(out + e0 * R::WORDS64, sA, ne);
        __syncthreads();
    }
}

template <int WARPS, int MINB>
static cudaError_t launch_ring_mul_half(const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    auto kern = bb_ring_mul_half_kernel<WARPS, MINB>;
    constexpr int TE = WARPS * 16;
    const size_t smem = (size_t)2 * TE * BBPolicy::ROW * sizeof(u32);
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

// Tuning (B200, n = 2^22, see profiles/r01_tuning.md): CRT / ICRT one thread per element, 128-thread CTAs, 3 per SM;
// NTT-form product 64-thread CTAs, 4 per SM (0.88 vs 0.80 of the roofline at 128); fused ring mul two threads per
// element, 2 warps per CTA, 8 CTAs per SM (16 warps, 126 registers).  Round 2, with the Shoup twiddle multiplications
// (0.817): 96-register builds with 9 / 10 CTAs per SM 0.651, 4 warps x 5 CTAs 0.749, 4 x 4 0.811; ONE copy of the forward
// transform run for both operands in a two-trip loop (11 KB less code, 28 bytes spilled) 0.797.
cudaError_t bb_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_batch_op<BBPolicy, OP_CRT, 128, 3>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<BBPolicy, OP_ICRT, 128, 3>(a, b, out, n, st, sms);
    case OP_NTT_MUL: return launch_batch_op<BBPolicy, OP_NTT_MUL, 64, 4>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_ring_mul_half<SR_BB_HALF_WARPS, SR_BB_HALF_MINB>(a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
