// BabyBear batch kernels (instantiations).
#include "bb_half.cuh"
#include "bb_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

// ---- fused ring multiplication, two threads per element (bb_half.cuh) -------------------------
__constant__ bb::HalfConsts BB_HALF_C[2] = {bb::half_consts(0), bb::half_consts(1)};

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
#ifdef SR_BB_L2_PREFETCH  // tried in round 1: no measurable effect (profiles/r01_tuning.md)
        {   // pull the CTA's next tile into L2 while this one is being multiplied
            const size_t nt = tile + gridDim.x;
            if (nt < ntiles) {
                const size_t ne2 = (n - nt * TE < (size_t)TE) ? (n - nt * TE) : (size_t)TE;
                const size_t lines = ne2 * R::WORDS64 * 8 / 128;  // 576 B per element: 4.5 lines
                const char* pa = reinterpret_cast<const char*>(a + nt * TE * R::WORDS64);
                const char* pb = reinterpret_cast<const char*>(b + nt * TE * R::WORDS64);
                for (size_t l = threadIdx.x; l < lines; l += T) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + l * 128));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + l * 128));
                }
            }
        }
#endif
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

// Warp-autonomous flavour: each warp stages, multiplies and stores its own 16-element tiles.
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
bb_ring_mul_half_warp_kernel(const u64* a, const u64* b, u64* out, size_t n) {
    typedef BBPolicy R;
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, h = lane >> 4, warp = threadIdx.x >> 5;
    u32* sA = reinterpret_cast<u32*>(smem_raw) + warp * 2 * 16 * R::ROW;
    u32* sB = sA + 16 * R::ROW;
    const bb::HalfConsts K = BB_HALF_C[h];
    const size_t ntiles = (n + 15) / 16;
    const size_t wstride = (size_t)gridDim.x * WARPS;
    for (size_t tile = (size_t)blockIdx.x * WARPS + warp; tile < ntiles; tile += wstride) {
        const size_t e0 = tile * 16;
        const int ne = (n - e0 < 16) ? (int)(n - e0) : 16;
        stage_in_warp<R>(sA, a + e0 * R::WORDS64, ne);
        stage_in_warp<R>(sB, b + e0 * R::WORDS64, ne);
        __syncwarp();
        u32 A[36], B[36], Y[36];
        bb::half_crt(A, sA + (lane & 15) * R::ROW, K);
        bb::half_crt(B, sB + (lane & 15) * R::ROW, K);
        bb::half_slots(B, A, K);
        bb::half_icrt_local(B, K);
#pragma unroll
        for (int i = 0; i < 36; i++) Y[i] = __shfl_xor_sync(0xffffffffu, B[i], 16);
        bb::half_final(A, B, Y, K);
        __syncwarp();
        u32* dst = sA + (lane & 15) * R::ROW + 36 * h;
#pragma unroll
        for (int i = 0; i < 9; i++)
            *reinterpret_cast<uint4*>(dst + 4 * i) = make_uint4(A[4 * i], A[4 * i + 1], A[4 * i + 2], A[4 * i + 3]);
        __syncwarp();
        stage_out_warp<R>(out + e0 * R::WORDS64, sA, ne);
        __syncwarp();
    }
}

template <int WARPS, int MINB>
static cudaError_t launch_ring_mul_half_warp(const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    auto kern = bb_ring_mul_half_warp_kernel<WARPS, MINB>;
    const size_t smem = (size_t)2 * WARPS * 16 * BBPolicy::ROW * sizeof(u32);
    static KernelCache cache;  // per instantiation, per device
    int blocks_per_sm = 0;
    cudaError_t e = cache.configure(kern, WARPS * 32, smem, &blocks_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (n + 15) / 16;
    if (ntiles == 0) return cudaSuccess;
    size_t grid = (size_t)sms * blocks_per_sm;
    const size_t need = (ntiles + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(a, b, out, n);
    return cudaGetLastError();
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

#ifndef SR_BB_T
#define SR_BB_T 128
#endif
#ifndef SR_BB_MINB
#define SR_BB_MINB 2
#endif

// Tuning (B200, n = 2^22, see profiles/r01_tuning.md): two-threads-per-element, 2 warps per CTA,
// 8 CTAs per SM (16 warps, 127 registers) is the fastest fused ring-mul configuration.
#if !defined(SR_BB_HALF) && !defined(SR_BB_HALFW) && !defined(SR_BB_BLOCKSYNC) && !defined(SR_BB_WARPTILE)
#define SR_BB_HALF
#endif
#ifndef SR_BB_HALF_WARPS
#define SR_BB_HALF_WARPS 2
#endif
#ifndef SR_BB_HALF_MINB
#define SR_BB_HALF_MINB 8
#endif
#ifndef SR_BB_WARPS
#define SR_BB_WARPS 8
#endif
#ifndef SR_BB_WMINB
#define SR_BB_WMINB 1
#endif

cudaError_t bb_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_batch_op<BBPolicy, OP_CRT, SR_BB_T, 3>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<BBPolicy, OP_ICRT, SR_BB_T, 3>(a, b, out, n, st, sms);
#ifndef SR_BB_NM_T
#define SR_BB_NM_T 64
#endif
#ifndef SR_BB_NM_MINB
#define SR_BB_NM_MINB 4
#endif
    case OP_NTT_MUL: return launch_batch_op<BBPolicy, OP_NTT_MUL, SR_BB_NM_T, SR_BB_NM_MINB>(a, b, out, n, st, sms);  // 0.88 vs 0.80 of roofline at T=128
#if defined(SR_BB_HALFW)
    case OP_RING_MUL: return launch_ring_mul_half_warp<SR_BB_HALF_WARPS, SR_BB_HALF_MINB>(a, b, out, n, st, sms);
#elif defined(SR_BB_HALF)
    case OP_RING_MUL: return launch_ring_mul_half<SR_BB_HALF_WARPS, SR_BB_HALF_MINB>(a, b, out, n, st, sms);
#elif defined(SR_BB_BLOCKSYNC)
    case OP_RING_MUL: return launch_batch_op<BBPolicy, OP_RING_MUL, SR_BB_T, SR_BB_MINB>(a, b, out, n, st, sms);
#else
    case OP_RING_MUL: return launch_batch_op_warp<BBPolicy, OP_RING_MUL, SR_BB_WARPS, SR_BB_WMINB>(a, b, out, n, st, sms);
#endif
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
