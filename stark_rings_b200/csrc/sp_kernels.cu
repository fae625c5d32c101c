// Starknet-prime batch kernels (instantiations).
#include "sp_half.cuh"
#include "sp_quad.cuh"
#include "sp_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

// ---- two threads per element (sp_half.cuh) -----------------------------------------------------
SR_D void sp_exchange(sp::Fe (&recv)[4], const sp::Fe (&send)[4]) {
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < sp::L; i++) recv[j].v[i] = __shfl_xor_sync(0xffffffffu, send[j].v[i], 16);
}
// parity layout (coefficients h + 2j) -> CRT positions 8h .. 8h+7
SR_D void sp_half_crt_row(sp::Fe (&out)[8], const u32* row, int h) {
    sp::Fe c[8], send[4], recv[4];
#pragma unroll
    for (int j = 0; j < 8; j++) SPPolicy::load_fe(c[j], row, h + 2 * j);
    sp::half_crt_local(c);
    sp::half_crt_send(send, c, h);
    sp_exchange(recv, send);
    sp::half_crt_cross(out, c, recv, h);
}
// CRT positions 8h .. 8h+7 -> parity layout
SR_D void sp_half_icrt_regs(sp::Fe (&c)[8], sp::Fe (&p)[8], int h) {
    sp::Fe send[4], recv[4];
    sp::half_icrt_first(p, h);
    sp::half_icrt_send(send, p, h);
    sp_exchange(recv, send);
    sp::half_icrt_gather(c, p, recv, h);
    sp::half_icrt_local(c);
}

template <int OP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
sp_half_kernel(const u64* a, const u64* b, u64* out, size_t n) {
    typedef SPPolicy R;
    constexpr int T = WARPS * 32, TE = WARPS * 16;
    extern __shared__ uint4 smem_raw[];
    u32* sA = reinterpret_cast<u32*>(smem_raw);
    u32* sB = sA + TE * R::ROW;
    const int lane = threadIdx.x & 31, h = lane >> 4;
    const int el = (threadIdx.x >> 5) * 16 + (lane & 15);
    const size_t ntiles = (n + TE - 1) / TE;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t e0 = tile * TE;
        const int ne = (n - e0 < (size_t)TE) ? (int)(n - e0) : TE;
        stage_in<R, T>(sA, a + e0 * R::WORDS64, ne);
        if (OP == OP_RING_MUL) stage_in<R, T>(sB, b + e0 * R::WORDS64, ne);
        __syncthreads();
        u32* rowA = sA + el * R::ROW;
        u32* rowB = sB + el * R::ROW;
        sp::Fe x[8], y[8];
        if (OP == OP_CRT) {
            sp_half_crt_row(x, rowA, h);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) R::store_fe(rowA, 8 * h + q, x[q]);
        } else if (OP == OP_ICRT) {
#pragma unroll
            for (int q = 0; q < 8; q++) R::load_fe(x[q], rowA, 8 * h + q);
            sp_half_icrt_regs(y, x, h);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) R::store_fe(rowA, h + 2 * j, y[j]);
        } else {  // fused ring multiplication
#if defined(SR_SP_RM_LOOP)
            // ONE copy of the forward transform, run for both operands in a 2-trip loop whose trip count is opaque to
            // the compiler (same code-size discipline as gl_fused6.cuh): crt(a) is parked in the row, crt(b) stays in x
            int trips = 2;
            asm volatile("" : "+r"(trips));
            const ptrdiff_t delta = rowB - rowA;
#pragma unroll 1
            for (int k = 0; k < trips; k++) {
                u32* row = rowA + k * delta;
                sp_half_crt_row(x, row, h);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++) R::store_fe(row, 8 * h + q, x[q]);
            }
#else
            sp_half_crt_row(x, rowA, h);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) R::store_fe(rowA, 8 * h + q, x[q]);  // park crt(a)
            sp_half_crt_row(x, rowB, h);
#endif
#if defined(SR_SP_RM_LOOP)
            // the eight slot products as a real loop over the rows (crt(b) was stored by the second trip above)
#pragma unroll 1
            for (int q = 0; q < 8; q++) {
                sp::Fe s, t;
                R::load_fe(s, rowB, 8 * h + q);
                R::load_fe(t, rowA, 8 * h + q);
                sp::mont_mul(s, s, t);
                R::store_fe(rowA, 8 * h + q, s);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) R::load_fe(x[q], rowA, 8 * h + q);
#else
#pragma unroll
            for (int q = 0; q < 8; q++) {
                sp::Fe t;
                R::load_fe(t, rowA, 8 * h + q);
                sp::mont_mul(x[q], x[q], t);
            }
#endif
            sp_half_icrt_regs(y, x, h);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) R::store_fe(rowA, h + 2 * j, y[j]);
        }
        __syncthreads();
        stage_out<R, T>(out + e0 * R::WORDS64, sA, ne);
        __syncthreads();
    }
}

template <int OP, int WARPS, int MINB>
static cudaError_t launch_sp_half(const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    auto kern = sp_half_kernel<OP, WARPS, MINB>;
    constexpr int TE = WARPS * 16;
    const size_t smem = (size_t)(OP == OP_RING_MUL ? 2 : 1) * TE * SPPolicy::ROW * sizeof(u32);
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
    for (int i = threadIdx.x; i < 256; i += T) wtab[i] = sp::SP_WTAB.w[i >> 3][i & 7];
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
                    sp::quad_fwd_stage(rowA + k * delta, wtab, s, t);
                    __syncwarp();
                }
            }
        }
        if (OP == OP_RING_MUL || OP == OP_NTT_MUL) {
            sp::quad_slots(rowA, rowB, t);
            __syncwarp();
        }
        if (OP == OP_ICRT || OP == OP_RING_MUL) {
#pragma unroll 1
            for (int s = 0; s < 3; s++) {
                sp::quad_inv_stage(rowA, wtab, s, t);
                __syncwarp();
            }
            sp::quad_inv_last(rowA, t);
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
// configuration for all three ops; the two-thread kernels (sp_half.cuh) remain selectable with -DSR_SP_HALF.
#if !defined(SR_SP_HALF) && !defined(SR_SP_NO_HALF)
#define SR_SP_QUAD
#endif
#ifndef SR_SP_QUAD_WARPS
#define SR_SP_QUAD_WARPS 2
#endif
#ifndef SR_SP_QUAD_MINB
#define SR_SP_QUAD_MINB 12
#endif

#ifndef SR_SP_HALF_WARPS
#define SR_SP_HALF_WARPS 2
#endif
#ifndef SR_SP_HALF_MINB
#define SR_SP_HALF_MINB 6
#endif

#ifndef SR_SP_T
#define SR_SP_T 64
#endif

cudaError_t sp_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
#if defined(SR_SP_QUAD)
    case OP_CRT: return launch_sp_quad<OP_CRT, SR_SP_QUAD_WARPS, SR_SP_QUAD_MINB>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_sp_quad<OP_ICRT, SR_SP_QUAD_WARPS, SR_SP_QUAD_MINB>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_sp_quad<OP_RING_MUL, SR_SP_QUAD_WARPS, SR_SP_QUAD_MINB>(a, b, out, n, st, sms);
#elif !defined(SR_SP_NO_HALF)
    case OP_CRT: return launch_sp_half<OP_CRT, SR_SP_HALF_WARPS, SR_SP_HALF_MINB>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_sp_half<OP_ICRT, SR_SP_HALF_WARPS, SR_SP_HALF_MINB>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_sp_half<OP_RING_MUL, SR_SP_HALF_WARPS, SR_SP_HALF_MINB>(a, b, out, n, st, sms);
#else
    case OP_CRT: return launch_batch_op<SPPolicy, OP_CRT, SR_SP_T, 4>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<SPPolicy, OP_ICRT, SR_SP_T, 4>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_batch_op<SPPolicy, OP_RING_MUL, SR_SP_T, 3>(a, b, out, n, st, sms);
#endif
#if defined(SR_SP_QUAD) && !defined(SR_SP_NTTMUL_BATCH)
    case OP_NTT_MUL: return launch_sp_quad<OP_NTT_MUL, SR_SP_QUAD_WARPS, SR_SP_QUAD_MINB>(a, b, out, n, st, sms);
#else
    case OP_NTT_MUL: return launch_batch_op<SPPolicy, OP_NTT_MUL, SR_SP_T, 4>(a, b, out, n, st, sms);
#endif
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
