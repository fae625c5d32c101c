// Ring matrix x vector product y_i = sum_j A[i][j] * v[j] over NTT-form elements
// (reference: linear_algebra/src/matrix.rs:168-178 with R = RqNTT; the inner product is
// ntt_form.rs:521-536 (slot-wise Mul) folded with ntt_form.rs:588-601 (Add) from ZERO).
//
// The product is independent per CRT slot, so the unit of work is one slot of one column: thread t
// owns slot (t mod SLOTS) of columns t / SLOTS, t / SLOTS + stride, ...; consecutive threads read
// consecutive slots, i.e. each warp reads one contiguous span of a row.  Every thread keeps one
// accumulator per matrix row (RB rows per pass), CTAs reduce per slot index through shared memory
// and write one partial element per row to scratch; a second small kernel adds the partials mod p
// in a fixed order, so the result is deterministic.  The same second kernel is the rank-0 modular
// sum of the multi-GPU commitment (sr_modsum_partials).
#include <cuda_runtime.h>

#define SR_GL_EPS_ON_ALU  // these kernels are bound by the multiply-add pipe: see gl_ring.cuh plus_eps_if

#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "sp_ring.cuh"
#include "sr_slots.cuh"
#include "sr_tma.cuh"

namespace sr {

// ---- Goldilocks: lazy accumulation ---------------------------------------------------------------
// sum_j a_j * x_j of 64 x 64 -> 128-bit products is kept UNREDUCED in a 160-bit accumulator held as two
// interleaved carry-save halves (E: limbs 0..4 takes lo*lo and hi*hi, O: limbs 1..3 takes the two
// cross products), so that every partial product is one IMAD.WIDE.U32 with carry and no modular
// reduction happens inside the column loop.  One reduction per thread at the end.
struct GLAcc {
    u32 e0, e1, e2, e3, e4, o1, o2, o3;
};
SR_D void gl_acc_zero(GLAcc& A) { A.e0 = A.e1 = A.e2 = A.e3 = A.e4 = A.o1 = A.o2 = A.o3 = 0; }
SR_D void gl_acc_mad(GLAcc& A, u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    asm(
        "mad.lo.cc.u32   %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32  %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32  %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32  %3, %6, %8, %3;\n\t"
        "addc.u32        %4, %4, 0;\n\t"
        : "+r"(A.e0), "+r"(A.e1), "+r"(A.e2), "+r"(A.e3), "+r"(A.e4)
        : "r"(al), "r"(ah), "r"(bl), "r"(bh));
    asm(
        "mad.lo.cc.u32   %0, %3, %6, %0;\n\t"
        "madc.hi.cc.u32  %1, %3, %6, %1;\n\t"
        "addc.u32        %2, %2, 0;\n\t"
        "mad.lo.cc.u32   %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32  %1, %4, %5, %1;\n\t"
        "addc.u32        %2, %2, 0;\n\t"
        : "+r"(A.o1), "+r"(A.o2), "+r"(A.o3)
        : "r"(al), "r"(ah), "r"(bl), "r"(bh));
}
// canonical residue of the accumulated value times 2^POST
template <int POST>
SR_D u64 gl_acc_reduce(const GLAcc& A) {
    // merge: limbs l0..l5 of E + O * 2^32
    u64 c = (u64)A.e1 + A.o1;
    const u32 l0 = A.e0, l1 = (u32)c;
    c = (c >> 32) + (u64)A.e2 + A.o2;
    const u32 l2 = (u32)c;
    c = (c >> 32) + (u64)A.e3 + A.o3;
    const u32 l3 = (u32)c;
    c = (c >> 32) + (u64)A.e4;
    const u32 l4 = (u32)c, l5 = (u32)(c >> 32);
    // 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32, 2^160 = 1 - 2^32 (mod p)
    u64 r = gl::reduce128((u64)l0 | ((u64)l1 << 32), (u64)l2 | ((u64)l3 << 32));
    r = gl::sub(r, (u64)l4 << 32);
    r = gl::add(r, (u64)l5);
    r = gl::sub(r, (u64)l5 << 32);
    return gl::canon(POST ? gl::mul_pow2<POST>(r) : r);
}

#ifndef SR_GLMV_MINB
#define SR_GLMV_MINB 2
#endif
#ifndef SR_GLMV_RB
#define SR_GLMV_RB 4
#endif
constexpr int GLMV_T = 128;
// Rows [row0, row0 + RB) of the product; same work split and partial layout as matvec_partial_kernel.
template <int RB>
__global__ void __launch_bounds__(GLMV_T, SR_GLMV_MINB)
gl_matvec_lazy_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                      const u64* __restrict__ v, u64* __restrict__ partial) {
    typedef GLSlot S;
    __shared__ S::Val red[GLMV_T];
    GLAcc acc[RB][3];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) gl_acc_zero(acc[r][k]);
    const u64* rp[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) rp[r] = (row0 + r < nrows) ? rows[row0 + r] : rows[nrows - 1];

    const size_t total = ncols * S::SLOTS;
    const size_t stride = (size_t)gridDim.x * GLMV_T;
    for (size_t g = (size_t)blockIdx.x * GLMV_T + threadIdx.x; g < total; g += stride) {
        u64 a[RB][3];
#pragma unroll
        for (int r = 0; r < RB; r++) {
            a[r][0] = __ldcs(rp[r] + g * 3);
            a[r][1] = __ldcs(rp[r] + g * 3 + 1);
            a[r][2] = __ldcs(rp[r] + g * 3 + 2);
        }
        const u64 x0 = v[g * 3], x1 = v[g * 3 + 1], x2 = v[g * 3 + 2];
        const u64 xr1 = gl::mul_pow2<gl::root_exp(1)>(x1), xr2 = gl::mul_pow2<gl::root_exp(1)>(x2);  // u^3 = r
#pragma unroll
        for (int r = 0; r < RB; r++) {
            gl_acc_mad(acc[r][0], a[r][0], x0);
            gl_acc_mad(acc[r][0], a[r][1], xr2);
            gl_acc_mad(acc[r][0], a[r][2], xr1);
            gl_acc_mad(acc[r][1], a[r][0], x1);
            gl_acc_mad(acc[r][1], a[r][1], x0);
            gl_acc_mad(acc[r][1], a[r][2], xr2);
            gl_acc_mad(acc[r][2], a[r][0], x2);
            gl_acc_mad(acc[r][2], a[r][1], x1);
            gl_acc_mad(acc[r][2], a[r][2], x0);
        }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        S::Val mine;  // Montgomery layout: the product of two raw limbs carries an extra 2^-64 = 2^128
#pragma unroll
        for (int k = 0; k < 3; k++) mine.c[k] = gl_acc_reduce<128>(acc[r][k]);
        red[threadIdx.x] = mine;
        __syncthreads();
        if (threadIdx.x < S::SLOTS) {
            S::Val s = red[threadIdx.x];
            for (int k = threadIdx.x + S::SLOTS; k < GLMV_T; k += S::SLOTS) S::acc(s, red[k]);
            S::store(partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + threadIdx.x * S::SLOT_U64, s);
        }
        __syncthreads();
    }
}

// TMA-pipelined flavour: the CTA streams chunks of GLTMA_T consecutive slots (3072 B) of v and of RB
// matrix rows through an NS-deep ring of shared-memory stages filled by 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx), so HBM requests are whole 128-byte lines issued far ahead of
// use and the per-thread 24-byte slot reads hit shared memory (stride 6 words: conflict-free LDS.64).
constexpr int GLTMA_T = 128, GLTMA_NS = 4;
// RB rows per launch, split over G thread groups of GLTMA_T threads (RB / G rows per thread): with G = 2
// a thread carries 6 instead of 12 lazy accumulators (about 110 registers), doubling the resident warps.
template <int RB, int G>
__global__ void __launch_bounds__(GLTMA_T * G, 2)
gl_matvec_tma_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                     const u64* __restrict__ v, u64* __restrict__ partial) {
    typedef GLSlot S;
    constexpr int CS = GLTMA_T;                  // slots per chunk
    constexpr int RBT = RB / G;                  // rows per thread
    constexpr uint32_t ROWB = CS * 24;           // bytes per row per stage
    static_assert(RB % G == 0, "row split");
    extern __shared__ __align__(128) unsigned char smem[];
    u64* stage = reinterpret_cast<u64*>(smem);   // [NS][RB + 1][CS * 3]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)GLTMA_NS * (RB + 1) * ROWB);
    __shared__ S::Val red[GLTMA_T];

    const int slot = threadIdx.x % CS, grp = threadIdx.x / CS;
    const size_t total = ncols * S::SLOTS;
    const size_t nchunks = (total + CS - 1) / CS;
    const size_t my_chunks = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GLTMA_NS; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](size_t it) {  // thread 0 only
        const int s = (int)(it % GLTMA_NS);
        const size_t chunk = blockIdx.x + it * gridDim.x;
        const size_t slot0 = chunk * CS;
        const uint32_t bytes = (uint32_t)(((total - slot0 < (size_t)CS) ? (total - slot0) : (size_t)CS) * 24);
        mbar_arrive_expect_tx(&full[s], bytes * (RB + 1));
        u64* dst = stage + (size_t)s * (RB + 1) * CS * 3;
        tma_load_1d(dst, v + slot0 * 3, bytes, &full[s]);
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const u64* rp = (row0 + r < nrows) ? rows[row0 + r] : rows[nrows - 1];
            tma_load_1d(dst + (size_t)(r + 1) * CS * 3, rp + slot0 * 3, bytes, &full[s]);
        }
    };
    if (threadIdx.x == 0)
        for (size_t it = 0; it < (size_t)GLTMA_NS && it < my_chunks; it++) issue(it);

    GLAcc acc[RBT][3];
#pragma unroll
    for (int r = 0; r < RBT; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) gl_acc_zero(acc[r][k]);

    for (size_t it = 0; it < my_chunks; it++) {
        const int s = (int)(it % GLTMA_NS);
        mbar_wait(&full[s], (uint32_t)((it / GLTMA_NS) & 1));
        const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
        const bool live = slot0 + slot < total;
        const u64* base = stage + (size_t)s * (RB + 1) * CS * 3 + slot * 3;
        u64 x0 = 0, x1 = 0, x2 = 0, a[RBT][3];
        if (live) { x0 = base[0]; x1 = base[1]; x2 = base[2]; }
#pragma unroll
        for (int r = 0; r < RBT; r++) {
            const u64* q = base + (size_t)(grp * RBT + r + 1) * CS * 3;
            a[r][0] = live ? q[0] : 0; a[r][1] = live ? q[1] : 0; a[r][2] = live ? q[2] : 0;
        }
        __syncthreads();  // every thread has read stage s: it can be refilled
        if (threadIdx.x == 0 && it + GLTMA_NS < my_chunks) issue(it + GLTMA_NS);
        const u64 xr1 = gl::mul_pow2<gl::root_exp(1)>(x1), xr2 = gl::mul_pow2<gl::root_exp(1)>(x2);
#pragma unroll
        for (int r = 0; r < RBT; r++) {
            gl_acc_mad(acc[r][0], a[r][0], x0);
            gl_acc_mad(acc[r][0], a[r][1], xr2);
            gl_acc_mad(acc[r][0], a[r][2], xr1);
            gl_acc_mad(acc[r][1], a[r][0], x1);
            gl_acc_mad(acc[r][1], a[r][1], x0);
            gl_acc_mad(acc[r][1], a[r][2], xr2);
            gl_acc_mad(acc[r][2], a[r][0], x2);
            gl_acc_mad(acc[r][2], a[r][1], x1);
            gl_acc_mad(acc[r][2], a[r][2], x0);
        }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        if (grp == r / RBT) {
            S::Val mine;
#pragma unroll
            for (int k = 0; k < 3; k++) mine.c[k] = gl_acc_reduce<128>(acc[r % RBT][k]);
            red[slot] = mine;
        }
        __syncthreads();
        if (threadIdx.x < S::SLOTS) {
            S::Val sacc = red[threadIdx.x];
            for (int k = threadIdx.x + S::SLOTS; k < GLTMA_T; k += S::SLOTS) S::acc(sacc, red[k]);
            S::store(partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + threadIdx.x * S::SLOT_U64, sacc);
        }
        __syncthreads();
    }
}

// Warp-specialised flavour: one producer warp (lane 0 issues the bulk copies) plus GLTMA_T consumer threads.
// Stages are recycled through per-stage "empty" mbarriers (one arrival per consumer warp), so there is no
// CTA barrier in the column loop and the consumer warps run up to NS stages apart.
template <int RB>
__global__ void __launch_bounds__(GLTMA_T + 32, 2)
gl_matvec_ws_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                    const u64* __restrict__ v, u64* __restrict__ partial) {
    typedef GLSlot S;
    constexpr int CS = GLTMA_T, NS = GLTMA_NS;
    constexpr uint32_t ROWB = CS * 24;
    extern __shared__ __align__(128) unsigned char smem[];
    u64* stage = reinterpret_cast<u64*>(smem);   // [NS][RB + 1][CS * 3]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * (RB + 1) * ROWB);
    uint64_t* empty = full + NS;
    __shared__ S::Val red[GLTMA_T];

    const bool producer = threadIdx.x >= CS;
    const int slot = threadIdx.x;  // consumers only
    const size_t total = ncols * S::SLOTS;
    const size_t nchunks = (total + CS - 1) / CS;
    const size_t my_chunks = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    GLAcc acc[RB][3];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) gl_acc_zero(acc[r][k]);

    if (producer) {
        if (threadIdx.x == CS) {
            for (size_t it = 0; it < my_chunks; it++) {
                const int s = (int)(it % NS);
                if (it >= (size_t)NS) mbar_wait(&empty[s], (uint32_t)(((it / NS) - 1) & 1));
                const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
                const uint32_t bytes = (uint32_t)(((total - slot0 < (size_t)CS) ? (total - slot0) : (size_t)CS) * 24);
                mbar_arrive_expect_tx(&full[s], bytes * (RB + 1));
                u64* dst = stage + (size_t)s * (RB + 1) * CS * 3;
                tma_load_1d(dst, v + slot0 * 3, bytes, &full[s]);
#pragma unroll
                for (int r = 0; r < RB; r++) {
                    const u64* rp = (row0 + r < nrows) ? rows[row0 + r] : rows[nrows - 1];
                    tma_load_1d(dst + (size_t)(r + 1) * CS * 3, rp + slot0 * 3, bytes, &full[s]);
                }
            }
        }
    } else {
        for (size_t it = 0; it < my_chunks; it++) {
            const int s = (int)(it % NS);
            mbar_wait(&full[s], (uint32_t)((it / NS) & 1));
            const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
            const bool live = slot0 + slot < total;
            const u64* base = stage + (size_t)s * (RB + 1) * CS * 3 + slot * 3;
            u64 x0 = 0, x1 = 0, x2 = 0, a[RB][3];
            if (live) { x0 = base[0]; x1 = base[1]; x2 = base[2]; }
#pragma unroll
            for (int r = 0; r < RB; r++) {
                const u64* q = base + (size_t)(r + 1) * CS * 3;
                a[r][0] = live ? q[0] : 0; a[r][1] = live ? q[1] : 0; a[r][2] = live ? q[2] : 0;
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
            const u64 xr1 = gl::mul_pow2<gl::root_exp(1)>(x1), xr2 = gl::mul_pow2<gl::root_exp(1)>(x2);
#pragma unroll
            for (int r = 0; r < RB; r++) {
                gl_acc_mad(acc[r][0], a[r][0], x0);
                gl_acc_mad(acc[r][0], a[r][1], xr2);
                gl_acc_mad(acc[r][0], a[r][2], xr1);
                gl_acc_mad(acc[r][1], a[r][0], x1);
                gl_acc_mad(acc[r][1], a[r][1], x0);
                gl_acc_mad(acc[r][1], a[r][2], xr2);
                gl_acc_mad(acc[r][2], a[r][0], x2);
                gl_acc_mad(acc[r][2], a[r][1], x1);
                gl_acc_mad(acc[r][2], a[r][2], x0);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        if (!producer) {
            S::Val mine;
#pragma unroll
            for (int k = 0; k < 3; k++) mine.c[k] = gl_acc_reduce<128>(acc[r][k]);
            red[slot] = mine;
        }
        __syncthreads();
        if (threadIdx.x < S::SLOTS) {
            S::Val sacc = red[threadIdx.x];
            for (int k = threadIdx.x + S::SLOTS; k < GLTMA_T; k += S::SLOTS) S::acc(sacc, red[k]);
            S::store(partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + threadIdx.x * S::SLOT_U64, sacc);
        }
        __syncthreads();
    }
}

// The same with the rows split over G consumer groups (as gl_matvec_tma_kernel): no CTA barrier in the column loop.
template <int RB, int G>
__global__ void __launch_bounds__(GLTMA_T * G + 32, 2)
gl_matvec_wsg_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                    const u64* __restrict__ v, u64* __restrict__ partial) {
    typedef GLSlot S;
    constexpr int CS = GLTMA_T, NS = GLTMA_NS, RBT = RB / G;
    static_assert(RB % G == 0, "row split");
    constexpr uint32_t ROWB = CS * 24;
    extern __shared__ __align__(128) unsigned char smem[];
    u64* stage = reinterpret_cast<u64*>(smem);   // [NS][RB + 1][CS * 3]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * (RB + 1) * ROWB);
    uint64_t* empty = full + NS;
    __shared__ S::Val red[GLTMA_T];

    const bool producer = threadIdx.x >= CS * G;
    const int slot = threadIdx.x % CS, grp = threadIdx.x / CS;  // consumers only
    const size_t total = ncols * S::SLOTS;
    const size_t nchunks = (total + CS - 1) / CS;
    const size_t my_chunks = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], G * CS / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    GLAcc acc[RBT][3];
#pragma unroll
    for (int r = 0; r < RBT; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) gl_acc_zero(acc[r][k]);

    if (producer) {
        if (threadIdx.x == CS * G) {
            for (size_t it = 0; it < my_chunks; it++) {
                const int s = (int)(it % NS);
                if (it >= (size_t)NS) mbar_wait(&empty[s], (uint32_t)(((it / NS) - 1) & 1));
                const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
                const uint32_t bytes = (uint32_t)(((total - slot0 < (size_t)CS) ? (total - slot0) : (size_t)CS) * 24);
                mbar_arrive_expect_tx(&full[s], bytes * (RB + 1));
                u64* dst = stage + (size_t)s * (RB + 1) * CS * 3;
                tma_load_1d(dst, v + slot0 * 3, bytes, &full[s]);
#pragma unroll
                for (int r = 0; r < RB; r++) {
                    const u64* rp = (row0 + r < nrows) ? rows[row0 + r] : rows[nrows - 1];
                    tma_load_1d(dst + (size_t)(r + 1) * CS * 3, rp + slot0 * 3, bytes, &full[s]);
                }
            }
        }
    } else {
        for (size_t it = 0; it < my_chunks; it++) {
            const int s = (int)(it % NS);
            mbar_wait(&full[s], (uint32_t)((it / NS) & 1));
            const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
            const bool live = slot0 + slot < total;
            const u64* base = stage + (size_t)s * (RB + 1) * CS * 3 + slot * 3;
            u64 x0 = 0, x1 = 0, x2 = 0, a[RBT][3];
            if (live) { x0 = base[0]; x1 = base[1]; x2 = base[2]; }
#pragma unroll
            for (int r = 0; r < RBT; r++) {
                const u64* q = base + (size_t)(grp * RBT + r + 1) * CS * 3;
                a[r][0] = live ? q[0] : 0; a[r][1] = live ? q[1] : 0; a[r][2] = live ? q[2] : 0;
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
            const u64 xr1 = gl::mul_pow2<gl::root_exp(1)>(x1), xr2 = gl::mul_pow2<gl::root_exp(1)>(x2);
#pragma unroll
            for (int r = 0; r < RBT; r++) {
                gl_acc_mad(acc[r][0], a[r][0], x0);
                gl_acc_mad(acc[r][0], a[r][1], xr2);
                gl_acc_mad(acc[r][0], a[r][2], xr1);
                gl_acc_mad(acc[r][1], a[r][0], x1);
                gl_acc_mad(acc[r][1], a[r][1], x0);
                gl_acc_mad(acc[r][1], a[r][2], xr2);
                gl_acc_mad(acc[r][2], a[r][0], x2);
                gl_acc_mad(acc[r][2], a[r][1], x1);
                gl_acc_mad(acc[r][2], a[r][2], x0);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        if (!producer && grp == r / RBT) {
            S::Val mine;
#pragma unroll
            for (int k = 0; k < 3; k++) mine.c[k] = gl_acc_reduce<128>(acc[r % RBT][k]);
            red[slot] = mine;
        }
        __syncthreads();
        if (threadIdx.x < S::SLOTS) {
            S::Val sacc = red[threadIdx.x];
            for (int k = threadIdx.x + S::SLOTS; k < GLTMA_T; k += S::SLOTS) S::acc(sacc, red[k]);
            S::store(partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + threadIdx.x * S::SLOT_U64, sacc);
        }
        __syncthreads();
    }
}

#ifndef SR_GLMV_G
#define SR_GLMV_G 2
#endif
// 4-row passes: warp-specialised row-split kernel (no CTA barrier in the column loop) unless -DSR_GLMV_CTASYNC:
// 0.582 vs 0.567 of the HBM roofline at m = 2^20, 0.652 vs 0.628 at 2^22 (kappa = 4)
#if !defined(SR_GLMV_CTASYNC) && !defined(SR_GLMV_WSG)
#define SR_GLMV_WSG
#endif
template <int RB>
static cudaError_t gl_tma_launch_rb(int grid, const u64* const* d_rows, size_t nrows, size_t row0, size_t ncols,
                                    const u64* v, u64* parts, cudaStream_t st) {
    // RB < 4: warp-specialised producer/consumer kernel (kappa = 1: 0.100 vs 0.157 ms at m = 2^20);
    // RB = 4: the row-split kernel (the 12-accumulator consumer would spill under the 168-register cap).
    constexpr bool WS = (RB < 4);
    constexpr int G = WS ? 1 : SR_GLMV_G;
#if defined(SR_GLMV_WSG)
    void (*kern)(const u64* const*, size_t, size_t, size_t, const u64*, u64*) =
        WS ? (void (*)(const u64* const*, size_t, size_t, size_t, const u64*, u64*))gl_matvec_ws_kernel<(RB < 4 ? RB : 1)>
           : (void (*)(const u64* const*, size_t, size_t, size_t, const u64*, u64*))gl_matvec_wsg_kernel<RB, (RB >= 4 ? SR_GLMV_G : 1)>;
    const int threads = WS ? GLTMA_T + 32 : GLTMA_T * G + 32;
#else
    void (*kern)(const u64* const*, size_t, size_t, size_t, const u64*, u64*) =
        WS ? (void (*)(const u64* const*, size_t, size_t, size_t, const u64*, u64*))gl_matvec_ws_kernel<(RB < 4 ? RB : 1)>
           : (void (*)(const u64* const*, size_t, size_t, size_t, const u64*, u64*))gl_matvec_tma_kernel<RB, (RB >= 4 ? SR_GLMV_G : 1)>;
    const int threads = WS ? GLTMA_T + 32 : GLTMA_T * G;
#endif
    const size_t smem = (size_t)GLTMA_NS * (RB + 1) * GLTMA_T * 24 + 2 * GLTMA_NS * sizeof(uint64_t);
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, threads, smem, st>>>(d_rows, nrows, row0, ncols, v, parts);
    return cudaGetLastError();
}

// Three threads per slot: the nine products of a slot product fall on five diagonals
//   d0 = a0 x0, d1 = a0 x1 + a1 x0, d2 = a0 x2 + a1 x1 + a2 x0, d3 = a1 x2 + a2 x1, d4 = a2 x2,
//   c0 = d0 + r d3, c1 = d1 + r d4, c2 = d2   (u^3 = r = 2^40),
// and are split 3 + 3 + 3 over the threads (part 0: d0 | d3, part 1: d4 | d1, part 2: a0 x2 | a1 x1 + a2 x0),
// each thread keeping two lazy accumulators per matrix row.  No multiplication by r and no reduction
// inside the column loop; operand indices are per-lane constants, so all lanes run one instruction
// stream.  ~100 registers -> 18 warps per SM instead of 8.
#ifndef SR_GL3_MINB
#define SR_GL3_MINB 2
#endif
constexpr int GL3_SLOTS = 64, GL3_T = 3 * GL3_SLOTS, GL3_NS = 4;
template <int RB>
__global__ void __launch_bounds__(GL3_T, SR_GL3_MINB)
gl_matvec_tma3_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                      const u64* __restrict__ v, u64* __restrict__ partial) {
    typedef GLSlot S;
    constexpr int CS = GL3_SLOTS;
    constexpr uint32_t ROWB = CS * 24;
    extern __shared__ __align__(128) unsigned char smem[];
    u64* stage = reinterpret_cast<u64*>(smem);   // [NS][RB + 1][CS * 3]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)GL3_NS * (RB + 1) * ROWB);
    __shared__ u64 red[GL3_SLOTS][3];

    const int slot = threadIdx.x / 3, part = threadIdx.x - 3 * slot;
    // operand indices: accA += a[ia] x[ja];  accB += a[ib1] x[jb1] + a[ib2] x[jb2]
    const int ia = (part == 1) ? 2 : 0, ja = (part == 0) ? 0 : 2;
    const int ib1 = (part == 1) ? 0 : 1, jb1 = (part == 0) ? 2 : 1;
    const int ib2 = (part == 1) ? 1 : 2, jb2 = (part == 0) ? 1 : 0;

    const u64* rp[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) rp[r] = (row0 + r < nrows) ? rows[row0 + r] : rows[nrows - 1];
    const size_t total = ncols * S::SLOTS;
    const size_t nchunks = (total + CS - 1) / CS;
    const size_t my_chunks = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GL3_NS; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](size_t it) {
        const int s = (int)(it % GL3_NS);
        const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
        const uint32_t bytes = (uint32_t)(((total - slot0 < (size_t)CS) ? (total - slot0) : (size_t)CS) * 24);
        mbar_arrive_expect_tx(&full[s], bytes * (RB + 1));
        u64* dst = stage + (size_t)s * (RB + 1) * CS * 3;
        tma_load_1d(dst, v + slot0 * 3, bytes, &full[s]);
#pragma unroll
        for (int r = 0; r < RB; r++) tma_load_1d(dst + (size_t)(r + 1) * CS * 3, rp[r] + slot0 * 3, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (size_t it = 0; it < (size_t)GL3_NS && it < my_chunks; it++) issue(it);

    GLAcc accA[RB], accB[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) { gl_acc_zero(accA[r]); gl_acc_zero(accB[r]); }

    for (size_t it = 0; it < my_chunks; it++) {
        const int s = (int)(it % GL3_NS);
        mbar_wait(&full[s], (uint32_t)((it / GL3_NS) & 1));
        const size_t slot0 = (blockIdx.x + it * gridDim.x) * CS;
        const bool live = slot0 + slot < total;
        const u64* base = stage + (size_t)s * (RB + 1) * CS * 3 + slot * 3;
        u64 xa = 0, xb1 = 0, xb2 = 0, aa[RB], ab1[RB], ab2[RB];
        if (live) { xa = base[ja]; xb1 = base[jb1]; xb2 = base[jb2]; }
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const u64* q = base + (size_t)(r + 1) * CS * 3;
            aa[r] = live ? q[ia] : 0; ab1[r] = live ? q[ib1] : 0; ab2[r] = live ? q[ib2] : 0;
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + GL3_NS < my_chunks) issue(it + GL3_NS);
#pragma unroll
        for (int r = 0; r < RB; r++) {
            gl_acc_mad(accA[r], aa[r], xa);
            gl_acc_mad(accB[r], ab1[r], xb1);
            gl_acc_mad(accB[r], ab2[r], xb2);
        }
    }
    // thread -> one coefficient of the slot: part 0: c0 = A + r B, part 1: c1 = B + r A, part 2: c2 = A + B
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        const u64 ra = gl_acc_reduce<0>(accA[r]), rb = gl_acc_reduce<0>(accB[r]);
        u64 c;
        if (part == 0) c = gl::add(gl::mul_pow2<gl::root_exp(1)>(rb), ra);
        else if (part == 1) c = gl::add(gl::mul_pow2<gl::root_exp(1)>(ra), rb);
        else c = gl::add(ra, rb);
        red[slot][part] = gl::canon(gl::mul_pow2<128>(c));  // Montgomery layout: extra 2^-64 = 2^128
        __syncthreads();
        if (threadIdx.x < S::SLOTS * 3) {  // 24 threads: (slot index s8, coefficient k)
            const int s8 = threadIdx.x / 3, k = threadIdx.x - 3 * s8;
            u64 acc = 0;
            for (int q = s8; q < GL3_SLOTS; q += S::SLOTS) acc = gl::add(acc, red[q][k]);
            partial[((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + s8 * 3 + k] = gl::canon(acc);
        }
        __syncthreads();
    }
}

template <int RB>
static cudaError_t gl_tma3_launch_rb(int grid, const u64* const* d_rows, size_t nrows, size_t row0, size_t ncols,
                                     const u64* v, u64* parts, cudaStream_t st) {
    auto kern = gl_matvec_tma3_kernel<RB>;
    const size_t smem = (size_t)GL3_NS * (RB + 1) * GL3_SLOTS * 24 + GL3_NS * sizeof(uint64_t);
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, GL3_T, smem, st>>>(d_rows, nrows, row0, ncols, v, parts);
    return cudaGetLastError();
}

#ifndef SR_MV_T
#define SR_MV_T 256
#endif
#ifndef SR_MV_RB
#define SR_MV_RB 4
#endif
constexpr int MV_T = SR_MV_T;  // threads per CTA (multiple of every SLOTS)

// partial[(blockIdx * nrows + row) * ELEM + slot*SLOT_U64 ...] = CTA-local sum for rows [row0, row0+RB)
template <class S, int RB>
__global__ void __launch_bounds__(MV_T)
matvec_partial_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                      const u64* __restrict__ v, u64* __restrict__ partial) {
    __shared__ typename S::Val red[MV_T];
    typename S::Val acc[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) acc[r] = S::zero();
    const u64* rp[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) rp[r] = (row0 + r < nrows) ? rows[row0 + r] : nullptr;

    const size_t total = ncols * S::SLOTS;  // slots per row
    const size_t stride = (size_t)gridDim.x * MV_T;
    for (size_t g = (size_t)blockIdx.x * MV_T + threadIdx.x; g < total; g += stride) {
        const typename S::Val x = S::load_cached(v + g * S::SLOT_U64);
        typename S::Val a[RB];
#pragma unroll
        for (int r = 0; r < RB; r++)
            if (rp[r]) a[r] = S::load(rp[r] + g * S::SLOT_U64);
#pragma unroll
        for (int r = 0; r < RB; r++)
            if (rp[r]) S::acc(acc[r], S::mul_lazy(a[r], x));
    }
    // CTA reduction per slot index: thread t holds slot t % SLOTS
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        S::finish(acc[r]);
        red[threadIdx.x] = acc[r];
        __syncthreads();
        if (threadIdx.x < S::SLOTS) {
            typename S::Val s = red[threadIdx.x];
            for (int k = threadIdx.x + S::SLOTS; k < MV_T; k += S::SLOTS) S::acc(s, red[k]);
            S::store(partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + threadIdx.x * S::SLOT_U64, s);
        }
        __syncthreads();
    }
}

// ---- NVLink peer-memory hand-off of the column-sharded commitment -----------------------------------
// The last kernel of a rank's partial product writes its nrows partial elements STRAIGHT into the root rank's
// mailbox (peer memory mapped through CUDA IPC; plain stores travel over NVLink), then publishes an epoch flag
// with release semantics at system scope.  The root's reduction kernel acquires the flags of all ranks and adds the
// partials mod p.  No NCCL call, no host synchronisation, no extra launch: the exchange is fused into the two
// kernels that produce and consume the data.  Slots are reused every MAILBOX_DEPTH epochs; a writer first checks
// the root's `consumed` counter (peer load) so that it never overruns a slot the root has not summed yet.
SR_D u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SR_D void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SR_D unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;  // a lost peer must not hang the GPU: flag an error
// spin until *p >= want; false (and *err = 1) on timeout
SR_D bool spin_until(const u64* p, u64 want, int* err) {
    if (ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(p) < want) {
        __nanosleep(200);
        if (global_ns() - t0 > PEER_TIMEOUT_NS) {
            atomicExch(err, 1);
            return false;
        }
    }
    return true;
}
// out[row] = sum_k parts[(k * stride_rows + row)]  (k < nparts).  One CTA per (row, slot), fixed-order tree: deterministic.
// (A warp per (row, slot) took 10.6 us for the 296 partials of a kappa = 4 commit, a quarter of the per-rank time at
// 8 GPUs: 32 warps on the whole GPU, each chaining ten dependent loads.)
// ps.role 1 (writer): `out` is replaced by this rank's mailbox slot of the epoch; the kernel first makes sure the
// root has summed the epoch that used the slot before, and publishes the epoch flag once every block has stored.
// ps.role 2 (root): `parts` is replaced by the epoch's nranks slots; the kernel first acquires all rank flags and
// publishes `consumed` at the end.
template <class S>
__global__ void __launch_bounds__(128)
sum_partials_kernel(const u64* __restrict__ parts, size_t nparts, size_t stride_rows, size_t nrows,
                    u64* __restrict__ out, PeerSync ps) {
    __shared__ typename S::Val red[128];
    u64 epoch = 0;
    if (ps.role != 0) {  // uniform per launch
        epoch = ps.epoch ? ps.epoch : *reinterpret_cast<volatile u64*>(ps.epoch_ctr) + 1;
        const size_t slot0 = (size_t)(epoch % MAILBOX_DEPTH) * ps.nranks;
        if (ps.role == 1) {
            out = ps.slots + (slot0 + ps.rank) * ps.slot_stride;
            if (threadIdx.x == 0 && epoch > (u64)MAILBOX_DEPTH) spin_until(ps.consumed, epoch - MAILBOX_DEPTH, ps.err);
        } else {
            parts = ps.slots + slot0 * ps.slot_stride;
            for (int r = threadIdx.x; r < ps.nranks; r += 128) spin_until(ps.flags + r, epoch, ps.err);  // a lane per rank
        }
        __syncthreads();
    }
    // one CTA per (row, slot): the 128 threads stride over the partials (independent loads, two or three each for the
    // 296 partials of a mat-vec), then a shared-memory tree adds the 128 thread sums in a fixed order
    const size_t idx = blockIdx.x;  // (row, slot)
    const size_t row = idx / S::SLOTS, slot = idx % S::SLOTS;
    typename S::Val s = S::zero();
    for (size_t k = threadIdx.x; k < nparts; k += 128)
        S::acc(s, S::load(parts + (k * stride_rows + row) * S::ELEM_U64 + slot * S::SLOT_U64));
    red[threadIdx.x] = s;
    __syncthreads();
#pragma unroll
    for (int w = 64; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) {
            typename S::Val t = red[threadIdx.x];
            S::acc(t, red[threadIdx.x + w]);
            red[threadIdx.x] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) S::store(out + row * S::ELEM_U64 + slot * S::SLOT_U64, red[0]);
    if (ps.role != 0) {
        // every block has read the epoch counter before the last one arrives, so it may be advanced here
        u64* flag = ps.role == 1 ? ps.flags + ps.rank : ps.consumed;
        __threadfence_system();  // this thread's stores (possibly to peer memory) before the arrival
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned arrived = atomicAdd(ps.counter, 1u);
            if (arrived == gridDim.x - 1) {
                *ps.counter = 0;  // ready for the next launch on this stream
                if (ps.role == 1) *ps.epoch_ctr = epoch;
                __threadfence_system();
                st_release_sys(flag, epoch);
            }
        }
    }
}

static int mv_grid(int sms) { return sms * 4; }

template <class S>
static cudaError_t final_sum(const u64* parts, size_t nparts, size_t nrows, u64* out, const PeerSync& ps,
                             cudaStream_t st) {
    const size_t n = nrows * S::SLOTS;
    sum_partials_kernel<S><<<(unsigned)n, 128, 0, st>>>(parts, nparts, nrows, nrows, out, ps);
    return cudaGetLastError();
}

static cudaError_t gl_matvec_launch(const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v, u64* out,
                                    void* scratch, cudaStream_t st, int sms, int* launches, const PeerSync& ps) {
    typedef GLSlot S;
    *launches = 0;
    if (nrows == 0) return cudaSuccess;
    if (ncols == 0) {  // Sum of nothing = ZERO (a sum over zero partials, so that a mailbox flag is still published)
        (*launches)++;
        return final_sum<S>(reinterpret_cast<u64*>(scratch), 0, nrows, out, ps, st);
    }
    size_t total = ncols * S::SLOTS;
    int grid = mv_grid(sms);  // scratch is sized for mv_grid(sms) partials
    size_t need = (total + GLMV_T - 1) / GLMV_T;
    if ((size_t)grid > need) grid = (int)need;
    u64* parts = reinterpret_cast<u64*>(scratch);
    size_t row0 = 0;
#if defined(SR_GLMV_TMA3)
    {
        int g2 = sms * SR_GL3_MINB;
        const size_t nchunks = (total + GL3_SLOTS - 1) / GL3_SLOTS;
        if ((size_t)g2 > nchunks) g2 = (int)nchunks;
        while (row0 < nrows) {
            const size_t left = nrows - row0;
            cudaError_t e;
            if (left >= 4) { e = gl_tma3_launch_rb<4>(g2, d_rows, nrows, row0, ncols, v, parts, st); row0 += 4; }
            else if (left >= 2) { e = gl_tma3_launch_rb<2>(g2, d_rows, nrows, row0, ncols, v, parts, st); row0 += 2; }
            else { e = gl_tma3_launch_rb<1>(g2, d_rows, nrows, row0, ncols, v, parts, st); row0 += 1; }
            if (e != cudaSuccess) return e;
            (*launches)++;
        }
        grid = g2;
    }
#elif !defined(SR_GLMV_NO_TMA)
    {
        int g2 = sms * 2;  // two resident CTAs per SM
        const size_t nchunks = (total + GLTMA_T - 1) / GLTMA_T;
        if ((size_t)g2 > nchunks) g2 = (int)nchunks;
        while (row0 < nrows) {
            const size_t left = nrows - row0;
            cudaError_t e;
            if (left >= 4) { e = gl_tma_launch_rb<4>(g2, d_rows, nrows, row0, ncols, v, parts, st); row0 += 4; }
            else if (left >= 2) { e = gl_tma_launch_rb<2>(g2, d_rows, nrows, row0, ncols, v, parts, st); row0 += 2; }
            else { e = gl_tma_launch_rb<1>(g2, d_rows, nrows, row0, ncols, v, parts, st); row0 += 1; }
            if (e != cudaSuccess) return e;
            (*launches)++;
        }
        grid = g2;
    }
#endif
    while (row0 < nrows) {
        const size_t left = nrows - row0;
        if (left >= 4 && SR_GLMV_RB >= 4) { gl_matvec_lazy_kernel<4><<<grid, GLMV_T, 0, st>>>(d_rows, nrows, row0, ncols, v, parts); row0 += 4; }
        else if (left >= 2) { gl_matvec_lazy_kernel<2><<<grid, GLMV_T, 0, st>>>(d_rows, nrows, row0, ncols, v, parts); row0 += 2; }
        else { gl_matvec_lazy_kernel<1><<<grid, GLMV_T, 0, st>>>(d_rows, nrows, row0, ncols, v, parts); row0 += 1; }
        (*launches)++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    (*launches)++;
    return final_sum<S>(parts, (size_t)grid, nrows, out, ps, st);
}

size_t matvec_scratch_bytes(int ring, size_t nrows, int sms) {
    const size_t w = ring == RING_GL ? 24 : ring == RING_BB ? 72 : 64;
    return (size_t)mv_grid(sms) * nrows * w * 8 + 16;
}

template <class S>
static cudaError_t matvec_launch_t(const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v, u64* out,
                                   void* scratch, cudaStream_t st, int sms, int* launches, const PeerSync& ps) {
    *launches = 0;
    if (nrows == 0) return cudaSuccess;
    if (ncols == 0) {  // Sum of nothing = ZERO
        (*launches)++;
        return final_sum<S>(reinterpret_cast<u64*>(scratch), 0, nrows, out, ps, st);
    }
    size_t total = ncols * S::SLOTS;
    int grid = mv_grid(sms);
    size_t need = (total + MV_T - 1) / MV_T;
    if ((size_t)grid > need) grid = (int)need;
    u64* parts = reinterpret_cast<u64*>(scratch);
    constexpr int RB = SR_MV_RB;
    for (size_t row0 = 0; row0 < nrows; row0 += RB) {
        matvec_partial_kernel<S, RB><<<grid, MV_T, 0, st>>>(d_rows, nrows, row0, ncols, v, parts);
        (*launches)++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    (*launches)++;
    return final_sum<S>(parts, (size_t)grid, nrows, out, ps, st);
}

// ps (optional): `out` is a slot of the root's mailbox; see PeerSync
cudaError_t matvec_launch(int ring, const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v, u64* out,
                          void* scratch, cudaStream_t st, int sms, int* launches, const PeerSync* ps) {
    const PeerSync none = {};
    const PeerSync& p = ps ? *ps : none;
    switch (ring) {
    case RING_GL: return gl_matvec_launch(d_rows, nrows, ncols, v, out, scratch, st, sms, launches, p);
    case RING_BB: return matvec_launch_t<BBSlot>(d_rows, nrows, ncols, v, out, scratch, st, sms, launches, p);
    case RING_SP: return matvec_launch_t<SPSlot>(d_rows, nrows, ncols, v, out, scratch, st, sms, launches, p);
    }
    return cudaErrorInvalidValue;
}

// out[i] = sum_r gathered[(r * stride_rows + i)] over nranks partials; ps (optional): the root's mailbox reduction
template <class S>
static cudaError_t modsum_t(const u64* g, size_t nranks, size_t stride_rows, size_t nrows, u64* out,
                            const PeerSync& ps, cudaStream_t st) {
    const size_t n = nrows * S::SLOTS;
    sum_partials_kernel<S><<<(unsigned)n, 128, 0, st>>>(g, nranks, stride_rows, nrows, out, ps);
    return cudaGetLastError();
}
cudaError_t modsum_launch(int ring, const u64* gathered, size_t nranks, size_t stride_rows, size_t nrows, u64* out,
                          cudaStream_t st, const PeerSync* ps) {
    const PeerSync none = {};
    const PeerSync& p = ps ? *ps : none;
    switch (ring) {
    case RING_GL: return modsum_t<GLSlot>(gathered, nranks, stride_rows, nrows, out, p, st);
    case RING_BB: return modsum_t<BBSlot>(gathered, nranks, stride_rows, nrows, out, p, st);
    case RING_SP: return modsum_t<SPSlot>(gathered, nranks, stride_rows, nrows, out, p, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
