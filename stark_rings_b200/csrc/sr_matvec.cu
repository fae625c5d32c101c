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

#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "sp_ring.cuh"

namespace sr {

// ---- per-ring slot traits ------------------------------------------------------------------
struct GLSlot {
    static constexpr int SLOTS = 8, SLOT_U64 = 3, ELEM_U64 = 24;
    struct Val { u64 c[3]; };
    SR_D static Val load(const u64* p) { Val v; v.c[0] = __ldcs(p); v.c[1] = __ldcs(p + 1); v.c[2] = __ldcs(p + 2); return v; }
    SR_D static Val load_cached(const u64* p) { Val v; v.c[0] = p[0]; v.c[1] = p[1]; v.c[2] = p[2]; return v; }
    SR_D static void store(u64* p, const Val& v) { p[0] = v.c[0]; p[1] = v.c[1]; p[2] = v.c[2]; }
    SR_D static Val zero() { Val v; v.c[0] = v.c[1] = v.c[2] = 0; return v; }
    SR_D static Val mul(const Val& a, const Val& b) { Val z; gl::slot_mul<gl::root_exp(1), 128>(z.c, a.c, b.c); return z; }
    SR_D static void acc(Val& s, const Val& x) {
#pragma unroll
        for (int i = 0; i < 3; i++) s.c[i] = gl::add(s.c[i], x.c[i]);
    }
};
struct BBSlot {
    static constexpr int SLOTS = 8, SLOT_U64 = 9, ELEM_U64 = 72;
    struct Val { u32 c[9]; };
    SR_D static Val load(const u64* p) {
        Val v;
        const u32* q = reinterpret_cast<const u32*>(p);
#pragma unroll
        for (int i = 0; i < 9; i++) v.c[i] = __ldcs(q + 2 * i);
        return v;
    }
    SR_D static Val load_cached(const u64* p) {
        Val v;
        const u32* q = reinterpret_cast<const u32*>(p);
#pragma unroll
        for (int i = 0; i < 9; i++) v.c[i] = q[2 * i];
        return v;
    }
    SR_D static void store(u64* p, const Val& v) {
#pragma unroll
        for (int i = 0; i < 9; i++) p[i] = v.c[i];
    }
    SR_D static Val zero() {
        Val v;
#pragma unroll
        for (int i = 0; i < 9; i++) v.c[i] = 0;
        return v;
    }
    SR_D static Val mul(const Val& a, const Val& b) { Val z; bb::slot_mul_ntt(z.c, a.c, b.c); return z; }
    SR_D static void acc(Val& s, const Val& x) {
#pragma unroll
        for (int i = 0; i < 9; i++) s.c[i] = bb::add(s.c[i], x.c[i]);
    }
};
struct SPSlot {
    static constexpr int SLOTS = 16, SLOT_U64 = 4, ELEM_U64 = 64;
    typedef sp::Fe Val;
    SR_D static Val load(const u64* p) {
        Val v;
        uint4 lo = __ldcs(reinterpret_cast<const uint4*>(p)), hi = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    SR_D static Val load_cached(const u64* p) {
        Val v;
        uint4 lo = reinterpret_cast<const uint4*>(p)[0], hi = reinterpret_cast<const uint4*>(p)[1];
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    SR_D static void store(u64* p, const Val& v) {
        reinterpret_cast<uint4*>(p)[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
        reinterpret_cast<uint4*>(p)[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
    SR_D static Val zero() {
        Val v;
#pragma unroll
        for (int i = 0; i < 8; i++) v.v[i] = 0;
        return v;
    }
    SR_D static Val mul(const Val& a, const Val& b) { Val z; sp::mont_mul(z, a, b); return z; }
    SR_D static void acc(Val& s, const Val& x) { Val t; sp::add(t, s, x); s = t; }
};

constexpr int MV_T = 256;  // threads per CTA (multiple of every SLOTS)

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
            if (rp[r]) S::acc(acc[r], S::mul(a[r], x));
    }
    // CTA reduction per slot index: thread t holds slot t % SLOTS
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
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

// out[row] = sum_k parts[k * nrows + row]  (k < nparts), one thread per (row, slot)
template <class S>
__global__ void sum_partials_kernel(const u64* __restrict__ parts, size_t nparts, size_t nrows, u64* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * S::SLOTS) return;
    const size_t row = idx / S::SLOTS, slot = idx % S::SLOTS;
    typename S::Val s = S::zero();
    for (size_t k = 0; k < nparts; k++)
        S::acc(s, S::load_cached(parts + (k * nrows + row) * S::ELEM_U64 + slot * S::SLOT_U64));
    S::store(out + row * S::ELEM_U64 + slot * S::SLOT_U64, s);
}

static int mv_grid(int sms) { return sms * 4; }

size_t matvec_scratch_bytes(int ring, size_t nrows, int sms) {
    const size_t w = ring == RING_GL ? 24 : ring == RING_BB ? 72 : 64;
    return (size_t)mv_grid(sms) * nrows * w * 8 + 16;
}

template <class S>
static cudaError_t matvec_launch_t(const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v, u64* out,
                                   void* scratch, cudaStream_t st, int sms, int* launches) {
    *launches = 0;
    if (nrows == 0) return cudaSuccess;
    if (ncols == 0) return cudaMemsetAsync(out, 0, nrows * S::ELEM_U64 * 8, st);  // Sum of nothing = ZERO
    size_t total = ncols * S::SLOTS;
    int grid = mv_grid(sms);
    size_t need = (total + MV_T - 1) / MV_T;
    if ((size_t)grid > need) grid = (int)need;
    u64* parts = reinterpret_cast<u64*>(scratch);
    constexpr int RB = 4;
    for (size_t row0 = 0; row0 < nrows; row0 += RB) {
        matvec_partial_kernel<S, RB><<<grid, MV_T, 0, st>>>(d_rows, nrows, row0, ncols, v, parts);
        (*launches)++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t n = nrows * S::SLOTS;
    sum_partials_kernel<S><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(parts, (size_t)grid, nrows, out);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t matvec_launch(int ring, const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v, u64* out,
                          void* scratch, cudaStream_t st, int sms, int* launches) {
    switch (ring) {
    case RING_GL: return matvec_launch_t<GLSlot>(d_rows, nrows, ncols, v, out, scratch, st, sms, launches);
    case RING_BB: return matvec_launch_t<BBSlot>(d_rows, nrows, ncols, v, out, scratch, st, sms, launches);
    case RING_SP: return matvec_launch_t<SPSlot>(d_rows, nrows, ncols, v, out, scratch, st, sms, launches);
    }
    return cudaErrorInvalidValue;
}

template <class S>
static cudaError_t modsum_t(const u64* g, size_t nranks, size_t nrows, u64* out, cudaStream_t st) {
    const size_t n = nrows * S::SLOTS;
    sum_partials_kernel<S><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, nranks, nrows, out);
    return cudaGetLastError();
}
cudaError_t modsum_launch(int ring, const u64* gathered, size_t nranks, size_t nrows, u64* out, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return modsum_t<GLSlot>(gathered, nranks, nrows, out, st);
    case RING_BB: return modsum_t<BBSlot>(gathered, nranks, nrows, out, st);
    case RING_SP: return modsum_t<SPSlot>(gathered, nranks, nrows, out, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
