// The callers' other linear maps over NTT-form ring elements, next to the dense mat-vec:
//   sparse mat-vec   SparseMatrix<R>::checked_mul_vec / try_mul_vec   linear_algebra/src/sparse_matrix.rs:201-217
//                    out[i] = sum over (r, j) in coeffs[i] of r * v[j], Sum folded from ZERO (ntt_form.rs:640-654)
//   dense mat-mat    Matrix<R>::checked_mul_mat / try_mul_mat         linear_algebra/src/matrix.rs:148-166
//                    out[i][j] = sum_k a[i][k] * m[k][j]
//   scalar scaling   MulAssign<&R> for Matrix<R> / SparseMatrix<R>    matrix.rs:207-211, sparse_matrix.rs:298-302
//                    every entry *= r
// All three are independent per CRT slot, so the unit of work is one slot of one output element (sr_slots.cuh):
// consecutive threads own consecutive slots, i.e. a group of SLOTS threads reads one whole element (192 / 576 /
// 512 contiguous bytes).  Sums are exact modular sums, so any association gives the reference's bits.
//
// The sparse matrix crosses the boundary as the CSR image of coeffs: Vec<Vec<(R, usize)>> (sparse_matrix.rs:17-21):
// row_ptr[nrows + 1] entry offsets, col_idx[nnz], vals = nnz elements in row order.  Short rows (constraint
// matrices: a handful of entries per row) take the thread-per-(row, slot) kernel; long rows the warp-per-row one.
#include <cuda_runtime.h>

#define SR_GL_EPS_ON_ALU  // these kernels are bound by the multiply-add pipe: see gl_ring.cuh plus_eps_if

#include "sr_slots.cuh"

namespace sr {

constexpr int SPMV_T = 128;

// thread per (row, slot)
template <class S>
__global__ void __launch_bounds__(SPMV_T)
sparse_matvec_kernel(const u64* __restrict__ row_ptr, const u64* __restrict__ col_idx, const u64* __restrict__ vals,
                     const u64* __restrict__ v, size_t nrows, size_t ncols, size_t nnz, u64* __restrict__ out,
                     int* __restrict__ bad) {
    const size_t idx = (size_t)blockIdx.x * SPMV_T + threadIdx.x;
    if (idx >= nrows * S::SLOTS) return;
    const size_t row = idx / S::SLOTS;
    const int slot = (int)(idx - row * S::SLOTS);
    typename S::Accum acc;
    S::accum_zero(acc);
    u64 e0 = row_ptr[row], e1 = row_ptr[row + 1];
    if (e0 > e1 || e1 > nnz) {  // a malformed row_ptr must not turn into out-of-bounds reads
        *bad = 2;
        e1 = e0;
    }
    for (u64 e = e0; e < e1; e++) {
        const u64 c = col_idx[e];
        if (c >= ncols) {  // the reference indexes v[*i] out of bounds and panics
            *bad = 1;
            continue;
        }
        const typename S::Val a = S::load(vals + e * S::ELEM_U64 + slot * S::SLOT_U64);
        const typename S::Val x = S::load_cached(v + c * S::ELEM_U64 + slot * S::SLOT_U64);
        S::accum_mad(acc, a, x);
    }
    S::store(out + row * S::ELEM_U64 + slot * S::SLOT_U64, S::accum_result(acc));
}

// warp per row: lane = (sub, slot); the 32 / SLOTS sub-lanes stride over the row's entries, then the sub-sums are
// added through shared memory in a fixed order
template <class S>
__global__ void __launch_bounds__(SPMV_T)
sparse_matvec_warp_kernel(const u64* __restrict__ row_ptr, const u64* __restrict__ col_idx,
                          const u64* __restrict__ vals, const u64* __restrict__ v, size_t nrows, size_t ncols,
                          size_t nnz, u64* __restrict__ out, int* __restrict__ bad) {
    constexpr int SUBS = 32 / S::SLOTS;
    __shared__ typename S::Val red[SPMV_T];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane % S::SLOTS, sub = lane / S::SLOTS;
    const size_t row = (size_t)blockIdx.x * (SPMV_T / 32) + warp;
    const bool live = row < nrows;
    typename S::Accum acc;
    S::accum_zero(acc);
    if (live) {
        u64 e0 = row_ptr[row], e1 = row_ptr[row + 1];
        if (e0 > e1 || e1 > nnz) {
            *bad = 2;
            e1 = e0;
        }
        for (u64 e = e0 + sub; e < e1; e += SUBS) {
            const u64 c = col_idx[e];
            if (c >= ncols) {
                *bad = 1;
                continue;
            }
            const typename S::Val a = S::load(vals + e * S::ELEM_U64 + slot * S::SLOT_U64);
            const typename S::Val x = S::load_cached(v + c * S::ELEM_U64 + slot * S::SLOT_U64);
            S::accum_mad(acc, a, x);
        }
    }
    red[threadIdx.x] = S::accum_result(acc);
    __syncwarp();
    if (live && sub == 0) {
        typename S::Val s = red[threadIdx.x];
#pragma unroll
        for (int k = 1; k < SUBS; k++) S::acc(s, red[threadIdx.x + k * S::SLOTS]);
        S::store(out + row * S::ELEM_U64 + slot * S::SLOT_U64, s);
    }
}

// out[i][j] = sum_k a[i][k] * m[k][j]; thread per (j, slot), blockIdx.y = i
template <class S>
__global__ void __launch_bounds__(SPMV_T)
matmat_kernel(const u64* const* __restrict__ a_rows, const u64* const* __restrict__ m_rows,
              u64* const* __restrict__ out_rows, size_t inner, size_t m_ncols) {
    const size_t idx = (size_t)blockIdx.x * SPMV_T + threadIdx.x;
    if (idx >= m_ncols * S::SLOTS) return;
    const size_t j = idx / S::SLOTS;
    const int slot = (int)(idx - j * S::SLOTS);
    const size_t i = blockIdx.y;
    const u64* arow = a_rows[i];
    typename S::Accum acc;
    S::accum_zero(acc);
    for (size_t k = 0; k < inner; k++) {
        const typename S::Val a = S::load_cached(arow + k * S::ELEM_U64 + slot * S::SLOT_U64);
        const typename S::Val x = S::load_cached(m_rows[k] + j * S::ELEM_U64 + slot * S::SLOT_U64);
        S::accum_mad(acc, a, x);
    }
    S::store(out_rows[i] + j * S::ELEM_U64 + slot * S::SLOT_U64, S::accum_result(acc));
}

// Sparse x sparse product, numeric phase (SparseMatrix::checked_mul_mat, sparse_matrix.rs:219-275).  The symbolic phase
// (which entries of row i of A meet which entries of column j of M: the reference's merge join over the stored index
// order) depends on the indices only and is done by the caller (sr_sparse_matmat_symbolic); it yields, per candidate
// output entry c, the pairs [pair_ptr[c], pair_ptr[c+1]) of (entry of A, entry of M).  Thread per (candidate, slot):
// out[c] = sum of the pair products; nonzero[c] |= (some product is not the zero ELEMENT) -- the reference only adds
// non-zero products, which changes nothing in the sum but decides whether the entry exists at all.
template <class S>
__global__ void __launch_bounds__(SPMV_T)
sparse_pairs_kernel(const u64* __restrict__ a_vals, const u64* __restrict__ m_vals, const u64* __restrict__ pair_ptr,
                    const u64* __restrict__ pair_a, const u64* __restrict__ pair_m, size_t ncand,
                    u64* __restrict__ out, int* __restrict__ nonzero) {
    const size_t idx = (size_t)blockIdx.x * SPMV_T + threadIdx.x;
    if (idx >= ncand * S::SLOTS) return;
    const size_t c = idx / S::SLOTS;
    const int slot = (int)(idx - c * S::SLOTS);
    typename S::Val sum = S::zero();
    bool any = false;
    for (u64 e = pair_ptr[c]; e < pair_ptr[c + 1]; e++) {
        const typename S::Val a = S::load_cached(a_vals + pair_a[e] * S::ELEM_U64 + slot * S::SLOT_U64);
        const typename S::Val x = S::load_cached(m_vals + pair_m[e] * S::ELEM_U64 + slot * S::SLOT_U64);
        const typename S::Val prod = S::mul(a, x);
        any = any || !S::is_zero(prod);
        S::acc(sum, prod);
    }
    S::store(out + c * S::ELEM_U64 + slot * S::SLOT_U64, sum);
    if (any) atomicOr(nonzero + c, 1);
}
template <class S>
static cudaError_t sparse_pairs_t(const u64* a_vals, const u64* m_vals, const u64* pair_ptr, const u64* pair_a,
                                  const u64* pair_m, size_t ncand, u64* out, int* nonzero, cudaStream_t st) {
    if (ncand == 0) return cudaSuccess;
    const size_t total = ncand * S::SLOTS;
    sparse_pairs_kernel<S><<<(unsigned)((total + SPMV_T - 1) / SPMV_T), SPMV_T, 0, st>>>(a_vals, m_vals, pair_ptr, pair_a,
                                                                                         pair_m, ncand, out, nonzero);
    return cudaGetLastError();
}
cudaError_t sparse_pairs_launch(int ring, const u64* a_vals, const u64* m_vals, const u64* pair_ptr, const u64* pair_a,
                                const u64* pair_m, size_t ncand, u64* out, int* nonzero, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return sparse_pairs_t<GLSlot>(a_vals, m_vals, pair_ptr, pair_a, pair_m, ncand, out, nonzero, st);
    case RING_BB: return sparse_pairs_t<BBSlot>(a_vals, m_vals, pair_ptr, pair_a, pair_m, ncand, out, nonzero, st);
    case RING_SP: return sparse_pairs_t<SPSlot>(a_vals, m_vals, pair_ptr, pair_a, pair_m, ncand, out, nonzero, st);
    }
    return cudaErrorInvalidValue;
}

// a[e] <- a[e] * r for every element e of the batch; thread per (e, slot)
template <class S>
__global__ void __launch_bounds__(256)
scale_kernel(u64* __restrict__ a, const u64* __restrict__ r, size_t n) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n * S::SLOTS) return;
    const int slot = (int)(idx % S::SLOTS);
    const typename S::Val x = S::load_cached(r + slot * S::SLOT_U64);
    u64* p = a + idx * S::SLOT_U64;
    S::store(p, S::mul(S::load(p), x));
}

template <class S>
static cudaError_t sparse_matvec_t(const u64* row_ptr, const u64* col_idx, const u64* vals, const u64* v, size_t nrows,
                                   size_t ncols, size_t nnz, u64* out, int* bad, cudaStream_t st) {
    if (nrows == 0) return cudaSuccess;
    if (nnz / nrows >= 8) {
        const unsigned grid = (unsigned)((nrows + SPMV_T / 32 - 1) / (SPMV_T / 32));
        sparse_matvec_warp_kernel<S><<<grid, SPMV_T, 0, st>>>(row_ptr, col_idx, vals, v, nrows, ncols, nnz, out, bad);
    } else {
        const size_t total = nrows * S::SLOTS;
        sparse_matvec_kernel<S><<<(unsigned)((total + SPMV_T - 1) / SPMV_T), SPMV_T, 0, st>>>(row_ptr, col_idx, vals, v,
                                                                                              nrows, ncols, nnz, out, bad);
    }
    return cudaGetLastError();
}
cudaError_t sparse_matvec_launch(int ring, const u64* row_ptr, const u64* col_idx, const u64* vals, const u64* v,
                                 size_t nrows, size_t ncols, size_t nnz, u64* out, int* bad, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return sparse_matvec_t<GLSlot>(row_ptr, col_idx, vals, v, nrows, ncols, nnz, out, bad, st);
    case RING_BB: return sparse_matvec_t<BBSlot>(row_ptr, col_idx, vals, v, nrows, ncols, nnz, out, bad, st);
    case RING_SP: return sparse_matvec_t<SPSlot>(row_ptr, col_idx, vals, v, nrows, ncols, nnz, out, bad, st);
    }
    return cudaErrorInvalidValue;
}

template <class S>
static cudaError_t matmat_t(const u64* const* a_rows, const u64* const* m_rows, u64* const* out_rows, size_t a_nrows,
                            size_t inner, size_t m_ncols, cudaStream_t st) {
    if (a_nrows == 0 || m_ncols == 0) return cudaSuccess;
    const size_t total = m_ncols * S::SLOTS;
    for (size_t i0 = 0; i0 < a_nrows; i0 += 65535) {  // gridDim.y limit
        const size_t ny = (a_nrows - i0 < 65535) ? a_nrows - i0 : 65535;
        dim3 grid((unsigned)((total + SPMV_T - 1) / SPMV_T), (unsigned)ny);
        matmat_kernel<S><<<grid, SPMV_T, 0, st>>>(a_rows + i0, m_rows, out_rows + i0, inner, m_ncols);
    }
    return cudaGetLastError();
}
cudaError_t matmat_launch(int ring, const u64* const* a_rows, const u64* const* m_rows, u64* const* out_rows,
                          size_t a_nrows, size_t inner, size_t m_ncols, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return matmat_t<GLSlot>(a_rows, m_rows, out_rows, a_nrows, inner, m_ncols, st);
    case RING_BB: return matmat_t<BBSlot>(a_rows, m_rows, out_rows, a_nrows, inner, m_ncols, st);
    case RING_SP: return matmat_t<SPSlot>(a_rows, m_rows, out_rows, a_nrows, inner, m_ncols, st);
    }
    return cudaErrorInvalidValue;
}

template <class S>
static cudaError_t scale_t(u64* a, const u64* r, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t total = n * S::SLOTS;
    scale_kernel<S><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, r, n);
    return cudaGetLastError();
}
cudaError_t scale_launch(int ring, u64* a, const u64* r, size_t n, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return scale_t<GLSlot>(a, r, n, st);
    case RING_BB: return scale_t<BBSlot>(a, r, n, st);
    case RING_SP: return scale_t<SPSlot>(a, r, n, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
