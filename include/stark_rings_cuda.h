/* stark_rings_cuda.h -- C ABI of libstarkrings_cuda.so (sm_100a).
 *
 * Drop-in boundary for the hot path of NethermindEth/stark-rings: the bodies of the Rust items
 * cited below become calls into these entry points (INTEGRATION.md shows the build.rs / extern "C"
 * shim).  All buffers are the reference's own memory layout: dense arrays of ring elements, each
 * element D field elements, each field element N little-endian u64 limbs holding x * 2^(64 N) mod p
 * (ark-ff MontBackend), canonical (< p).
 *
 *   ring            p                          D   N   limbs/element   bytes/element
 *   SR_GOLDILOCKS   2^64 - 2^32 + 1            24  1   24              192
 *   SR_BABYBEAR     15 * 2^27 + 1              72  1   72              576
 *   SR_STARK        2^251 + 17 * 2^192 + 1     16  4   64              512
 *
 * Lengths are given in u64 LIMBS of the flat slice (what `Flatten::flatten_to_coeffs`,
 * flatten.rs:10-18, exposes), so a wrong slice length is detectable: a length that is not a
 * multiple of limbs/element returns SR_ERR_BAD_LENGTH where the reference panics
 * (assert_eq!(coefficients.len(), D): goldilocks/ntt.rs:136,241, babybear/ntt.rs:144,239,
 * stark_prime/ntt.rs:122,246).
 *
 * `loc` says where the caller's buffers live: SR_HOST (any host memory; pinned memory from
 * sr_host_alloc moves fastest) or SR_DEVICE (device memory of the context's GPU, e.g. from
 * sr_dev_alloc; must be 16-byte aligned).  Host calls are synchronous; device calls are enqueued
 * on the context's stream (sr_sync / sr_set_stream).
 *
 * There is no CPU fallback: without a usable CUDA device sr_init fails with SR_ERR_CUDA.
 * Thread safety: a context serialises its own calls with an internal mutex; use one context per
 * host thread for concurrency.  The library never frees or retains caller pointers past return.
 */
#ifndef STARK_RINGS_CUDA_H
#define STARK_RINGS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sr_ctx sr_ctx;

enum sr_status {
    SR_OK = 0,
    SR_ERR_BAD_LENGTH = 1,  /* slice length not a multiple of limbs/element, or ncols != v.len() */
    SR_ERR_CUDA = 2,        /* CUDA runtime error; text in sr_last_error */
    SR_ERR_INVALID = 3,     /* null/misaligned pointer, unknown ring/loc */
    SR_ERR_NOMEM = 4
};
enum sr_ring { SR_GOLDILOCKS = 0, SR_BABYBEAR = 1, SR_STARK = 2 };
enum sr_loc { SR_HOST = 0, SR_DEVICE = 1 };

/* ---- context ---------------------------------------------------------------------------- */
int sr_init(int device, sr_ctx** out);
int sr_destroy(sr_ctx* ctx);
const char* sr_last_error(sr_ctx* ctx);       /* message of the last failing call on ctx */
const char* sr_version(void);
int sr_set_stream(sr_ctx* ctx, void* cuda_stream); /* enqueue SR_DEVICE calls on this cudaStream_t (NULL = default stream) */
int sr_reset_stream(sr_ctx* ctx);                  /* back to the context's own non-blocking stream */
int sr_sync(sr_ctx* ctx);
/* Pipelined products (off by default).  When on, the mat-vec / commitment kernels of this context are launched with
 * programmatic dependent launch: the column loop of a product may start while the previous kernel on the stream is
 * still in its tail (the cross-CTA reduction and, in a commitment, the NVLink hand-off), which hides that tail in a
 * sequence of products.  Only the column loop runs early, and it only READS the matrix rows and the vector: turn
 * this on when those inputs are resident (not written by the kernel enqueued immediately before the product). */
int sr_set_pipelined(sr_ctx* ctx, int on);
size_t sr_elem_limbs(int ring);               /* 24 / 72 / 64; 0 for an unknown ring */
uint64_t sr_kernel_launches(sr_ctx* ctx);     /* kernels launched by this context so far */

/* ---- memory ------------------------------------------------------------------------------ */
int sr_dev_alloc(sr_ctx* ctx, size_t bytes, void** dptr);
int sr_dev_free(sr_ctx* ctx, void* dptr);
int sr_host_alloc(sr_ctx* ctx, size_t bytes, void** hptr); /* pinned host memory */
int sr_host_free(sr_ctx* ctx, void* hptr);
int sr_h2d(sr_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);  /* async on ctx stream */
int sr_d2h(sr_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);  /* async on ctx stream */

/* ---- timing on the context's stream (CUDA events) ---------------------------------------- */
int sr_timer_start(sr_ctx* ctx);
int sr_timer_stop(sr_ctx* ctx, float* ms);    /* synchronises; elapsed since sr_timer_start */

/* Measured integer multiply-add peaks of the context's device, in 10^12 operations per second: tops3[0] 32-bit IMAD,
 * tops3[1] IMAD.WIDE.U32 (32 x 32 + 64 -> 64), tops3[2] the carry-chained IMAD.WIDE.U32.X the multi-limb kernels use.
 * The roofline denominator of the Starknet-prime kernels (bench.py measures it in the run it reports).  Synchronous,
 * about 10 ms. */
int sr_imad_peak(sr_ctx* ctx, double* tops3);

/* ---- batched conversions and products ------------------------------------------------------
 * Replaces  CRT::elementwise_crt / ICRT::elementwise_icrt            (crt.rs:10-25, 34-49)
 *           CyclotomicConfig::{crt_in_place, icrt_in_place}           (ring_config.rs:27,34)
 *           impl_crt_icrt_for_a_ring!: crt(self) / icrt(self)         (crt.rs:52-77), n = 1 element
 * In place: `buf` holds n_limbs limbs of coefficient-form (crt) / NTT-form (icrt) elements and is
 * overwritten with the other form, exactly as the reference retypes the same allocation. */
int sr_crt_batch(sr_ctx* ctx, int ring, uint64_t* buf, size_t n_limbs, int loc);
int sr_icrt_batch(sr_ctx* ctx, int ring, uint64_t* buf, size_t n_limbs, int loc);

/* Replaces  Mul / MulAssign / MulUnchecked for CyclotomicPolyRingNTTGeneral
 *           (ntt_form.rs:159-189, 213-225, 521-550): a[i] <- a[i] * b[i] slot-wise, NTT form. */
int sr_ntt_mul_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, const uint64_t* b, size_t n_limbs, int loc);

/* Replaces  Add / AddAssign / Sub / SubAssign / Neg / Sum for CyclotomicPolyRingNTTGeneral (ntt_form.rs:588-601,
 *           603-626, 640-654) and the same operators of CyclotomicPolyRingGeneral (coeff_form.rs): both forms add field
 *           element by field element, so these serve RqPoly and RqNTT batches alike.
 * a[i] <- a[i] + b[i] / a[i] - b[i] / -a[i] in place; sr_sum_batch folds n elements into one (`out`: one element,
 * ZERO for an empty slice). */
int sr_add_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, const uint64_t* b, size_t n_limbs, int loc);
int sr_sub_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, const uint64_t* b, size_t n_limbs, int loc);
int sr_neg_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, size_t n_limbs, int loc);
int sr_sum_batch(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint64_t* out, int loc);

/* Replaces  Mul for CyclotomicPolyRingGeneral (coeff_form.rs:54-67, 250-258): out[i] = a[i] * b[i]
 * in F_p[X]/Phi, coefficient form in and out; computed as icrt(crt(a) * crt(b)) in one kernel
 * (the identity the reference pins in test_mul_crt, e.g. goldilocks/mod.rs:231-247).
 * out may alias a or b. */
int sr_ring_mul_batch(sr_ctx* ctx, int ring, const uint64_t* a, const uint64_t* b, uint64_t* out,
                      size_t n_limbs, int loc);

/* ---- ring matrix x vector (Ajtai-style commitment) -------------------------------------------
 * Replaces  Matrix<R>::checked_mul_vec / try_mul_vec / Mul<&[R]>  with R = RqNTT
 *           (linear_algebra/src/matrix.rs:168-183, 199-205).
 * rows[i] points at row i: ncols NTT-form elements (each row is its own allocation, as in
 * Matrix.vals: Vec<Vec<R>>, matrix.rs:17-21); v holds v_limbs limbs; out receives nrows elements.
 * ncols * limbs/element != v_limbs  ->  SR_ERR_BAD_LENGTH (the reference returns None /
 * AlgebraError::DifferentLengths(ncols, v.len()), matrix.rs:169-171,180-183). */
int sr_matvec(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols,
              const uint64_t* v, size_t v_limbs, uint64_t* out, int loc);

/* Column-sharded commitment, one rank's share: the same product over this rank's
 * columns only, leaving nrows partial elements in `partial_out` (device or host).  The partials of
 * all ranks are gathered by the caller (NCCL all-gather of raw limbs) and summed mod p on rank 0: */
int sr_matvec_partial(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols,
                      const uint64_t* v, size_t v_limbs, uint64_t* partial_out, int loc);
/* out[i] = sum_r gathered[r * nrows + i]  (mod p, NTT form; an NCCL sum cannot reduce mod p). */
int sr_modsum_partials(sr_ctx* ctx, int ring, const uint64_t* gathered, size_t nranks, size_t nrows,
                       uint64_t* out, int loc);

/* ---- column-sharded commitment over NVLink peer memory -----------------------------------------
 * One process per GPU.  The ROOT rank creates a mailbox (device memory) and exports its CUDA IPC handle
 * (SR_IPC_HANDLE_BYTES bytes), which the caller ships to the other ranks out of band (MPI, torch.distributed
 * all_gather_object, a pipe ...); they map it with sr_mailbox_open.  Per commitment (epoch = 1, 2, 3 ... the same on
 * every rank) every rank launches ONE kernel per four matrix rows and nothing else:
 *   other ranks: sr_commit_send    its share of the columns -> nrows partial elements, written by the tail of the
 *                                  product kernel straight into the root's mailbox over NVLink, followed by a
 *                                  release-flag; no NCCL call, no host synchronisation, no extra kernel launch
 *   root:        sr_commit_root    the same for the root's own share; the tail of the root's kernel then acquires the
 *                                  flags of all ranks and adds the partials mod p (an NCCL sum cannot reduce mod p)
 *                                  into `out` (device memory of the root)
 * Alternatively the root calls sr_commit_send like everybody else and then sr_commit_reduce, a separate kernel that
 * does the acquire + modular sum (what a caller uses when the reduction must run on another stream, and what the
 * single-GPU emulation of several ranks uses, where the root's share cannot wait for shares enqueued after it).
 * epoch = 0 selects device-resident epochs: each rank's kernels count their own commitments, so a step issues
 * identical launches every time and can be captured in a CUDA graph (every rank must then make exactly one
 * sr_commit_send / sr_commit_root call per commitment).  Do not mix the two modes on one mailbox.
 * All calls are asynchronous on the context's stream.  A wait that exceeds the mailbox's budget (default 4 s,
 * sr_mailbox_set_timeout; a lost peer) sets an error flag, readable with sr_mailbox_error, instead of hanging the
 * GPU: a writer that timed out does not overwrite the slot and does not publish, a root that timed out fills `out`
 * with all-ones limbs (not a canonical residue) so that the value cannot pass for a commitment.  Start the first
 * commitment only after every rank has opened the mailbox (a barrier), or raise the budget above the start-up
 * skew.  The reference has no counterpart (single process); single-GPU semantics are those of
 * Matrix::checked_mul_vec (matrix.rs:168-178).
 * CUDA graphs: the kernels read the context's row-pointer table; a context whose calls were captured must keep
 * serving that matrix (use one context per captured matrix). */
typedef struct sr_mailbox sr_mailbox;
#define SR_IPC_HANDLE_BYTES 64
int sr_mailbox_create(sr_ctx* ctx, int ring, size_t nrows_max, int nranks, sr_mailbox** out,
                      unsigned char* handle_out /* SR_IPC_HANDLE_BYTES, may be NULL */);
int sr_mailbox_open(sr_ctx* ctx, int ring, size_t nrows_max, int nranks, const unsigned char* handle,
                    sr_mailbox** out);
int sr_mailbox_destroy(sr_ctx* ctx, sr_mailbox* box);
int sr_mailbox_error(sr_ctx* ctx, sr_mailbox* box, int* timed_out); /* synchronises the context's stream */
int sr_mailbox_set_timeout(sr_ctx* ctx, sr_mailbox* box, uint64_t nanoseconds); /* budget of one in-kernel wait */
int sr_commit_send(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols, const uint64_t* v,
                   size_t v_limbs, sr_mailbox* root_box, int rank, uint64_t epoch);
int sr_commit_root(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols, const uint64_t* v,
                   size_t v_limbs, sr_mailbox* own_box, int rank, uint64_t epoch, uint64_t* out);
int sr_commit_reduce(sr_ctx* ctx, int ring, sr_mailbox* own_box, size_t nrows, uint64_t epoch, uint64_t* out);

/* ---- coefficient-form helpers next to the hot path (out of place, out != in) ----
 * sr_reduce_batch replaces CyclotomicConfig::reduce_in_place (goldilocks/mod.rs:75-98, babybear/mod.rs:87-110,
 * stark_prime/mod.rs:40-47) on a batch: `in` holds polynomials of coeffs_per_poly field elements each
 * (D <= coeffs_per_poly <= 2D, shorter inputs are the caller's zero padding), `out` receives D per polynomial.
 * sr_rot_batch replaces Cyclotomic::rot (multiplication by X; goldilocks/mod.rs:138-149, babybear/mod.rs:150-161,
 * stark_prime/mod.rs:87-95). */
int sr_reduce_batch(sr_ctx* ctx, int ring, const uint64_t* in, size_t in_limbs, size_t coeffs_per_poly,
                    uint64_t* out, int loc);
int sr_rot_batch(sr_ctx* ctx, int ring, const uint64_t* in, uint64_t* out, size_t n_limbs, int loc);

/* ---- balanced gadget decomposition feeding the commitment (Goldilocks and BabyBear) ----
 * sr_gadget_decompose replaces GadgetDecompose for &[R] / Vec<R> with R = RqPoly (balanced_decomposition/mod.rs:163-175,
 * per element coeff_form.rs:588-606, per coefficient decompose_balanced_in_place mod.rs:62-103): `in` holds n
 * coefficient-form elements, `out` receives n * padding_size elements, out[j * padding_size + t] = t-th digit element of
 * in[j]; digits lie in [-b/2, b/2].  The basis b = b_lo + 2^64 b_hi must be even and >= 2 (the reference asserts
 * this); bases >= 2^62 are not supported.  A decomposition longer than padding_size returns SR_ERR_BAD_LENGTH
 * (the reference indexes out of bounds and panics).  sr_gadget_recompose is the inverse (mod.rs:177-190):
 * n * padding_size digit elements -> n elements.  Both are out of place and synchronous. */
int sr_gadget_decompose(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint64_t b_lo, uint64_t b_hi,
                        size_t padding_size, uint64_t* out, int loc);
int sr_gadget_recompose(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint64_t b_lo, uint64_t b_hi,
                        size_t padding_size, uint64_t* out, int loc);

/* ---- the callers' other linear maps over NTT-form elements ------------------------------
 * sr_sparse_matvec replaces SparseMatrix<R>::checked_mul_vec / try_mul_vec / Mul<&[R]> with R = RqNTT
 * (linear_algebra/src/sparse_matrix.rs:201-217, 278-286): out[i] = sum over (r, j) in coeffs[i] of r * v[j].
 * The matrix crosses the boundary as the CSR image of coeffs: Vec<Vec<(R, usize)>> (sparse_matrix.rs:17-21):
 * row_ptr[nrows + 1] entry offsets (row_ptr[0] = 0, non-decreasing, row_ptr[nrows] = nnz), col_idx[nnz], and vals =
 * nnz NTT-form elements in row order; all of them, v and out live in `loc`.  ncols != v.len() returns
 * SR_ERR_BAD_LENGTH (the reference returns None / DifferentLengths(ncols, v.len())); a column index >= ncols
 * returns SR_ERR_INVALID (the reference panics on v[*i]).  Synchronous (the index check reads a device flag). */
int sr_sparse_matvec(sr_ctx* ctx, int ring, size_t nrows, size_t ncols, const uint64_t* row_ptr,
                     const uint64_t* col_idx, const uint64_t* vals, const uint64_t* v, size_t v_limbs, uint64_t* out,
                     int loc);

/* Sparse x sparse product: SparseMatrix<R>::checked_mul_mat / try_mul_mat / Mul<&SparseMatrix<R>> with R = RqNTT
 * (sparse_matrix.rs:219-275), in two phases.  sr_sparse_matmat_symbolic (host only, no context: indices are host
 * metadata) is the reference's own structure pass: the columns of M gathered in row order and, for every (row i of A,
 * column j of M), the merge join of the two index lists in their stored order.  Call it once with the output arrays
 * NULL to size them (*ncand candidates, *npairs matches), then again to fill cand_row / cand_col [ncand],
 * pair_ptr [ncand + 1], pair_a / pair_m [npairs] (entry numbers into A's and M's vals).  sr_sparse_matmat_values is
 * the arithmetic: out_vals[c] = the sum of candidate c's products, nonzero_host[c] = 1 iff one of them is not the zero
 * element; the reference keeps exactly the candidates with nonzero = 1, in (row, column) order (:249-262).  a_vals,
 * m_vals and out_vals live in `loc`; the pair arrays and nonzero_host are host arrays.  Synchronous.
 * A column index of M that is >= m_ncols returns SR_ERR_INVALID (the reference panics). */
int sr_sparse_matmat_symbolic(size_t a_nrows, const uint64_t* a_row_ptr, const uint64_t* a_col_idx, size_t m_nrows,
                              size_t m_ncols, const uint64_t* m_row_ptr, const uint64_t* m_col_idx, size_t* ncand,
                              size_t* npairs, uint64_t* cand_row, uint64_t* cand_col, uint64_t* pair_ptr,
                              uint64_t* pair_a, uint64_t* pair_m);
int sr_sparse_matmat_values(sr_ctx* ctx, int ring, const uint64_t* a_vals, const uint64_t* m_vals, size_t ncand,
                            const uint64_t* pair_ptr, const uint64_t* pair_a, const uint64_t* pair_m,
                            uint64_t* out_vals, int* nonzero_host, int loc);

/* sr_matmat replaces Matrix<R>::checked_mul_mat / try_mul_mat / Mul<&Matrix<R>> with R = RqNTT
 * (linear_algebra/src/matrix.rs:148-166, 185-197): out[i][j] = sum_k a[i][k] * m[k][j].  a_rows / m_rows / out_rows
 * are host arrays of row pointers (rows are separate allocations, matrix.rs:17-21) whose targets live in `loc`.
 * a_ncols != m_nrows returns SR_ERR_BAD_LENGTH (None / DifferentLengths(self.ncols, m.nrows)). */
int sr_matmat(sr_ctx* ctx, int ring, const uint64_t* const* a_rows, size_t a_nrows, size_t a_ncols,
              const uint64_t* const* m_rows, size_t m_nrows, size_t m_ncols, uint64_t* const* out_rows, int loc);

/* sr_ntt_scale_batch replaces MulAssign<&R> for Matrix<R> / SparseMatrix<R> (matrix.rs:207-211,
 * sparse_matrix.rs:298-302) on one row / on the vals array: every NTT-form element of a_inout *= r (one element). */
int sr_ntt_scale_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, size_t n_limbs, const uint64_t* r, int loc);

/* ---- canonical (de)serialization of batches on the device -------------------------------------
 * Replaces, for a batch, CanonicalSerialize / CanonicalDeserialize of RqPoly / RqNTT (coeff_form.rs:154-189,
 * ntt_form.rs:24), i.e. ark-serialize 0.4 on each field element in memory order: the little-endian bytes of the
 * STANDARD-form integer, ceil(modulus bits / 8) = 8 / 4 / 32 bytes per field element (Goldilocks / BabyBear / Starknet
 * prime), no length prefix (the u64 prefix of a Vec is the caller's).  sr_deserialize_batch returns SR_ERR_INVALID when
 * an integer is not below the modulus (SerializationError::InvalidData).  Out of place.  The reference holds no
 * serialized test vector: the byte format is the restated ark-serialize rule ("parity unpinned"). */
size_t sr_serialized_bytes(int ring, size_t n_elems);
int sr_serialize_batch(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint8_t* out_bytes, int loc);
int sr_deserialize_batch(sr_ctx* ctx, int ring, const uint8_t* in_bytes, size_t n_bytes, uint64_t* out, int loc);

/* ---- per-prime entry points (what each model module binds) --------------------------------- */
#define SR_DECLARE_RING(tag)                                                                          \
    int sr_##tag##_crt_batch(sr_ctx* ctx, uint64_t* buf, size_t n_limbs, int loc);                   \
    int sr_##tag##_icrt_batch(sr_ctx* ctx, uint64_t* buf, size_t n_limbs, int loc);                  \
    int sr_##tag##_ntt_mul_batch(sr_ctx* ctx, uint64_t* a_inout, const uint64_t* b, size_t n_limbs,  \
                                 int loc);                                                            \
    int sr_##tag##_ring_mul_batch(sr_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out,  \
                                  size_t n_limbs, int loc);                                           \
    int sr_##tag##_matvec(sr_ctx* ctx, const uint64_t* const* rows, size_t nrows, size_t ncols,      \
                          const uint64_t* v, size_t v_limbs, uint64_t* out, int loc);

SR_DECLARE_RING(gl) /* goldilocks::{RqPoly,RqNTT},  models/goldilocks/mod.rs:31-32,100-118,151 */
SR_DECLARE_RING(bb) /* babybear::{RqPoly,RqNTT},    models/babybear/mod.rs:30-31,112-130,163   */
SR_DECLARE_RING(sp) /* stark_prime::{RqPoly,RqNTT}, models/stark_prime/mod.rs:29-30,49-67,97   */

#ifdef __cplusplus
}
#endif
#endif /* STARK_RINGS_CUDA_H */
