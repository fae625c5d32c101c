// stark_rings.hpp -- header-only C++ mirror of the reference's operator surface for the hot path,
// over the C ABI in stark_rings_cuda.h.  (The reference is Rust; no Rust toolchain exists in the
// build image, so the compiled-language host side is C++ -- see INTEGRATION.md for the Rust shim.)
//
//   reference (crates/ring/src/cyclotomic_ring/...)              here
//   ring_config.rs:11-35  CyclotomicConfig::{crt,icrt}_in_place    RingConfig<...>::crt_in_place / icrt_in_place
//   coeff_form.rs:31-33   CyclotomicPolyRingGeneral  (RqPoly)      RqPoly<C>   (a batch of n >= 1 elements)
//   ntt_form.rs:25-27     CyclotomicPolyRingNTTGeneral (RqNTT)     RqNTT<C>
//   crt.rs:6-50           CRT / ICRT, elementwise_crt / _icrt      RqPoly::crt(), RqNTT::icrt(), CRT<C>, ICRT<C>
//   ntt_form.rs:159-189   Mul / MulUnchecked                       RqNTT::operator*=, mul_unchecked
//   coeff_form.rs:250-258 Mul (poly_mul + reduce)                  RqPoly::operator*  (fused kernel)
//   linear_algebra/src/matrix.rs:168-183  checked/try_mul_vec      Matrix<C>::checked_mul_vec / try_mul_vec
//   linear_algebra/src/matrix.rs:148-166,207-211  mul_mat, *= R      Matrix<C>::checked_mul_mat / try_mul_mat / operator*=
//   linear_algebra/src/sparse_matrix.rs:17-21,201-217,298-302     SparseMatrix<C> (CSR image of coeffs)
//   coeff_form.rs:154-189 CanonicalSerialize / Deserialize         RqPoly / RqNTT ::serialize, ::deserialize (batch)
//   linear_algebra/src/error.rs:3-8       AlgebraError             DifferentLengths
// Buffers are raw ark-ff limbs (little-endian u64, Montgomery form) in host memory; every call goes
// to libstarkrings_cuda.so -- there is no CPU implementation behind this header.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "stark_rings_cuda.h"

namespace stark_rings {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
// wrong slice length: the reference panics (assert_eq!(coefficients.len(), D))
struct LengthPanic : std::logic_error {
    using std::logic_error::logic_error;
};
// AlgebraError::DifferentLengths(usize, usize)
struct DifferentLengths : std::runtime_error {
    size_t lhs, rhs;
    DifferentLengths(size_t a, size_t b)
        : std::runtime_error("Unexpected different lengths: " + std::to_string(a) + " and " + std::to_string(b)),
          lhs(a), rhs(b) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        if (sr_init(device, &h_) != SR_OK) throw Error("sr_init failed: no usable sm_100 GPU (no CPU fallback)");
    }
    ~Context() { if (h_) sr_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sr_ctx* get() const { return h_; }
    void check(int rc, const char* what) const {
        if (rc == SR_OK) return;
        std::string msg = std::string(what) + ": " + sr_last_error(h_);
        if (rc == SR_ERR_BAD_LENGTH) throw LengthPanic(msg);
        throw Error(msg);
    }
    static Context& global() { static Context c(0); return c; }
private:
    sr_ctx* h_ = nullptr;
};

// CyclotomicConfig<N>: RING = sr_ring id, D = ring dimension, N = u64 limbs per field element
template <int RING, size_t D_, size_t N_, size_t EXT>
struct RingConfig {
    static constexpr int ring = RING;
    static constexpr size_t D = D_, N = N_, LIMBS = D_ * N_, CRT_FIELD_EXTENSION_DEGREE = EXT;
    // one element (D field elements), in place; any other length panics like the reference
    static void crt_in_place(uint64_t* coefficients, size_t n_limbs) {
        if (n_limbs != LIMBS) throw LengthPanic("crt_in_place: wrong slice length");
        Context::global().check(sr_crt_batch(Context::global().get(), RING, coefficients, n_limbs, SR_HOST), "crt");
    }
    static void icrt_in_place(uint64_t* evaluations, size_t n_limbs) {
        if (n_limbs != LIMBS) throw LengthPanic("icrt_in_place: wrong slice length");
        Context::global().check(sr_icrt_batch(Context::global().get(), RING, evaluations, n_limbs, SR_HOST), "icrt");
    }
};
using GoldilocksRingConfig = RingConfig<SR_GOLDILOCKS, 24, 1, 3>;
using BabyBearRingConfig = RingConfig<SR_BABYBEAR, 72, 1, 9>;
using StarkRingConfig = RingConfig<SR_STARK, 16, 4, 1>;

template <class C> struct RqNTT;

// SerializationError::InvalidData (an integer not below the modulus)
struct InvalidData : std::runtime_error {
    using std::runtime_error::runtime_error;
};
template <class C>
std::vector<uint8_t> serialize_limbs(const std::vector<uint64_t>& limbs) {
    auto& c = Context::global();
    std::vector<uint8_t> out(sr_serialized_bytes(C::ring, limbs.size() / C::LIMBS));
    c.check(sr_serialize_batch(c.get(), C::ring, limbs.data(), limbs.size(), out.data(), SR_HOST), "serialize");
    return out;
}
template <class C>
std::vector<uint64_t> deserialize_limbs(const std::vector<uint8_t>& bytes) {
    auto& c = Context::global();
    const size_t per = sr_serialized_bytes(C::ring, 1);
    std::vector<uint64_t> out(bytes.size() / per * C::LIMBS);
    int rc = sr_deserialize_batch(c.get(), C::ring, bytes.data(), bytes.size(), out.data(), SR_HOST);
    if (rc == SR_ERR_INVALID) throw InvalidData(sr_last_error(c.get()));
    c.check(rc, "deserialize");
    return out;
}

// Add / Sub / Neg / Sum (ntt_form.rs:588-626, 640-654 and the same operators of coeff_form.rs): element-wise on the batch
template <class C>
inline void addsub_limbs(int op, std::vector<uint64_t>& a, const std::vector<uint64_t>* b) {
    if (b && b->size() != a.size()) throw LengthPanic("operands differ in length");
    auto& c = Context::global();
    int rc = op == 0 ? sr_add_batch(c.get(), C::ring, a.data(), b->data(), a.size(), SR_HOST)
           : op == 1 ? sr_sub_batch(c.get(), C::ring, a.data(), b->data(), a.size(), SR_HOST)
                     : sr_neg_batch(c.get(), C::ring, a.data(), a.size(), SR_HOST);
    c.check(rc, "add/sub/neg");
}
template <class C>
inline std::vector<uint64_t> sum_limbs(const std::vector<uint64_t>& a) {
    std::vector<uint64_t> out(C::LIMBS);
    auto& c = Context::global();
    c.check(sr_sum_batch(c.get(), C::ring, a.data(), a.size(), out.data(), SR_HOST), "sum");
    return out;
}

// A batch of coefficient-form elements over one flat limb buffer (len() == 1: a single element).
template <class C>
struct RqPoly {
    std::vector<uint64_t> limbs;
    RqPoly() = default;
    explicit RqPoly(std::vector<uint64_t> raw) : limbs(std::move(raw)) {
        if (limbs.size() % C::LIMBS) throw LengthPanic("buffer is not a whole number of ring elements");
    }
    size_t len() const { return limbs.size() / C::LIMBS; }
    static constexpr size_t dimension() { return C::D; }
    RqNTT<C> crt() &&;                       // CRT::crt / elementwise_crt: consumes self, same allocation
    RqPoly operator*(const RqPoly& rhs) const {  // coeff_form.rs Mul == icrt(crt(a) * crt(b)), one kernel
        if (rhs.limbs.size() != limbs.size()) throw LengthPanic("operands differ in length");
        RqPoly out;
        out.limbs.resize(limbs.size());
        auto& c = Context::global();
        c.check(sr_ring_mul_batch(c.get(), C::ring, limbs.data(), rhs.limbs.data(), out.limbs.data(), limbs.size(),
                                  SR_HOST), "ring_mul");
        return out;
    }
    RqPoly& operator+=(const RqPoly& rhs) { addsub_limbs<C>(0, limbs, &rhs.limbs); return *this; }
    RqPoly& operator-=(const RqPoly& rhs) { addsub_limbs<C>(1, limbs, &rhs.limbs); return *this; }
    RqPoly operator+(const RqPoly& rhs) const { RqPoly t(*this); t += rhs; return t; }
    RqPoly operator-(const RqPoly& rhs) const { RqPoly t(*this); t -= rhs; return t; }
    RqPoly operator-() const { RqPoly t(*this); addsub_limbs<C>(2, t.limbs, nullptr); return t; }
    RqPoly sum() const { return RqPoly(sum_limbs<C>(limbs)); }
    bool operator==(const RqPoly& o) const { return limbs == o.limbs; }
    // CanonicalSerialize of the batch (coeff_form.rs:154-167): standard-form little-endian bytes, no length prefix
    std::vector<uint8_t> serialize() const { return serialize_limbs<C>(limbs); }
    static RqPoly deserialize(const std::vector<uint8_t>& bytes) { return RqPoly(deserialize_limbs<C>(bytes)); }
};

template <class C>
struct RqNTT {
    std::vector<uint64_t> limbs;
    RqNTT() = default;
    explicit RqNTT(std::vector<uint64_t> raw) : limbs(std::move(raw)) {
        if (limbs.size() % C::LIMBS) throw LengthPanic("buffer is not a whole number of ring elements");
    }
    size_t len() const { return limbs.size() / C::LIMBS; }
    RqPoly<C> icrt() && {                    // ICRT::icrt / elementwise_icrt
        auto& c = Context::global();
        c.check(sr_icrt_batch(c.get(), C::ring, limbs.data(), limbs.size(), SR_HOST), "icrt");
        return RqPoly<C>(std::move(limbs));
    }
    RqNTT& operator*=(const RqNTT& rhs) {    // ntt_form.rs:159-175, slot-wise
        if (rhs.limbs.size() != limbs.size()) throw LengthPanic("operands differ in length");
        auto& c = Context::global();
        c.check(sr_ntt_mul_batch(c.get(), C::ring, limbs.data(), rhs.limbs.data(), limbs.size(), SR_HOST), "ntt_mul");
        return *this;
    }
    RqNTT operator*(const RqNTT& rhs) const { RqNTT t(*this); t *= rhs; return t; }
    RqNTT mul_unchecked(const RqNTT& rhs) const { return *this * rhs; }  // ntt_form.rs:177-189
    RqNTT& operator+=(const RqNTT& rhs) { addsub_limbs<C>(0, limbs, &rhs.limbs); return *this; }
    RqNTT& operator-=(const RqNTT& rhs) { addsub_limbs<C>(1, limbs, &rhs.limbs); return *this; }
    RqNTT operator+(const RqNTT& rhs) const { RqNTT t(*this); t += rhs; return t; }
    RqNTT operator-(const RqNTT& rhs) const { RqNTT t(*this); t -= rhs; return t; }
    RqNTT operator-() const { RqNTT t(*this); addsub_limbs<C>(2, t.limbs, nullptr); return t; }
    RqNTT sum() const { return RqNTT(sum_limbs<C>(limbs)); }
    static constexpr size_t dimension() { return C::D; }
    bool operator==(const RqNTT& o) const { return limbs == o.limbs; }
    // every element of the batch *= r, r one element (the body of MulAssign<&R> for the matrix types)
    void scale(const RqNTT& r) {
        if (r.len() != 1) throw LengthPanic("the scalar must be one ring element");
        auto& c = Context::global();
        c.check(sr_ntt_scale_batch(c.get(), C::ring, limbs.data(), limbs.size(), r.limbs.data(), SR_HOST), "scale");
    }
    std::vector<uint8_t> serialize() const { return serialize_limbs<C>(limbs); }
    static RqNTT deserialize(const std::vector<uint8_t>& bytes) { return RqNTT(deserialize_limbs<C>(bytes)); }
};

template <class C>
RqNTT<C> RqPoly<C>::crt() && {
    auto& c = Context::global();
    c.check(sr_crt_batch(c.get(), C::ring, limbs.data(), limbs.size(), SR_HOST), "crt");
    return RqNTT<C>(std::move(limbs));
}

template <class C> struct CRT { static RqNTT<C> elementwise_crt(RqPoly<C>&& v) { return std::move(v).crt(); } };
template <class C> struct ICRT { static RqPoly<C> elementwise_icrt(RqNTT<C>&& v) { return std::move(v).icrt(); } };

// Matrix { nrows, ncols, vals: Vec<Vec<R>> } with R = RqNTT (each row its own allocation)
template <class C>
struct Matrix {
    size_t nrows = 0, ncols = 0;
    std::vector<RqNTT<C>> vals;
    explicit Matrix(std::vector<RqNTT<C>> rows) : vals(std::move(rows)) {
        nrows = vals.size();
        ncols = nrows ? vals[0].len() : 0;
        for (auto& r : vals) if (r.len() != ncols) throw std::invalid_argument("ragged matrix");
    }
    std::optional<RqNTT<C>> checked_mul_vec(const RqNTT<C>& v) const {  // None when ncols != v.len()
        std::vector<const uint64_t*> ptrs(nrows);
        for (size_t i = 0; i < nrows; i++) ptrs[i] = vals[i].limbs.data();
        RqNTT<C> out;
        out.limbs.resize(nrows * C::LIMBS);
        auto& c = Context::global();
        int rc = sr_matvec(c.get(), C::ring, ptrs.data(), nrows, ncols, v.limbs.data(), v.limbs.size(),
                           out.limbs.data(), SR_HOST);
        if (rc == SR_ERR_BAD_LENGTH) return std::nullopt;
        c.check(rc, "matvec");
        return out;
    }
    RqNTT<C> try_mul_vec(const RqNTT<C>& v) const {  // Err(DifferentLengths(ncols, v.len()))
        auto r = checked_mul_vec(v);
        if (!r) throw DifferentLengths(ncols, v.len());
        return *r;
    }
    // matrix.rs:148-166: out[i][j] = sum_k self[i][k] * m[k][j]; None when self.ncols != m.nrows
    std::optional<Matrix> checked_mul_mat(const Matrix& m) const {
        if (ncols != m.nrows) return std::nullopt;
        std::vector<RqNTT<C>> out(nrows);
        std::vector<const uint64_t*> pa(nrows), pm(m.nrows);
        std::vector<uint64_t*> po(nrows);
        for (size_t i = 0; i < nrows; i++) {
            out[i].limbs.resize(m.ncols * C::LIMBS);
            pa[i] = vals[i].limbs.data();
            po[i] = out[i].limbs.data();
        }
        for (size_t k = 0; k < m.nrows; k++) pm[k] = m.vals[k].limbs.data();
        auto& c = Context::global();
        c.check(sr_matmat(c.get(), C::ring, pa.data(), nrows, ncols, pm.data(), m.nrows, m.ncols, po.data(), SR_HOST),
                "matmat");
        return Matrix(std::move(out));
    }
    Matrix try_mul_mat(const Matrix& m) const {  // matrix.rs:185-188
        auto r = checked_mul_mat(m);
        if (!r) throw DifferentLengths(ncols, m.nrows);
        return std::move(*r);
    }
    Matrix& operator*=(const RqNTT<C>& r) {  // matrix.rs:207-211
        for (auto& row : vals) row.scale(r);
        return *this;
    }
};

// SparseMatrix { nrows, ncols, coeffs: Vec<Vec<(R, usize)>> } (sparse_matrix.rs:17-21) held as the CSR image of coeffs
template <class C>
struct SparseMatrix {
    size_t nrows = 0, ncols = 0;
    std::vector<uint64_t> row_ptr, col_idx;  // row_ptr: nrows + 1 entry offsets
    RqNTT<C> vals;                           // nnz elements in row order
    SparseMatrix(size_t nr, size_t nc, const std::vector<std::vector<std::pair<std::vector<uint64_t>, size_t>>>& coeffs)
        : nrows(nr), ncols(nc), row_ptr(nr + 1, 0) {
        for (size_t i = 0; i < coeffs.size(); i++) {
            for (auto& e : coeffs[i]) {
                vals.limbs.insert(vals.limbs.end(), e.first.begin(), e.first.end());
                col_idx.push_back(e.second);
            }
            row_ptr[i + 1] = col_idx.size();
        }
        for (size_t i = coeffs.size(); i < nr; i++) row_ptr[i + 1] = col_idx.size();  // pad_rows
    }
    std::optional<RqNTT<C>> checked_mul_vec(const RqNTT<C>& v) const {  // sparse_matrix.rs:201-212
        RqNTT<C> out;
        out.limbs.resize(nrows * C::LIMBS);
        auto& c = Context::global();
        int rc = sr_sparse_matvec(c.get(), C::ring, nrows, ncols, row_ptr.data(), col_idx.data(), vals.limbs.data(),
                                  v.limbs.data(), v.limbs.size(), out.limbs.data(), SR_HOST);
        if (rc == SR_ERR_BAD_LENGTH) return std::nullopt;
        c.check(rc, "sparse_matvec");
        return out;
    }
    RqNTT<C> try_mul_vec(const RqNTT<C>& v) const {  // sparse_matrix.rs:214-217
        auto r = checked_mul_vec(v);
        if (!r) throw DifferentLengths(ncols, v.len());
        return *r;
    }
    SparseMatrix& operator*=(const RqNTT<C>& r) {  // sparse_matrix.rs:298-302
        if (vals.len()) vals.scale(r);
        return *this;
    }
};

}  // namespace stark_rings
