// Per-ring CRT-slot traits shared by the mat-vec kernels (sr_matvec.cu) and the sparse / matrix-matrix kernels
// (sr_sparse.cu): one Val = one slot of an NTT-form element (Fq3 / Fq9 / Fq), with the slot product of
// ntt_form.rs:159-175 on raw Montgomery limbs and the slot addition of ntt_form.rs:588-601.
#pragma once
#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "sp_ring.cuh"

namespace sr {

struct GLSlot {
    static constexpr int SLOTS = 8, SLOT_U64 = 3, ELEM_U64 = 24;
    struct Val { u64 c[3]; };
    SR_D static Val load(const u64* p) { Val v; v.c[0] = __ldcs(p); v.c[1] = __ldcs(p + 1); v.c[2] = __ldcs(p + 2); return v; }
    SR_D static Val load_cached(const u64* p) { Val v; v.c[0] = p[0]; v.c[1] = p[1]; v.c[2] = p[2]; return v; }
    // coherent load (bypasses L1): data another CTA or another GPU has just written
    SR_D static Val load_cv(const u64* p) { Val v; v.c[0] = __ldcv(p); v.c[1] = __ldcv(p + 1); v.c[2] = __ldcv(p + 2); return v; }
    SR_D static void store(u64* p, const Val& v) { p[0] = v.c[0]; p[1] = v.c[1]; p[2] = v.c[2]; }
    SR_D static void store_poison(u64* p) { p[0] = p[1] = p[2] = ~0ull; }
    SR_D static Val zero() { Val v; v.c[0] = v.c[1] = v.c[2] = 0; return v; }
    SR_D static bool is_zero(const Val& v) { return (v.c[0] | v.c[1] | v.c[2]) == 0; }
    // gl:: arithmetic is weak-form (gl_ring.cuh); values stored in Val are kept canonical
    SR_D static Val mul(const Val& a, const Val& b) {
        Val z;
        gl::slot_mul<gl::root_exp(1), 128>(z.c, a.c, b.c);
#pragma unroll
        for (int i = 0; i < 3; i++) z.c[i] = gl::canon(z.c[i]);
        return z;
    }
    SR_D static void acc(Val& s, const Val& x) {
#pragma unroll
        for (int i = 0; i < 3; i++) s.c[i] = gl::canon(gl::add(s.c[i], x.c[i]));
    }
    // mul_lazy / finish: a product that may leave a linear factor to be applied once to a sum of products
    SR_D static Val mul_lazy(const Val& a, const Val& b) { return mul(a, b); }
    SR_D static void finish(Val&) {}
    // Accum: a running sum of slot products.  Goldilocks keeps the three output coefficients as UNREDUCED 160-bit
    // carry-save sums (gl::Acc) and reduces once per result (valid for up to 2^30 products per sum).
    struct Accum { gl::Acc d[3]; };
    SR_D static void accum_zero(Accum& A) { gl::acc_zero(A.d[0]); gl::acc_zero(A.d[1]); gl::acc_zero(A.d[2]); }
    SR_D static void accum_mad(Accum& A, const Val& a, const Val& x) {
        const u64 r1 = gl::mul_pow2<gl::root_exp(1)>(x.c[1]), r2 = gl::mul_pow2<gl::root_exp(1)>(x.c[2]);  // u^3 = r
        gl::acc_mad(A.d[0], a.c[0], x.c[0]); gl::acc_mad(A.d[0], a.c[1], r2); gl::acc_mad(A.d[0], a.c[2], r1);
        gl::acc_mad(A.d[1], a.c[0], x.c[1]); gl::acc_mad(A.d[1], a.c[1], x.c[0]); gl::acc_mad(A.d[1], a.c[2], r2);
        gl::acc_mad(A.d[2], a.c[0], x.c[2]); gl::acc_mad(A.d[2], a.c[1], x.c[1]); gl::acc_mad(A.d[2], a.c[2], x.c[0]);
    }
    // Prep: the vector operand prepared once and shared by all rows it multiplies (nothing to prepare here)
    typedef Val Prep;
    SR_D static Prep prep(const Val& x) { return x; }
    SR_D static void accum_mad_p(Accum& A, const Val& a, const Prep& x) { accum_mad(A, a, x); }
    SR_D static Val accum_result(const Accum& A) {  // Montgomery layout: products of raw limbs carry 2^-64 = 2^128
        Val z;
#pragma unroll
        for (int i = 0; i < 3; i++) z.c[i] = gl::acc_reduce_m128(A.d[i]);
        return z;
    }
};
// Accum for the rings without a lazy representation: the running sum is a Val (mul_lazy / acc / finish)
template <class S>
struct ValAccum {
    typename S::Val v;
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
    SR_D static Val load_cv(const u64* p) {
        Val v;
#pragma unroll
        for (int i = 0; i < 9; i++) v.c[i] = (u32)__ldcv(p + i);
        return v;
    }
    SR_D static void store(u64* p, const Val& v) {
#pragma unroll
        for (int i = 0; i < 9; i++) p[i] = v.c[i];
    }
    SR_D static void store_poison(u64* p) {
#pragma unroll
        for (int i = 0; i < 9; i++) p[i] = ~0ull;
    }
    SR_D static Val zero() {
        Val v;
#pragma unroll
        for (int i = 0; i < 9; i++) v.c[i] = 0;
        return v;
    }
    SR_D static bool is_zero(const Val& v) {
        u32 o = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) o |= v.c[i];
        return o == 0;
    }
    SR_D static Val mul(const Val& a, const Val& b) { Val z; bb::slot_mul_ntt(z.c, a.c, b.c); return z; }
    SR_D static void acc(Val& s, const Val& x) {
#pragma unroll
        for (int i = 0; i < 9; i++) s.c[i] = bb::add(s.c[i], x.c[i]);
    }
    // the second 2^-32 of the Montgomery-64 product is applied once per accumulated sum (bb_ring.cuh)
    SR_D static Val mul_lazy(const Val& a, const Val& b) { Val z; bb::slot_mul_ntt_lazy(z.c, a.c, b.c); return z; }
    SR_D static void finish(Val& s) {
#pragma unroll
        for (int i = 0; i < 9; i++) s.c[i] = bb::red((u64)s.c[i]);
    }
    // Accum: the nine 64-bit accumulators of the slot product kept unreduced over the whole sum (bb::SlotAcc)
    typedef bb::SlotAcc Accum;
    typedef bb::SlotPrep Prep;
    SR_D static Prep prep(const Val& x) { Prep p; bb::slot_prep_ntt(p, x.c); return p; }
    SR_D static void accum_zero(Accum& A) { bb::slot_acc_zero(A); }
    SR_D static void accum_mad_p(Accum& A, const Val& a, const Prep& x) { bb::slot_acc_mad(A, a.c, x); }
    SR_D static void accum_mad(Accum& A, const Val& a, const Val& x) { accum_mad_p(A, a, prep(x)); }
    SR_D static Val accum_result(const Accum& A) { Val r; bb::slot_acc_result(r.c, A); return r; }
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
    SR_D static Val load_cv(const u64* p) {
        Val v;
        uint4 lo = __ldcv(reinterpret_cast<const uint4*>(p)), hi = __ldcv(reinterpret_cast<const uint4*>(p) + 1);
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    SR_D static void store(u64* p, const Val& v) {
        reinterpret_cast<uint4*>(p)[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
        reinterpret_cast<uint4*>(p)[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
    SR_D static void store_poison(u64* p) { p[0] = p[1] = p[2] = p[3] = ~0ull; }
    SR_D static Val zero() {
        Val v;
#pragma unroll
        for (int i = 0; i < 8; i++) v.v[i] = 0;
        return v;
    }
    SR_D static bool is_zero(const Val& v) {
        u32 o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v.v[i];
        return o == 0;
    }
    SR_D static Val mul(const Val& a, const Val& b) { Val z; sp::mont_mul(z, a, b); return z; }
    SR_D static void acc(Val& s, const Val& x) { Val t; sp::add(t, s, x); s = t; }
    SR_D static Val mul_lazy(const Val& a, const Val& b) { return mul(a, b); }
    SR_D static void finish(Val&) {}
    // Accum: the running sum kept unreduced (sp::DotAcc: products without their final subtraction, plain additions)
    typedef sp::DotAcc Accum;
    typedef Val Prep;
    SR_D static Prep prep(const Val& x) { return x; }
    SR_D static void accum_zero(Accum& A) { sp::dot_zero(A); }
    SR_D static void accum_mad(Accum& A, const Val& a, const Val& x) { sp::dot_mad(A, a, x); }
    SR_D static void accum_mad_p(Accum& A, const Val& a, const Prep& x) { sp::dot_mad(A, a, x); }
    SR_D static Val accum_result(const Accum& A) { Val r; sp::dot_result(r, A); return r; }
};

}  // namespace sr
