// TEST INFRASTRUCTURE: compiles the kernels' __host__ __device__ per-element arithmetic for the
// CPU so that it can be checked against the oracle without a GPU (tests/test_hostcheck.py).
// Not part of the product; libstarkrings_cuda.so has no CPU path.
#include <cstdint>
#include <cstring>

#include "bb_half.cuh"
#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "gl_fused6.cuh"
#include "sp_quad.cuh"
#include "sp_ring.cuh"

using namespace sr;

extern "C" {

void hc_bb_crt(uint64_t* e) {
    u32 c[72];
    for (int i = 0; i < 72; i++) c[i] = (u32)e[i];
    bb::crt(c);
    for (int i = 0; i < 72; i++) e[i] = c[i];
}
void hc_bb_icrt(uint64_t* e) {
    u32 c[72];
    for (int i = 0; i < 72; i++) c[i] = (u32)e[i];
    bb::icrt(c);
    for (int i = 0; i < 72; i++) e[i] = c[i];
}
void hc_bb_ntt_mul(uint64_t* a, const uint64_t* b) {
    u32 x[72], y[72];
    for (int i = 0; i < 72; i++) { x[i] = (u32)a[i]; y[i] = (u32)b[i]; }
    bb::ntt_mul(x, y);
    for (int i = 0; i < 72; i++) a[i] = x[i];
}
void hc_bb_ring_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    u32 x[72], y[72];
    for (int i = 0; i < 72; i++) { x[i] = (u32)a[i]; y[i] = (u32)b[i]; }
    bb::crt_stages(x);
    bb::crt_stages(y);
    bb::fused_mul_icrt(y, x);
    for (int i = 0; i < 72; i++) out[i] = y[i];
}


// two-threads-per-element formulation, both halves run in sequence
void hc_bb_ring_mul_half(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    u32 ra[72], rb[72];
    for (int i = 0; i < 72; i++) { ra[i] = (u32)a[i]; rb[i] = (u32)b[i]; }
    u32 B[2][36], O[2][36];
    for (int h = 0; h < 2; h++) {
        const bb::HalfConsts K = bb::half_consts(h);
        u32 A[36];
        bb::half_crt(A, ra, K);
        bb::half_crt(B[h], rb, K);
        bb::half_slots(B[h], A, K);
        bb::half_icrt_local(B[h], K);
    }
    for (int h = 0; h < 2; h++) {
        const bb::HalfConsts K = bb::half_consts(h);
        bb::half_final(O[h], B[h], B[1 - h], K);
        for (int i = 0; i < 36; i++) out[36 * h + i] = O[h][i];
    }
}

// lazily accumulated sum of NTT-form slot products (the BabyBear mat-vec family): out = sum_c a[c] * x[c], n elements each
void hc_bb_dot(const uint64_t* a, const uint64_t* x, size_t n, uint64_t* out) {
    for (int s = 0; s < 8; s++) {
        bb::SlotAcc A;
        bb::slot_acc_zero(A);
        for (size_t c = 0; c < n; c++) {
            u32 as[9], xs[9];
            for (int i = 0; i < 9; i++) { as[i] = (u32)a[72 * c + 9 * s + i]; xs[i] = (u32)x[72 * c + 9 * s + i]; }
            bb::SlotPrep p;
            bb::slot_prep_ntt(p, xs);
            bb::slot_acc_mad(A, as, p);
            for (int k = 0; k < 9; k++)
                if (A.d[k] >= ((u64)bb::P << 32)) __builtin_trap();  // the invariant of the unreduced sums
        }
        u32 z[9];
        bb::slot_acc_result(z, A);
        for (int i = 0; i < 9; i++) out[9 * s + i] = z[i];
    }
}
// unreduced sum of Starknet-prime products (sp::DotAcc, the mat-vec family): out = sum_c a[c] * x[c] slot-wise
void hc_sp_dot(const uint64_t* a, const uint64_t* x, size_t n, uint64_t* out) {
    for (int s = 0; s < 16; s++) {
        sp::DotAcc A;
        sp::dot_zero(A);
        for (size_t c = 0; c < n; c++) {
            sp::Fe fa, fx;
            memcpy(&fa, a + 64 * c + 4 * s, 32);
            memcpy(&fx, x + 64 * c + 4 * s, 32);
            sp::dot_mad(A, fa, fx);
        }
        sp::Fe r;
        sp::dot_result(r, A);
        memcpy(out + 4 * s, &r, 32);
    }
}
// the two reductions of a 160-bit carry-save accumulator given as raw limbs (e0..e4, o1..o3)
void hc_gl_acc_reduce(const uint32_t* limbs, uint64_t* out) {
    gl::Acc A;
    A.e0 = limbs[0]; A.e1 = limbs[1]; A.e2 = limbs[2]; A.e3 = limbs[3]; A.e4 = limbs[4];
    A.o1 = limbs[5]; A.o2 = limbs[6]; A.o3 = limbs[7];
    out[0] = gl::acc_reduce(A);
    out[1] = gl::acc_reduce_m128(A);
    out[2] = gl::acc_reduce<false>(A);
}
void hc_gl_crt(uint64_t* e) { u64 c[24]; memcpy(c, e, 192); gl::crt(c); memcpy(e, c, 192); }
void hc_gl_icrt(uint64_t* e) { u64 c[24]; memcpy(c, e, 192); gl::icrt(c); memcpy(e, c, 192); }
void hc_gl_ntt_mul(uint64_t* a, const uint64_t* b) {
    u64 x[24], y[24]; memcpy(x, a, 192); memcpy(y, b, 192);
    gl::ntt_mul(x, y); memcpy(a, x, 192);
}
void hc_gl_ring_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    u64 x[24], y[24]; memcpy(x, a, 192); memcpy(y, b, 192);
    gl::crt_stages(x); gl::crt_stages(y); gl::fused_mul_icrt(y, x); memcpy(out, y, 192);
}

// degree-6 formulation (gl_fused6.cuh) used by the fused kernel
void hc_gl_ring_mul_fused6(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    u64 rows[48]; memcpy(rows, a, 192); memcpy(rows + 24, b, 192);
    gl::ring_mul_fused6(rows, rows + 24, 2); memcpy(out, rows, 192);
}
void hc_gl_ntt_mul_rolled(uint64_t* a, const uint64_t* b) { gl::ntt_mul_rolled(a, b); }
void hc_gl_icrt_row(uint64_t* e) { gl::icrt_row(e); }

void hc_sp_crt(uint64_t* e) { sp::Fe c[16]; memcpy(c, e, 512); sp::crt(c); memcpy(e, c, 512); }
void hc_sp_icrt(uint64_t* e) { sp::Fe c[16]; memcpy(c, e, 512); sp::icrt(c); memcpy(e, c, 512); }
void hc_sp_ntt_mul(uint64_t* a, const uint64_t* b) {
    sp::Fe x[16], y[16]; memcpy(x, a, 512); memcpy(y, b, 512);
    for (int i = 0; i < 16; i++) sp::mont_mul(x[i], x[i], y[i]);
    memcpy(a, x, 512);
}
void hc_sp_ring_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    sp::Fe x[16], y[16]; memcpy(x, a, 512); memcpy(y, b, 512);
    sp::crt(x); sp::crt(y);
    for (int i = 0; i < 16; i++) sp::mont_mul(y[i], y[i], x[i]);
    sp::icrt(y); memcpy(out, y, 512);
}


// four-threads-per-element formulation (sp_quad.cuh): the four "threads" in turn, stage by stage
static void sp_quad_fwd(uint32_t* row) {
    const uint32_t* wtab = &sp::ROOTS_MONT.w[0][0];
    for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) sp::quad_fwd_stage(row, wtab, s, t);
}
static void sp_quad_inv(uint32_t* row) {
    const uint32_t* wtab = &sp::ROOTS_MONT.w[0][0];
    for (int s = 0; s < 3; s++) for (int t = 0; t < 4; t++) sp::quad_inv_stage(row, wtab, s, t);
    for (int t = 0; t < 4; t++) sp::quad_inv_last(row, t);
}
// the unreduced ("lazy") schedule of the fused ring product: every bound is checked at run time (__builtin_trap)
void hc_sp_ring_mul_quad_lazy(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    const uint32_t* wtab = &sp::ROOTS_MONT.w[0][0];
    uint32_t ra[128], rb[128]; memcpy(ra, a, 512); memcpy(rb, b, 512);
    for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) sp::quad_fwd_stage<true>(ra, wtab, s, t);
    for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) sp::quad_fwd_stage<true>(rb, wtab, s, t);
    for (int t = 0; t < 4; t++) sp::quad_slots<true>(ra, rb, t);
    for (int s = 0; s < 3; s++) for (int t = 0; t < 4; t++) sp::quad_inv_stage<true>(ra, wtab, s, t);
    for (int t = 0; t < 4; t++) sp::quad_inv_last<true>(ra, t);
    memcpy(out, ra, 512);
}
// the unreduced schedules of the stand-alone CRT / ICRT kernels (bounds trap on the host)
void hc_sp_crt_quad_lazy(uint64_t* e) {
    const uint32_t* wtab = &sp::ROOTS_MONT.w[0][0];
    uint32_t* row = (uint32_t*)e;
    for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) sp::quad_fwd_stage<true, true>(row, wtab, s, t);
}
void hc_sp_icrt_quad_lazy(uint64_t* e) {
    const uint32_t* wtab = &sp::ROOTS_MONT.w[0][0];
    uint32_t* row = (uint32_t*)e;
    for (int s = 0; s < 3; s++) for (int t = 0; t < 4; t++) sp::quad_inv_stage<true>(row, wtab, s, t);
    for (int t = 0; t < 4; t++) sp::quad_inv_last<true, true>(row, t);
}
void hc_sp_crt_quad(uint64_t* e) { sp_quad_fwd((uint32_t*)e); }
void hc_sp_icrt_quad(uint64_t* e) { sp_quad_inv((uint32_t*)e); }
void hc_sp_ring_mul_quad(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    uint32_t ra[128], rb[128]; memcpy(ra, a, 512); memcpy(rb, b, 512);
    sp_quad_fwd(ra); sp_quad_fwd(rb);
    for (int t = 0; t < 4; t++) sp::quad_slots(ra, rb, t);
    sp_quad_inv(ra); memcpy(out, ra, 512);
}

}  // extern "C"
