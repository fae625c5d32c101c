// Goldilocks fused ring multiplication, "degree-6" formulation.
//
// Where the time went (ncu, profiles/r01b_gl_ncu.md): the fully unrolled three-stage formulation
// (gl_ring.cuh: crt_stages / fused_mul_icrt) is 7.5k straight-line instructions = 120 KB against 32 KB of L1.5
// instruction cache (`no_instructions` = 49% of the stall samples), and 70% of its instructions are 64-bit modular
// add / sub / shift sequences on the ALU pipe (64% busy) while the multiply-add pipe idles at 14%.
//
// This formulation moves work from the ALU pipe to the multiply-add pipe and shrinks the code:
//   * only the first two butterfly stages are run (crt_stages12): the four quarters of the element are then
//     f mod X^6 - rho_q, rho_q = r^2, r^14, r^10, r^22.  The products are taken modulo these sextics: 36 lazy
//     64 x 64 multiply-accumulates and 6 reductions per quarter instead of a third butterfly stage on both
//     operands, its inverse, and the 3 x 3 products (-36 butterflies, +72 multiply-accumulates per ring mul);
//   * the forward transform exists once and runs for both operands in a 2-trip loop, the four sextic products
//     and the final inverse stage are real loops over the thread's shared-memory row (128-bit, conflict-free
//     accesses), so the kernel is about 2k instructions.
// Bit-identical to icrt(crt(a) * crt(b)): the slot isomorphisms and the third stage only conjugate the product.
// Reference: goldilocks/ntt.rs:135-319 (stages), coeff_form.rs:54-67 + goldilocks/mod.rs test_mul_crt (identity).
// All functions are __host__ __device__; tests/hostcheck runs them on the CPU against the oracle.
#pragma once
#include "gl_ring.cuh"

namespace sr {
namespace gl {

// quarter q holds f mod X^6 - rho_q with rho_q the squares of the stage-3 roots 1, 7, 5, 11 (ntt.rs:196-225):
// rho = r^2, r^14, r^10, r^22 = 2^80, -2^80, 2^16, -2^16.  The quarter index is uniform across the warp (it is the loop
// trip), so the multiplication by rho_q is a uniform four-way branch over COMPILE-TIME shift-reductions (about a
// third of the instructions of a shift by a runtime exponent: no funnel shifts by a register, no word-rotation
// branches, the negation folded into the last subtraction).
template <int Q>
SR_HD u64 mul_rho(u64 y) {
    constexpr int E[4] = {root_exp(2), root_exp(14), root_exp(10), root_exp(22)};
    return mul_pow2<E[Q]>(y);
}
#if defined(__CUDA_ARCH__)
#define SR_GL_ROLL _Pragma("unroll 1")
#else
#define SR_GL_ROLL
#endif

// two consecutive limbs of a row (16-byte aligned: one 128-bit shared-memory access on the device)
SR_HD void ld2(const u64* p, u64& a, u64& b) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}
SR_HD void st2(u64* p, u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
SR_HD void row_load(u64 (&c)[D], const u64* row) {
#pragma unroll
    for (int i = 0; i < D; i += 2) ld2(row + i, c[i], c[i + 1]);
}
SR_HD void row_store(u64* row, const u64 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < D; i += 2) st2(row + i, c[i], c[i + 1]);
}

// z <- x * y modulo X^6 - rho_q (q warp-uniform); output canonical for odd q, weak for even q: the inverse stages
// (tail_dot2) use quarters 1 and 3 as the canonical operands of their additions and subtractions and accept weak
// values from quarters 0 and 2.
// y_1..y_5 are pre-multiplied by rho (five shift-reductions), after which output k is one lazy sum of six products:
// sum_{i <= k} x_i y_{k-i} + sum_{i > k} x_i (rho y_{k+6-i}).
SR_HD void sextic_mul(u64 (&z)[6], const u64 (&x)[6], const u64 (&y)[6], int q) {
    u64 ry[6];
    ry[0] = 0;
    if (q == 0) {
#pragma unroll
        for (int i = 1; i < 6; i++) ry[i] = mul_rho<0>(y[i]);
    } else if (q == 1) {
#pragma unroll
        for (int i = 1; i < 6; i++) ry[i] = mul_rho<1>(y[i]);
    } else if (q == 2) {
#pragma unroll
        for (int i = 1; i < 6; i++) ry[i] = mul_rho<2>(y[i]);
    } else {
#pragma unroll
        for (int i = 1; i < 6; i++) ry[i] = mul_rho<3>(y[i]);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        Acc d;
        acc_mul(d, x[0], y[k]);
#pragma unroll
        for (int i = 1; i < 6; i++) acc_mad(d, x[i], i <= k ? y[k - i] : ry[k + 6 - i]);
        z[k] = acc_reduce<false>(d);
    }
    if (q & 1) {
#pragma unroll
        for (int k = 0; k < 6; k++) z[k] = canon(z[k]);
    }
}

// rowA[6q .. 6q+6) <- rowA[6q ..] * rowB[6q ..] modulo X^6 - 2^e_q for the four quarters
SR_HD void sextic_products(u64* rowA, const u64* rowB) {
    SR_GL_ROLL
    for (int q = 0; q < 4; q++) {
        u64 x[6], y[6], z[6];
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
            ld2(rowA + 6 * q + i, x[i], x[i + 1]);
            ld2(rowB + 6 * q + i, y[i], y[i + 1]);
        }
        sextic_mul(z, x, y, q);
#pragma unroll
        for (int i = 0; i < 6; i += 2) st2(rowA + 6 * q + i, z[i], z[i + 1]);
    }
}

// The last two inverse stages (ntt.rs:272-318) in place on the row as constant-coefficient sums (gl_ring.cuh, TailK),
// two coefficient quadruples (i, i + 6, i + 12, i + 18) per trip.  CANON_IN: the second and fourth quarter of the row
// hold canonical values (sums and differences first, two products per output); otherwise weak values (the subtrahends are canonicalised first: four
// products per output with no canonicalisation were measured slower, twice the wide multiply-adds).  Output canonical.
template <class K, bool CANON_IN>
SR_HD void tail_rows(u64* row) {
    SR_GL_ROLL
    for (int t = 0; t < 3; t++) {
        u64 u0[2], u1[2], u2[2], u3[2], o0[2], o1[2], o2[2], o3[2];
        ld2(row + 2 * t, u0[0], u0[1]);
        ld2(row + 6 + 2 * t, u1[0], u1[1]);
        ld2(row + 12 + 2 * t, u2[0], u2[1]);
        ld2(row + 18 + 2 * t, u3[0], u3[1]);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (CANON_IN) tail_dot2<K>(o0[u], o1[u], o2[u], o3[u], u0[u], u1[u], u2[u], u3[u]);
            else tail_dot2<K>(o0[u], o1[u], o2[u], o3[u], u0[u], canon(u1[u]), u2[u], canon(u3[u]));
        }
        st2(row + 2 * t, o0[0], o0[1]);
        st2(row + 6 + 2 * t, o1[0], o1[1]);
        st2(row + 12 + 2 * t, o2[0], o2[1]);
        st2(row + 18 + 2 * t, o3[0], o3[1]);
    }
}
// Inverse of crt_stages12 with 2^EXTRA folded into the final scalings: the 1/8 and 1/4 of the three-stage inverse become
// 1/4 = 2^190 and 1/2 = 2^191 (one halving stage fewer).  Input: the sextic products (quarters 1 and 3 canonical),
// output canonical.
template <int EXTRA>
SR_HD void icrt_stages12(u64* row) {
    tail_rows<TailK<(190 + EXTRA) % 192, (191 + EXTRA) % 192>, true>(row);
}
// ICRT (ntt.rs:240-319 after the slot isomorphism ntt.rs:385-437) in place on the row: dehomogenisation and the first
// inverse stage in registers, the last two stages as a loop over the row.  Canonical in, canonical out.
SR_HD void icrt_row(u64* row) {
    {
        u64 c[D], o[D];
        row_load(c, row);
        dehomogenize_c(o, c);
        icrt_stage1(o);
        row_store(row, o);
    }
    tail_rows<TailK<189, 190>, false>(row);
}

// NTT-form product (ntt_form.rs:159-175 on raw Montgomery limbs) as a real loop, two slots (48 bytes of each row)
// per trip: every slot is multiplied modulo u^3 - r with the same r = 2^40, so the body is shared verbatim.
SR_HD void ntt_mul_rolled(u64* rowA, const u64* rowB) {
    SR_GL_ROLL
    for (int t = 0; t < 4; t++) {
        u64 x[6], y[6];
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
            ld2(rowA + 6 * t + i, x[i], x[i + 1]);
            ld2(rowB + 6 * t + i, y[i], y[i + 1]);
        }
        slot_mul<root_exp(1), 128>(x, x, y);
        slot_mul<root_exp(1), 128>(x + 3, x + 3, y + 3);
#pragma unroll
        for (int i = 0; i < 6; i += 2) st2(rowA + 6 * t + i, canon(x[i]), canon(x[i + 1]));
    }
}

// rowA <- a * b in F_p[X]/(X^24 - X^12 + 1), coefficient form; rowB is clobbered.  `trips` must be 2: it is passed
// in (opaquely) so that the compiler keeps ONE copy of the forward transform for both operands.
SR_HD void ring_mul_fused6(u64* rowA, u64* rowB, int trips) {
    const ptrdiff_t delta = rowB - rowA;
    SR_GL_ROLL
    for (int k = 0; k < trips; k++) {
        u64 c[D];
        row_load(c, rowA + k * delta);
        crt_stages12(c);
        row_store(rowA + k * delta, c);
    }
    sextic_products(rowA, rowB);
    icrt_stages12<128>(rowA);  // Montgomery layout: the product of two raw limbs carries 2^-64 = 2^128
}

}  // namespace gl
}  // namespace sr
