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

// z <- x * y modulo X^6 - rho_q (q warp-uniform); output CANONICAL.
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
#if defined(__CUDA_ARCH__)
        Acc d;
        acc_zero(d);
#pragma unroll
        for (int i = 0; i < 6; i++) acc_mad(d, x[i], i <= k ? y[k - i] : ry[k + 6 - i]);
        z[k] = acc_reduce(d);
#else
        u64 acc = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) acc = add(acc, canon(mul(x[i], i <= k ? y[k - i] : ry[k + 6 - i])));
        z[k] = canon(acc);
#endif
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

// Last inverse stage (ntt.rs:292-318) on the coefficient pairs (i, i + 12), i in [2 t0, 2 t1), two pairs per trip,
// with 2^EXTRA folded into the scalings; output canonical.
template <int EXTRA>
SR_HD void final_pairs(u64* row, int t0, int t1) {
    SR_GL_ROLL
    for (int t = t0; t < t1; t++) {
        u64 a[2], b[2], lo[2], hi[2];
        ld2(row + 2 * t, a[0], a[1]);
        ld2(row + 12 + 2 * t, b[0], b[1]);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const u64 cb = canon(b[u]);
            const u64 kd = canon(mul(sub(a[u], cb), (u64)SR_GL_KAPPA));
            // 1/8 and 1/4 of the three-stage inverse become 1/4 = 2^190 and 1/2 = 2^191: one halving stage fewer
            lo[u] = canon(mul_pow2<(190 + EXTRA) % 192>(sub(add(a[u], cb), kd)));
            hi[u] = canon(mul_pow2<(191 + EXTRA) % 192>(kd));
        }
        st2(row + 2 * t, lo[0], lo[1]);
        st2(row + 12 + 2 * t, hi[0], hi[1]);
    }
}

// Inverse of crt_stages12 with 2^EXTRA folded into the final scalings, in place on the row.  Input CANONICAL
// (the sextic products), output canonical.
template <int EXTRA>
SR_HD void icrt_stages12(u64* row) {
    {
        u64 c[D];
        row_load(c, row);
        ibfly<0, 6, 22, true>(c);   // ntt.rs:272-290
        ibfly<12, 6, 14, true>(c);
        row_store(row, c);
    }
    final_pairs<EXTRA>(row, 0, 6);
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
