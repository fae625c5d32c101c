// BabyBear ring F_p[X]/(X^72 - X^36 + 1), p = 15 * 2^27 + 1: per-element transforms on a
// register-resident u32[72].
//
// Mirrors (reference, crates/ring/src/cyclotomic_ring/models/babybear/):
//   ntt.rs:143-236  serial_babybear_crt_in_place      -> bb::crt
//   ntt.rs:238-317  serial_babybear_icrt_in_place     -> bb::icrt
//   ntt.rs:324-588  homogenize_fq9 / dehomogenize_fq9 -> SR_BB_HOMOG / SR_BB_DEHOMOG
//   fq9.rs:19-58 + ark-ff CubicExtField               -> bb::slot_mul (Y^9 = r product)
//
// Data in memory is ark-ff Fp64<MontBackend<_,1>>: one u64 limb holding x * 2^64 mod p (< 2^31).
// Only the low 32-bit word is ever non-zero.  CRT/ICRT are linear with standard-form constants,
// so they act on the raw words directly; the constants are applied with a 32-bit Montgomery
// multiplication (operand w * 2^32 mod p).  The slot product of two raw words needs one factor
// 2^-64 = (2^-32)^2: one comes from the Montgomery reduction of the accumulated products, the
// other from a second reduction (ntt_mul) or from the final ICRT scaling constants (fused ring_mul).
#pragma once
#include "sr_common.cuh"
#include "sr_consts_gen.cuh"
#include "sr_homog_gen.cuh"

namespace sr {
namespace bb {

constexpr u32 P = 0x78000001u;      // 2013265921
constexpr u32 NPINV = 0x77FFFFFFu;  // -p^-1 mod 2^32
constexpr int D = 72;               // coefficients per ring element
constexpr int SLOT = 9;             // slot width

// ROOTS_OF_UNITY_24[k] * 2^32 mod p (operand form) and standard form, as compile-time values
constexpr u32 w_m32(int k) {
    constexpr u32 t[24] = SR_BB_ROOTS_M32;
    return t[k];
}
constexpr u32 w_std(int k) {
    constexpr u32 t[24] = SR_BB_ROOTS_STD;
    return t[k];
}

constexpr u32 cmulmod(u32 a, u32 b) { return (u32)(((u64)a * b) % P); }
constexpr u32 R32 = (u32)((1ull << 32) % P);             // 2^32 mod p
constexpr u32 R32_INV = 943718400u;                      // 2^-32 mod p
static_assert(cmulmod(R32, R32_INV) == 1, "2^-32");
constexpr u32 to_m32(u32 x) { return cmulmod(x, R32); }  // operand form for mulc

// canonical ([0,p)) add / sub / neg
SR_HD u32 add(u32 a, u32 b) { u32 s = a + b; return umin32(s, s - P); }
SR_HD u32 sub(u32 a, u32 b) { u32 d = a - b; return umin32(d, d + P); }
SR_HD u32 neg(u32 a) { return sub(0u, a); }

// Montgomery reduction t -> t * 2^-32 mod p.  t < p * 2^32  =>  result in [0, 2p).
SR_HD u32 red_lazy(u64 t) {
    u32 m = (u32)t * NPINV;
    return (u32)(((u64)m * P + t) >> 32);
}
SR_HD u32 red(u64 t) { u32 u = red_lazy(t); return umin32(u, u - P); }
// a * w for a constant given in operand form wm = w * 2^32 mod p; a < 2p allowed.
SR_HD u32 mulc(u32 a, u32 wm) { return red((u64)a * wm); }

// Shoup multiplication by a constant w < p with its companion w' = floor(w 2^32 / p): q = hi(a w'), r = a w - q p lies in
// [0, 2p) for ANY 32-bit a, so the 32-bit wrap-around arithmetic is exact.  One IMAD.HI and two 32-bit IMAD against the
// wide multiply, the 32-bit multiply and the IMAD.HI of a Montgomery multiplication: measured 4.56 against 3.42 T
// multiplications/s (tools/int_pipe_peak.cu), on the pipe that bounds the fused ring product.
struct Tw {
    u32 w, wp;
};
constexpr Tw shoup(u32 w_std) { return Tw{w_std, (u32)(((u64)w_std << 32) / P)}; }
SR_HD u32 muls(u32 a, const Tw& t) {
    const u32 q = mulhi32(a, t.wp);
    const u32 r = a * t.w - q * P;
    return umin32(r, r - P);
}

template <int K>
SR_HD u32 mulw(u32 a) {  // a * ROOTS_OF_UNITY_24[K], a < 2^32, result canonical
    constexpr Tw t = shoup(w_std(K));
    return muls(a, t);
}
#define SR_BB_MULW(k, x) ::sr::bb::mulw<k>(x)

// (a, b) <- (a + w b, a - w b)
template <int LO, int SPAN, int K>
SR_HD void bfly(u32 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        u32 a = c[LO + i], t = mulw<K>(c[LO + SPAN + i]);
        c[LO + i] = add(a, t);
        c[LO + SPAN + i] = sub(a, t);
    }
}
// (a, b) <- (a + b, w (a - b))
template <int LO, int SPAN, int K>
SR_HD void ibfly(u32 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        u32 a = c[LO + i], b = c[LO + SPAN + i];
        c[LO + i] = add(a, b);
        // a - b + p in (0, 2p): valid lazy operand for the constant multiplication
        c[LO + SPAN + i] = mulw<K>(a - b + P);
    }
}

// Butterfly stages of the CRT (ntt.rs:154-233): slot s then holds f mod X^9 - r^k_s.
SR_HD void crt_stages(u32 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < 36; i++) {
        u32 a = c[i], b = c[36 + i];
        u32 z = mulw<4>(b);
        c[i] = add(a, z);
        c[36 + i] = sub(add(a, b), z);
    }
    bfly<0, 18, 2>(c);
    bfly<36, 18, 10>(c);
    bfly<0, 9, 1>(c);
    bfly<18, 9, 7>(c);
    bfly<36, 9, 5>(c);
    bfly<54, 9, 11>(c);
}

// Inverse stages (ntt.rs:249-316).  SCALE_M32 is an extra constant (operand form) folded into
// the final 1/8 and 1/4 scalings; to_m32(1) for the plain ICRT.
template <u32 SCALE_STD>
SR_HD void icrt_stages(u32 (&c)[D]) {
    ibfly<0, 9, 23>(c);
    ibfly<18, 9, 17>(c);
    ibfly<36, 9, 19>(c);
    ibfly<54, 9, 13>(c);
    ibfly<0, 18, 22>(c);
    ibfly<36, 18, 14>(c);
    constexpr Tw KAPPA_T = shoup(SR_BB_KAPPA_STD);
    constexpr Tw E8_T = shoup(cmulmod(SR_BB_EIGHT_INV_STD, SCALE_STD));
    constexpr Tw E4_T = shoup(cmulmod(SR_BB_FOUR_INV_STD, SCALE_STD));
#pragma unroll
    for (int i = 0; i < 36; i++) {
        u32 a = c[i], b = c[36 + i];
        u32 kd = muls(a - b + P, KAPPA_T);
        c[i] = muls(sub(add(a, b), kd), E8_T);
        c[36 + i] = muls(kd, E4_T);
    }
}

SR_HD void homogenize(u32 (&o)[D], const u32 (&c)[D]) {
#define MULW SR_BB_MULW
#define NEG ::sr::bb::neg
    SR_BB_HOMOG(o, c)
#undef MULW
#undef NEG
}
SR_HD void dehomogenize(u32 (&o)[D], const u32 (&c)[D]) {
#define MULW SR_BB_MULW
#define NEG ::sr::bb::neg
    SR_BB_DEHOMOG(o, c)
#undef MULW
#undef NEG
}

SR_HD void crt(u32 (&c)[D]) {
    crt_stages(c);
    u32 o[D];
    homogenize(o, c);
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = o[i];
}
SR_HD void icrt(u32 (&c)[D]) {
    u32 o[D];
    dehomogenize(o, c);
    icrt_stages<1u>(o);
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = o[i];
}

// Product of x, y (power order: index j = coefficient of Y^j) in F_p[Y]/(Y^9 - rho), times 2^-32:
//   z_k = 2^-32 * ( sum_{i+j=k} x_i y_j + rho * sum_{i+j=k+9} x_i y_j ).
// RHO_STD = rho (standard form).  Inputs canonical, output canonical.
// 64-bit accumulation: every product < p^2 < B/2.13 with B = p 2^32; the running sum is kept below
// 2^64 by folding the high word with min(hi, hi - p) (a subtraction of B when hi >= p).
template <u32 RHO_STD>
SR_HD void slot_mul_pow(u32 (&z)[SLOT], const u32 (&x)[SLOT], const u32 (&y)[SLOT]) {
    constexpr Tw rho = shoup(RHO_STD);
    u32 yr[SLOT];  // rho * y_j
#pragma unroll
    for (int j = 1; j < SLOT; j++) yr[j] = muls(y[j], rho);
#pragma unroll
    for (int k = 0; k < SLOT; k++) {
        u64 acc = 0;
        int n = 0;
#pragma unroll
        for (int i = 0; i < SLOT; i++) {
            // term i: x_i * y_{k-i} if i <= k else x_i * yr_{k+9-i}
            u32 f = (i <= k) ? y[(i <= k) ? k - i : 0] : yr[(i <= k) ? 1 : k + SLOT - i];
            acc += (u64)x[i] * f;
            n++;
            if (n == 4 || n == 6 || n == 8 || n == 9) {  // keep acc < B before the next additions
                u32 hi = (u32)(acc >> 32);
                hi = umin32(hi, hi - P);
                acc = ((u64)hi << 32) | (u32)acc;
            }
        }
        z[k] = red(acc);
    }
}

// NTT-form slot product (memory order of Fq9 = 3 x Fq3: index 3 (j mod 3) + j / 3 holds Y^j),
// result scaled by 2^-32 * EXTRA where EXTRA_M32 is in operand form (R32 => plain 2^-32...).
// For the Montgomery-64 layout a full product needs 2^-64: slot_mul_ntt applies a second reduction.
SR_HD void slot_mul_ntt(u32* z, const u32* x, const u32* y) {
    u32 xp[SLOT], yp[SLOT], zp[SLOT];
#pragma unroll
    for (int j = 0; j < SLOT; j++) {
        xp[j] = x[3 * (j % 3) + j / 3];
        yp[j] = y[3 * (j % 3) + j / 3];
    }
    slot_mul_pow<w_std(1)>(zp, xp, yp);
#pragma unroll
    for (int j = 0; j < SLOT; j++) z[3 * (j % 3) + j / 3] = red((u64)zp[j]);
}

// The same product with ONE of the two 2^-32 factors still missing (result = x y 2^-32 in memory order): sums of such
// products are finished with a single extra reduction per coefficient (the mat-vec family: nine wide multiply-adds
// less per product).
SR_HD void slot_mul_ntt_lazy(u32* z, const u32* x, const u32* y) {
    u32 xp[SLOT], yp[SLOT], zp[SLOT];
#pragma unroll
    for (int j = 0; j < SLOT; j++) {
        xp[j] = x[3 * (j % 3) + j / 3];
        yp[j] = y[3 * (j % 3) + j / 3];
    }
    slot_mul_pow<w_std(1)>(zp, xp, yp);
#pragma unroll
    for (int j = 0; j < SLOT; j++) z[3 * (j % 3) + j / 3] = zp[j];
}

// ---- sums of NTT-form slot products (the mat-vec family) -----------------------------------------------------------
// sum_c a_c * x_c for one slot, with the 64-bit accumulators of slot_mul_pow kept UNREDUCED across the whole sum: the
// running value of every output coefficient stays below B = p 2^32 (fold of the high word after every second
// product: B + 2 p^2 < 2 B), one Montgomery reduction pair per coefficient at the very end instead of one reduction
// and one modular addition per coefficient and product.  The vector operand is prepared once per slot (power order,
// multiples by rho) and shared by all matrix rows.
struct SlotPrep {
    u32 y[SLOT], yr[SLOT];  // power order; yr[j] = rho y[j] (j >= 1)
};
struct SlotAcc {
    u64 d[SLOT];  // power order; invariant: d[k] < p 2^32
};
SR_HD void slot_prep_ntt(SlotPrep& p, const u32* x) {  // x: one slot in memory order
    constexpr Tw rho = shoup(w_std(1));
#pragma unroll
    for (int j = 0; j < SLOT; j++) p.y[j] = x[3 * (j % 3) + j / 3];
    p.yr[0] = 0;
#pragma unroll
    for (int j = 1; j < SLOT; j++) p.yr[j] = muls(p.y[j], rho);
}
SR_HD void slot_acc_zero(SlotAcc& A) {
#pragma unroll
    for (int k = 0; k < SLOT; k++) A.d[k] = 0;
}
SR_HD void slot_acc_mad(SlotAcc& A, const u32* a, const SlotPrep& p) {  // a: one slot in memory order
    u32 x[SLOT];
#pragma unroll
    for (int j = 0; j < SLOT; j++) x[j] = a[3 * (j % 3) + j / 3];
#pragma unroll
    for (int k = 0; k < SLOT; k++) {
        u64 acc = A.d[k];
#pragma unroll
        for (int i = 0; i < SLOT; i++) {
            const u32 f = (i <= k) ? p.y[(i <= k) ? k - i : 0] : p.yr[(i <= k) ? 1 : k + SLOT - i];
            acc += (u64)x[i] * f;
            if ((i & 1) || i == SLOT - 1) {  // after every second product, and after the last: back below p 2^32
                u32 hi = (u32)(acc >> 32);
                hi = umin32(hi, hi - P);
                acc = ((u64)hi << 32) | (u32)acc;
            }
        }
        A.d[k] = acc;
    }
}
SR_HD void slot_acc_result(u32* z, const SlotAcc& A) {  // z: memory order; both 2^-32 of the Montgomery-64 layout
#pragma unroll
    for (int j = 0; j < SLOT; j++) z[3 * (j % 3) + j / 3] = red((u64)red(A.d[j]));
}

// ntt_form.rs:159-175 on the memory layout: a <- a * b slot-wise (raw Montgomery-64 words).
SR_HD void ntt_mul(u32 (&a)[D], const u32 (&b)[D]) {
#pragma unroll
    for (int s = 0; s < 8; s++) slot_mul_ntt(&a[SLOT * s], &a[SLOT * s], &b[SLOT * s]);
}

// Fused unit of the metric, icrt(crt(a) * crt(b)), without the slot isomorphisms: the slot-s
// product is taken directly in F_p[X]/(X^9 - r^k_s) (k_s = 1,13,7,19,5,17,11,23), which the
// homogenize / dehomogenize pair would only conjugate.  `bs` holds crt_stages(b); `as` points at
// crt_stages(a) (registers on the host, the thread's shared-memory row on the device).
template <int S>
SR_HD void fused_slot(u32 (&bs)[D], const u32* as) {
    constexpr int KS[8] = {1, 13, 7, 19, 5, 17, 11, 23};
    u32 x[SLOT], y[SLOT], z[SLOT];
#pragma unroll
    for (int j = 0; j < SLOT; j++) {
        x[j] = as[SLOT * S + j];
        y[j] = bs[SLOT * S + j];
    }
    slot_mul_pow<w_std(KS[S])>(z, x, y);
#pragma unroll
    for (int j = 0; j < SLOT; j++) bs[SLOT * S + j] = z[j];
}
// bs <- ring product (coefficient form); inputs: as = crt_stages(a), bs = crt_stages(b).
SR_HD void fused_mul_icrt(u32 (&bs)[D], const u32* as) {
    fused_slot<0>(bs, as);
    fused_slot<1>(bs, as);
    fused_slot<2>(bs, as);
    fused_slot<3>(bs, as);
    fused_slot<4>(bs, as);
    fused_slot<5>(bs, as);
    fused_slot<6>(bs, as);
    fused_slot<7>(bs, as);
    icrt_stages<R32_INV>(bs);  // the second 2^-32 of the Montgomery-64 product rides on 1/8, 1/4
}

}  // namespace bb
}  // namespace sr
