// Goldilocks ring F_p[X]/(X^24 - X^12 + 1), p = 2^64 - 2^32 + 1: per-element transforms on a
// register-resident u64[24].
//
// Mirrors (reference, crates/ring/src/cyclotomic_ring/models/goldilocks/):
//   ntt.rs:135-228  serial_goldilock_crt_in_place      -> gl::crt
//   ntt.rs:240-319  serial_goldilock_icrt_in_place     -> gl::icrt
//   ntt.rs:326-437  homogenize_fq3 / dehomogenize_fq3  -> SR_GL_HOMOG / SR_GL_DEHOMOG
//   mod.rs:34-52 (Fq3 = Fq[u]/(u^3 - 2^40)) + ark-ff CubicExtField -> gl::slot_mul
//
// Every root of unity the reference tabulates is a power of two (ROOTS_OF_UNITY_24[k] =
// 2^(40 k mod 192), 1/8 = 2^189, 1/4 = 2^190; 2 has order 192 and 2^96 = -1), so all twiddle
// multiplications are shift-and-reduce with the special form of p; only KAPPA and the slot
// products need a 64 x 64 multiplication.  Memory holds x * 2^64 mod p (Montgomery); CRT/ICRT are
// linear so they act on the raw limbs directly, and the slot product's extra 2^-64 = 2^128 is a
// power of two as well (folded into the 1/8, 1/4 scalings in the fused path).
#pragma once
#include "sr_common.cuh"
#include "sr_consts_gen.cuh"
#include "sr_homog_gen.cuh"

namespace sr {
namespace gl {

constexpr u64 P = 0xFFFFFFFF00000001ull;
constexpr u64 EPS = 0xFFFFFFFFull;  // 2^64 mod p
constexpr int D = 24;

constexpr int root_exp(int k) {
    constexpr int t[24] = SR_GL_ROOT_EXPS;
    return t[k];
}

// canonical add / sub / neg (inputs < p)
SR_HD u64 add(u64 a, u64 b) {
    u64 s = a + b;
    bool over = (s < a) | (s >= P);
    return over ? s + EPS : s;  // s - p == s + EPS (mod 2^64)
}
SR_HD u64 sub(u64 a, u64 b) {
    u64 d = a - b;
    return (a < b) ? d - EPS : d;  // d + p == d - EPS (mod 2^64)
}
SR_HD u64 neg(u64 a) { return a ? P - a : 0; }

// (hi, lo) = 128-bit value -> canonical residue.  2^64 = 2^32 - 1, 2^96 = -1 (mod p).
SR_HD u64 reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= EPS;          // borrowed 2^64 = EPS
    u64 t1 = (hl << 32) - hl;        // hl * (2^32 - 1)
    u64 r = t0 + t1;
    if (r < t1) r += EPS;            // carried 2^64 = EPS
    return r >= P ? r - P : r;
}
SR_HD u64 mul(u64 a, u64 b) { return reduce128(a * b, mul64hi(a, b)); }

// x * 2^K mod p for a compile-time K in [0, 192)
template <int K>
SR_HD u64 mul_pow2(u64 x) {
    static_assert(K >= 0 && K < 192, "exponent");
    constexpr int KK = K % 96, s = KK % 32, j = KK / 32;
    // v = x << s as 96 bits (v2:v1:v0)
    u64 lo = x << s;
    u32 v2 = s ? (u32)(x >> (64 - s)) : 0u;
    u32 v0 = (u32)lo, v1 = (u32)(lo >> 32);
    u64 r;
    if (j == 0) {
        r = reduce128(lo, v2);
    } else if (j == 1) {  // v0 2^32 + v1 2^64 - v2
        r = sub(reduce128((u64)v0 << 32, v1), (u64)v2);
    } else {              // v0 2^64 - v1 - v2 2^32
        r = sub(reduce128(0, v0), (u64)v1 | ((u64)v2 << 32));
    }
    return K >= 96 ? neg(r) : r;
}
template <int K>
SR_HD u64 mulw(u64 x) {  // x * ROOTS_OF_UNITY_24[K]
    return mul_pow2<root_exp(K)>(x);
}

template <int LO, int SPAN, int K>
SR_HD void bfly(u64 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        u64 a = c[LO + i], t = mulw<K>(c[LO + SPAN + i]);
        c[LO + i] = add(a, t);
        c[LO + SPAN + i] = sub(a, t);
    }
}
template <int LO, int SPAN, int K>
SR_HD void ibfly(u64 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        u64 a = c[LO + i], b = c[LO + SPAN + i];
        c[LO + i] = add(a, b);
        c[LO + SPAN + i] = mulw<K>(sub(a, b));
    }
}

// ntt.rs:146-225
SR_HD void crt_stages(u64 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 a = c[i], b = c[12 + i];
        u64 z = mulw<4>(b);
        c[i] = add(a, z);
        c[12 + i] = sub(add(a, b), z);
    }
    bfly<0, 6, 2>(c);
    bfly<12, 6, 10>(c);
    bfly<0, 3, 1>(c);
    bfly<6, 3, 7>(c);
    bfly<12, 3, 5>(c);
    bfly<18, 3, 11>(c);
}

// ntt.rs:250-318.  EXTRA: additional power of two folded into the 1/8 and 1/4 scalings.
template <int EXTRA>
SR_HD void icrt_stages(u64 (&c)[D]) {
    ibfly<0, 3, 23>(c);
    ibfly<6, 3, 17>(c);
    ibfly<12, 3, 19>(c);
    ibfly<18, 3, 13>(c);
    ibfly<0, 6, 22>(c);
    ibfly<12, 6, 14>(c);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 a = c[i], b = c[12 + i];
        u64 kd = mul(sub(a, b), (u64)SR_GL_KAPPA);
        c[i] = mul_pow2<(189 + EXTRA) % 192>(sub(add(a, b), kd));
        c[12 + i] = mul_pow2<(190 + EXTRA) % 192>(kd);
    }
}

SR_HD void homogenize(u64 (&o)[D], const u64 (&c)[D]) {
#define MULW(k, x) ::sr::gl::mulw<k>(x)
#define NEG ::sr::gl::neg
    SR_GL_HOMOG(o, c)
#undef MULW
#undef NEG
}
SR_HD void dehomogenize(u64 (&o)[D], const u64 (&c)[D]) {
#define MULW(k, x) ::sr::gl::mulw<k>(x)
#define NEG ::sr::gl::neg
    SR_GL_DEHOMOG(o, c)
#undef MULW
#undef NEG
}

SR_HD void crt(u64 (&c)[D]) {
    crt_stages(c);
    u64 o[D];
    homogenize(o, c);
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = o[i];
}
SR_HD void icrt(u64 (&c)[D]) {
    u64 o[D];
    dehomogenize(o, c);
    icrt_stages<0>(o);
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = o[i];
}

// z = x * y in F_p[u]/(u^3 - 2^RHO_EXP), then times 2^POST_EXP.
template <int RHO_EXP, int POST_EXP>
SR_HD void slot_mul(u64* z, const u64* x, const u64* y) {
    u64 x0 = x[0], x1 = x[1], x2 = x[2], y0 = y[0], y1 = y[1], y2 = y[2];
    u64 c0 = add(mul(x0, y0), mul_pow2<RHO_EXP>(add(mul(x1, y2), mul(x2, y1))));
    u64 c1 = add(add(mul(x0, y1), mul(x1, y0)), mul_pow2<RHO_EXP>(mul(x2, y2)));
    u64 c2 = add(add(mul(x0, y2), mul(x1, y1)), mul(x2, y0));
    if (POST_EXP != 0) {
        c0 = mul_pow2<POST_EXP>(c0);
        c1 = mul_pow2<POST_EXP>(c1);
        c2 = mul_pow2<POST_EXP>(c2);
    }
    z[0] = c0; z[1] = c1; z[2] = c2;
}

// ntt_form.rs:159-175 on raw Montgomery limbs: a <- a * b * 2^-64 slot-wise, 2^-64 = 2^128
SR_HD void ntt_mul(u64 (&a)[D], const u64 (&b)[D]) {
#pragma unroll
    for (int s = 0; s < 8; s++) slot_mul<root_exp(1), 128>(&a[3 * s], &a[3 * s], &b[3 * s]);
}

// Fused unit of the metric without the slot isomorphisms: slot s multiplied directly modulo
// X^3 - r^k_s; the Montgomery 2^-64 rides on the final scalings (EXTRA = 128).
template <int S>
SR_HD void fused_slot(u64 (&bs)[D], const u64* as) {
    constexpr int KS[8] = {1, 13, 7, 19, 5, 17, 11, 23};
    slot_mul<root_exp(KS[S]), 0>(&bs[3 * S], &as[3 * S], &bs[3 * S]);
}
SR_HD void fused_mul_icrt(u64 (&bs)[D], const u64* as) {
    fused_slot<0>(bs, as);
    fused_slot<1>(bs, as);
    fused_slot<2>(bs, as);
    fused_slot<3>(bs, as);
    fused_slot<4>(bs, as);
    fused_slot<5>(bs, as);
    fused_slot<6>(bs, as);
    fused_slot<7>(bs, as);
    icrt_stages<128>(bs);
}

}  // namespace gl
}  // namespace sr
