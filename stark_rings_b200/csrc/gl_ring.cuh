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
//
// Value discipline.  Intermediate values are WEAK: any u64 congruent to the value mod p.  The
// primitives below keep that cheap on the GPU (PTX carry chains, no compare/select):
//   add(a, b), sub(a, b)   a weak, b CANONICAL (< p, or == p)  -> weak      5 instructions
//   canon(x)               weak -> canonical
//   mul, mul_pow2<K>       weak inputs -> weak
// (a + b with b <= p cannot overflow twice: a + b - 2^64 < p, so one conditional +EPS suffices;
// likewise for the borrow.)  On the host the same functions are plain canonical C, which is a
// special case of weak, so tests/hostcheck exercises the same call graph.
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

// compile-time arithmetic mod p (constants of the merged inverse stages, icrt_tail)
constexpr u64 caddm(u64 a, u64 b) {  // a, b < p
    u64 s = a + b;
    return (s < a || s >= P) ? s - P : s;
}
constexpr u64 csubm(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
constexpr u64 cmulm(u64 a, u64 b) {
    u64 r = 0;
    for (int i = 63; i >= 0; i--) {
        r = caddm(r, r);
        if ((b >> i) & 1) r = caddm(r, a);
    }
    return r;
}
constexpr u64 cpow2(int e) {
    u64 r = 1;
    for (int i = 0; i < e; i++) r = caddm(r, r);
    return r;
}

#if defined(__CUDACC__)
static __constant__ u32 c_eps = 0xFFFFFFFFu;  // 2^64 mod p as a constant-bank operand (see plus_eps_if)
#endif
SR_HD u64 mk64(u32 lo, u32 hi) { return (u64)lo | ((u64)hi << 32); }
#if defined(__CUDA_ARCH__)
// x >= p  <=>  hi == 2^32 - 1 and lo != 0, and then x - p = lo - 1: two predicate tests and two predicated
// word updates instead of a 64-bit compare, a 64-bit subtraction and two selects
SR_D u64 canon(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    if (hi == 0xFFFFFFFFu && lo != 0u) {
        lo -= 1u;
        hi = 0u;
    }
    return mk64(lo, hi);
}
#else
SR_HD u64 canon(u64 x) { return x >= P ? x - P : x; }
#endif

#if defined(__CUDA_ARCH__)
// x + m * (2^32 - 1) for m in {0, 1} when the sum is known not to overflow 64 bits: ONE wide multiply-add on the
// multiply-add pipe (IMAD.WIDE.U32) instead of a three-instruction carry chain on the ALU pipe, which is the
// binding unit of every Goldilocks kernel (profiles/r01b_gl_ncu.md)
// (EPS comes from constant memory: with an immediate, ptxas strength-reduces the product into a four-instruction
// shift / subtract / carry sequence, which is exactly what this is meant to avoid.)
#if defined(SR_GL_EPS_ON_ALU)
#define SR_GL_ADD_ALU 1
#define SR_GL_AEM_ALU 1
#endif
#ifndef SR_GL_ADD_ALU
#define SR_GL_ADD_ALU 0  // the correction of add() on the ALU pipe
#endif
#ifndef SR_GL_AEM_ALU
#define SR_GL_AEM_ALU 0  // the correction of add_eps_mul() (shift-reductions, accumulator folds) on the ALU pipe
#endif
template <bool ON_ALU>
SR_D u64 plus_eps_if(u32 lo, u32 hi, u32 m) {
    if (ON_ALU) {
    // kernels whose binding unit is the multiply-add pipe (the mat-vec family) keep the correction on the ALU pipe
    u32 r0, r1;
    asm("sub.cc.u32   %0, %2, %4;\n\t"        // + 2^32 - 1 = - m, + m * 2^32
        "subc.u32     %1, %3, 0;\n\t"
        "add.u32      %1, %1, %4;\n\t"
        : "=&r"(r0), "=&r"(r1)
        : "r"(lo), "r"(hi), "r"(m));
    return mk64(r0, r1);
    }
    u32 r0, r1;  // the lo / hi pair is what ptxas fuses into one IMAD.WIDE.U32 with a 64-bit addend
    asm("mad.lo.cc.u32   %0, %4, %5, %2;\n\t"
        "madc.hi.u32     %1, %4, %5, %3;\n\t"
        : "=&r"(r0), "=r"(r1)
        : "r"(lo), "r"(hi), "r"(m), "r"(c_eps));
    return mk64(r0, r1);
}
// a weak, b canonical -> weak
SR_D u64 add(u64 a, u64 b) {
    u32 lo, hi, m;
    asm("add.cc.u32   %0, %3, %5;\n\t"
        "addc.cc.u32  %1, %4, %6;\n\t"
        "addc.u32     %2, 0, 0;\n\t"          // m = carry (add-flags are never fed to subc)
        : "=&r"(lo), "=&r"(hi), "=r"(m)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return plus_eps_if<SR_GL_ADD_ALU>(lo, hi, m);  // + EPS when the sum wrapped: a + b - 2^64 < p, no second wrap
}
// The same sum with the correction as two PREDICATED adds of 2^32 - 1 (lo + 0xFFFFFFFF, carry into hi): four
// instructions on the ALU pipe and no multiply-add (the mat-vec kernels, whose wide multiply-add pipe is the
// binding unit).  a weak, b canonical -> weak.
SR_D u64 add_cc(u64 a, u64 b) {
    u32 lo, hi;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        ".reg .pred p;\n\t"
        "add.cc.u32   %0, %2, %4;\n\t"
        "addc.cc.u32  %1, %3, %5;\n\t"
        "addc.u32     m, 0, 0;\n\t"
        "setp.ne.u32  p, m, 0;\n\t"
        "@p add.cc.u32  %0, %0, 0xffffffff;\n\t"
        "@p addc.u32    %1, %1, 0;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return mk64(lo, hi);
}
SR_D u64 sub(u64 a, u64 b) {
    u32 lo, hi;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32   %0, %2, %4;\n\t"
        "subc.cc.u32  %1, %3, %5;\n\t"
        "subc.u32     m, 0, 0;\n\t"          // m = -borrow
        "sub.cc.u32   %0, %0, m;\n\t"        // - EPS when the difference wrapped
        "subc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return mk64(lo, hi);
}
// lo + hl * (2^32 - 1) for a 32-bit hl (weak result)
SR_D u64 add_eps_mul(u64 lo, u32 hl) {
    u32 r0, r1, m;
    asm("mad.lo.cc.u32   %0, %5, 0xFFFFFFFF, %3;\n\t"
        "madc.hi.cc.u32  %1, %5, 0xFFFFFFFF, %4;\n\t"
        "addc.u32        %2, 0, 0;\n\t"
        : "=&r"(r0), "=&r"(r1), "=r"(m)
        : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"(hl));
    return plus_eps_if<SR_GL_AEM_ALU>(r0, r1, m);  // the wrapped sum is < hl * EPS <= 2^64 - 2^33 + 1: adding EPS cannot wrap again
}
// (hi, lo) = 128-bit value -> weak residue.  2^64 = 2^32 - 1, 2^96 = -1 (mod p).
SR_D u64 reduce128(u64 lo, u64 hi) {
    const u64 t0 = sub(lo, hi >> 32);  // hi >> 32 < 2^32: canonical
    return add_eps_mul(t0, (u32)hi);
}
SR_D u64 neg(u64 a) { return P - canon(a); }  // weak (p itself for a = 0)
#else
// host: canonical arithmetic (a special case of the weak discipline)
SR_HD u64 add(u64 a, u64 b) {
    a = canon(a);
    u64 s = a + b;
    bool over = (s < a) | (s >= P);
    return over ? s + EPS : s;
}
SR_HD u64 add_cc(u64 a, u64 b) { return add(a, b); }
SR_HD u64 sub(u64 a, u64 b) {
    a = canon(a);
    u64 d = a - b;
    return (a < b) ? d - EPS : d;
}
SR_HD u64 reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= EPS;
    u64 t1 = (hl << 32) - hl;
    u64 r = t0 + t1;
    if (r < t1) r += EPS;
    return r >= P ? r - P : r;
}
SR_HD u64 neg(u64 a) { a = canon(a); return a ? P - a : 0; }
#endif

SR_HD u64 mul(u64 a, u64 b) { return reduce128(a * b, mul64hi(a, b)); }

// x * 2^K mod p for a compile-time K in [0, 192); weak in, weak out
template <int K>
SR_HD u64 mul_pow2(u64 x) {
    static_assert(K >= 0 && K < 192, "exponent");
    constexpr int KK = K % 96, s = KK % 32, j = KK / 32;
    // v = x << s as 96 bits (v2:v1:v0)
    const u64 lo = x << s;
    const u32 v2 = s ? (u32)(x >> (64 - s)) : 0u;
    const u32 v0 = (u32)lo, v1 = (u32)(lo >> 32);
    u64 r;
#if defined(__CUDA_ARCH__)
    // the high part of the shifted value has at most 32 bits: skip the 2^96 term of reduce128
    if (j == 0) {
        r = add_eps_mul(lo, v2);                                    // v0 + v1 2^32 + v2 2^64
    } else if (j == 1) {
        r = sub(add_eps_mul((u64)v0 << 32, v1), (u64)v2);           // v0 2^32 + v1 2^64 - v2
    } else {
        r = sub((u64)v0 * EPS, (u64)v1 | ((u64)v2 << 32));          // v0 2^64 - v1 - v2 2^32; v0 EPS < p
    }
#else
    if (j == 0) {
        r = reduce128(lo, v2);
    } else if (j == 1) {  // v0 2^32 + v1 2^64 - v2
        r = sub(reduce128((u64)v0 << 32, v1), (u64)v2);
    } else {              // v0 2^64 - v1 - v2 2^32   (v2 < 2^31: the subtrahend is canonical)
        r = sub(reduce128(0, v0), (u64)v1 | ((u64)v2 << 32));
    }
#endif
    return K >= 96 ? neg(r) : r;
}
// mul_pow2<K> returns a CANONICAL value when its last step is the subtraction of two canonical numbers (sub of
// canonical operands is canonical): the word-rotation-by-two case, v0 (2^32 - 1) < p minus (v1, v2) < 2^63.
// For K >= 96 the result is p - canon(r): canonical, or p itself for a zero input, which add / sub accept as their second
// operand just the same.
constexpr bool pow2_canonical(int K) { return K >= 96 || K / 32 == 2; }
template <int K>
SR_HD u64 mul_pow2c(u64 x) {  // canonical x * 2^K: the canonicalisation is skipped where mul_pow2 already delivers it
    const u64 r = mul_pow2<K>(x);
    return pow2_canonical(K) ? r : canon(r);
}
template <int K>
SR_HD u64 mulw(u64 x) {  // x * ROOTS_OF_UNITY_24[K]
    return mul_pow2<root_exp(K)>(x);
}

// (a, b) <- (a + w b, a - w b); a, b weak.  For w = -2^e the two outputs simply swap roles.
template <int LO, int SPAN, int K, int N>
SR_HD void bfly(u64 (&c)[N]) {
    constexpr int E = root_exp(K);
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        const u64 a = c[LO + i], t = mul_pow2c<E % 96>(c[LO + SPAN + i]);
        c[LO + i] = (E >= 96) ? sub(a, t) : add(a, t);
        c[LO + SPAN + i] = (E >= 96) ? add(a, t) : sub(a, t);
    }
}
// (a, b) <- (a + b, w (a - b)); w = -2^e turns a - b into b - a
// CANON_IN: the inputs are already canonical (skips two canonicalisations per butterfly)
template <int LO, int SPAN, int K, bool CANON_IN = false, int N = D>
SR_HD void ibfly(u64 (&c)[N]) {
    constexpr int E = root_exp(K);
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        const u64 a = c[LO + i], b = c[LO + SPAN + i];
        const u64 ca = CANON_IN ? a : canon(a), cb = CANON_IN ? b : canon(b);
        c[LO + i] = add(a, cb);
        c[LO + SPAN + i] = mul_pow2<E % 96>((E >= 96) ? sub(b, ca) : sub(a, cb));
    }
}

// ntt.rs:146-225.  Input CANONICAL (as stored in memory); output weak.
// first two stages: the four quarters then hold f mod X^6 - r^2, r^14, r^10, r^22
SR_HD void crt_stages12(u64 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        // zeta = ROOTS_OF_UNITY_24[4] = 2^160 = -2^64:  z = -z'  with z' = 2^64 b
        const u64 a = c[i], b = c[12 + i];
        const u64 zp = mul_pow2c<64>(b);
        c[i] = sub(a, zp);                 // a + zeta b
        c[12 + i] = add(add(a, b), zp);    // a + b - zeta b
    }
    bfly<0, 6, 2>(c);
    bfly<12, 6, 10>(c);
}
SR_HD void crt_stages(u64 (&c)[D]) {
    crt_stages12(c);
    bfly<0, 3, 1>(c);
    bfly<6, 3, 7>(c);
    bfly<12, 3, 5>(c);
    bfly<18, 3, 11>(c);
}

// ntt.rs:250-318.  EXTRA: additional power of two folded into the 1/8 and 1/4 scalings.  Output weak.
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
        const u64 a = c[i], cb = canon(c[12 + i]);
        const u64 kd = canon(mul(sub(a, cb), (u64)SR_GL_KAPPA));
        c[i] = mul_pow2<(189 + EXTRA) % 192>(sub(add(a, cb), kd));
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

// full transforms: canonical in, canonical out
SR_HD void crt(u64 (&c)[D]) {
    crt_stages(c);
    u64 o[D];
    homogenize(o, c);
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = canon(o[i]);
}

// ---- lazy accumulation of 64 x 64 -> 128-bit products -------------------------------------------
// The sum is kept UNREDUCED in a 160-bit accumulator held as two interleaved carry-save halves
// (E: limbs 0..4 takes lo*lo and hi*hi, O: limbs 1..3 takes the two cross products), so every partial
// product is one IMAD.WIDE.U32 with carry and no modular reduction happens until the end.
// (The host versions are the same limb arithmetic in portable C, so that tests/hostcheck runs the lazy code paths.)
struct Acc {
    u32 e0, e1, e2, e3, e4, o1, o2, o3;
};
SR_HD void acc_zero(Acc& A) { A.e0 = A.e1 = A.e2 = A.e3 = A.e4 = A.o1 = A.o2 = A.o3 = 0; }
SR_HD void acc_mad(Acc& A, u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32   %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32  %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32  %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32  %3, %6, %8, %3;\n\t"
        "addc.u32        %4, %4, 0;\n\t"
        : "+r"(A.e0), "+r"(A.e1), "+r"(A.e2), "+r"(A.e3), "+r"(A.e4)
        : "r"(al), "r"(ah), "r"(bl), "r"(bh));
    asm("mad.lo.cc.u32   %0, %3, %6, %0;\n\t"
        "madc.hi.cc.u32  %1, %3, %6, %1;\n\t"
        "addc.u32        %2, %2, 0;\n\t"
        "mad.lo.cc.u32   %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32  %1, %4, %5, %1;\n\t"
        "addc.u32        %2, %2, 0;\n\t"
        : "+r"(A.o1), "+r"(A.o2), "+r"(A.o3)
        : "r"(al), "r"(ah), "r"(bl), "r"(bh));
#else
    const u64 ll = (u64)al * bl, hh = (u64)ah * bh, lh = (u64)al * bh, hl = (u64)ah * bl;
    u64 lo = mk64(A.e0, A.e1), hi = mk64(A.e2, A.e3), t;
    t = lo + ll; u64 c = t < lo; lo = t;
    t = hi + hh; u64 c2 = t < hi; hi = t;
    t = hi + c; c2 += t < hi; hi = t;
    A.e0 = (u32)lo; A.e1 = (u32)(lo >> 32); A.e2 = (u32)hi; A.e3 = (u32)(hi >> 32); A.e4 += (u32)c2;
    u64 o = mk64(A.o1, A.o2);
    t = o + lh; A.o3 += (u32)(t < o); o = t;
    t = o + hl; A.o3 += (u32)(t < o); o = t;
    A.o1 = (u32)o; A.o2 = (u32)(o >> 32);
#endif
}
// A <- a * b (the first product of a sum: no carries into the top limbs yet)
SR_HD void acc_mul(Acc& A, u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    const u64 ll = (u64)al * bl, hh = (u64)ah * bh, lh = (u64)al * bh;
    A.e0 = (u32)ll; A.e1 = (u32)(ll >> 32); A.e2 = (u32)hh; A.e3 = (u32)(hh >> 32); A.e4 = 0;
#if defined(__CUDA_ARCH__)
    A.o1 = (u32)lh; A.o2 = (u32)(lh >> 32);
    asm("mad.lo.cc.u32   %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32  %1, %3, %4, %1;\n\t"
        "addc.u32        %2, 0, 0;\n\t"
        : "+r"(A.o1), "+r"(A.o2), "=r"(A.o3)
        : "r"(ah), "r"(bl));
#else
    const u64 o = lh + (u64)ah * bl;
    A.o1 = (u32)o; A.o2 = (u32)(o >> 32); A.o3 = (u32)(o < lh);
#endif
}
// Canonical residue of the accumulated value; valid for sums of fewer than 2^31 products.
// Merge the two halves into limbs l0..l4 (one carry chain), then with T = 2^32, T^2 = 2^64 = 2^32 - 1 and T^3 = -1 (mod p):
//   l0 + l1 T + l2 T^2 + (l3 + l4 T) T^3  =  [(l0, l1) + l2 (2^32 - 1)]  -  (l3, l4)
// i.e. one multiply-add fold of the third limb and ONE modular subtraction of the 64-bit number (l3, l4), which is
// canonical because l4 (the number of carries out of 2^128) is tiny: 16 instructions, against 26 for the signed
// X + Y T formulation this replaced.
template <bool CANON = true>
SR_HD u64 acc_reduce(const Acc& A) {
    u32 l1, l2, l3, l4;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32   %0, %4, %7;\n\t"
        "addc.cc.u32  %1, %5, %8;\n\t"
        "addc.cc.u32  %2, %6, %9;\n\t"
        "addc.u32     %3, %10, 0;\n\t"
        : "=&r"(l1), "=&r"(l2), "=&r"(l3), "=r"(l4)
        : "r"(A.e1), "r"(A.e2), "r"(A.e3), "r"(A.o1), "r"(A.o2), "r"(A.o3), "r"(A.e4));
    const u64 r = sub(add_eps_mul(mk64(A.e0, l1), l2), mk64(l3, l4));
    return CANON ? canon(r) : r;
#else
    u64 t = (u64)A.e1 + A.o1;
    l1 = (u32)t;
    t = (u64)A.e2 + A.o2 + (t >> 32);
    l2 = (u32)t;
    t = (u64)A.e3 + A.o3 + (t >> 32);
    l3 = (u32)t;
    l4 = A.e4 + (u32)(t >> 32);
    return canon(sub(reduce128(mk64(A.e0, l1), (u64)l2), mk64(l3, l4)));
#endif
}
// canonical residue of (accumulated value) * 2^128, i.e. with the Montgomery factor 2^-64 = 2^128 of a product of two
// raw limbs folded into the reduction.  Limbs l0..l4 as in acc_reduce; with T = 2^32, T^2 = 2^32 - 1 and T^3 = -1 the powers
// T^4 .. T^8 are -T, -T^2, 1, T, T^2, so
//   (l0 + l1 T + l2 T^2 + l3 T^3 + l4 T^4) T^4 = [(l2, l3) + l4 (2^32 - 1)] - [(0, l0) + l1 (2^32 - 1)]:
// two multiply-add folds and one modular subtraction, 23 instructions (the round-1 form, two signed 64-bit sums X + Y T
// folded with a bias of 16 p, took 26: ntt_mul 0.742 -> 0.760 of the HBM roofline).
SR_HD u64 acc_reduce_m128(const Acc& A) {
    u32 l1, l2, l3, l4;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32   %0, %4, %7;\n\t"
        "addc.cc.u32  %1, %5, %8;\n\t"
        "addc.cc.u32  %2, %6, %9;\n\t"
        "addc.u32     %3, %10, 0;\n\t"
        : "=&r"(l1), "=&r"(l2), "=&r"(l3), "=r"(l4)
        : "r"(A.e1), "r"(A.e2), "r"(A.e3), "r"(A.o1), "r"(A.o2), "r"(A.o3), "r"(A.e4));
    const u64 U = add_eps_mul(mk64(l2, l3), l4);
    const u64 W = canon(add_eps_mul(mk64(0u, A.e0), l1));
    return canon(sub(U, W));
#else
    u64 t = (u64)A.e1 + A.o1;
    l1 = (u32)t;
    t = (u64)A.e2 + A.o2 + (t >> 32);
    l2 = (u32)t;
    t = (u64)A.e3 + A.o3 + (t >> 32);
    l3 = (u32)t;
    l4 = A.e4 + (u32)(t >> 32);
    const u64 U = reduce128(mk64(l2, l3), (u64)l4);
    const u64 W = reduce128(mk64(0u, A.e0), (u64)l1);
    return canon(sub(U, W));
#endif
}
// canonical a ka + b kb: weak inputs, one reduction
SR_HD u64 dot2(u64 a, u64 ka, u64 b, u64 kb) {
    Acc d;
    acc_mul(d, a, ka);
    acc_mad(d, b, kb);
    return acc_reduce(d);
}

// ---- the last two inverse stages as constant-coefficient sums -------------------------------------
// ntt.rs:272-318: with u0..u3 = c[i], c[6+i], c[12+i], c[18+i] after the first inverse stage,
//   a  = u0 + u1,  a' = w22 (u0 - u1),  b = u2 + u3,  b' = w14 (u2 - u3)           (second stage)
//   out[i]    = SA (a + b - kappa (a - b))    = A0 (u0 + u1) + A1 (u2 + u3)        A0 = SA (1 - kappa), A1 = SA (1 + kappa)
//   out[12+i] = SB kappa (a - b)              = B0 (u0 + u1) + B1 (u2 + u3)        B0 = SB kappa, B1 = -B0
//   out[6+i]  = SA (a' + b' - kappa (a' - b')) = C0 (u0 - u1) + C1 (u2 - u3)       C0 = A0 w22, C1 = A1 w14
//   out[18+i] = SB kappa (a' - b')            = D0 (u0 - u1) + D1 (u2 - u3)        D0 = B0 w22, D1 = B1 w14
// (SA, SB: the final scalings 1/8, 1/4 times any extra power of two).  Each output is one lazy sum of 64 x 64 products and ONE
// reduction, instead of a butterfly (two canonicalisations, add, sub, shift-reduction), the multiplication by kappa with
// its own reduction, two more shift-reductions and the final canonicalisations: the inverse transform's last stage
// was 45% of the instructions of the ICRT kernel, all on the ALU pipe, which bounds it.
template <int EA, int EB>
struct TailK {
    static constexpr u64 KAP = SR_GL_KAPPA;
    static constexpr u64 SA = cpow2(EA), SB = cpow2(EB);
    static constexpr u64 W22 = cpow2(root_exp(22)), W14 = cpow2(root_exp(14));
    static constexpr u64 A0 = cmulm(SA, csubm(1, KAP)), A1 = cmulm(SA, caddm(1, KAP));
    static constexpr u64 B0 = cmulm(SB, KAP), B1 = P - B0;
    static constexpr u64 C0 = cmulm(A0, W22), C1 = cmulm(A1, W14);
    static constexpr u64 D0 = cmulm(B0, W22), D1 = cmulm(B1, W14);
};
// u0, u2 weak, u1, u3 CANONICAL -> the four canonical outputs; sums and differences first, two products per output
template <class K>
SR_HD void tail_dot2(u64& o0, u64& o1, u64& o2, u64& o3, u64 u0, u64 u1, u64 u2, u64 u3) {
    const u64 s01 = add(u0, u1), d01 = sub(u0, u1), s23 = add(u2, u3), d23 = sub(u2, u3);
    o0 = dot2(s01, K::A0, s23, K::A1);
    o1 = dot2(d01, K::C0, d23, K::C1);
    o2 = dot2(s01, K::B0, s23, K::B1);
    o3 = dot2(d01, K::D0, d23, K::D1);
}

// one butterfly of the first inverse stage; CA / CB: the operand is already canonical (or == p)
template <int K, bool CA, bool CB>
SR_HD void ibfly1(u64& x, u64& y) {
    constexpr int E = root_exp(K);
    const u64 a = x, b = y;
    const u64 ca = CA ? a : canon(a), cb = CB ? b : canon(b);
    x = add(a, cb);
    y = mul_pow2<E % 96>((E >= 96) ? sub(b, ca) : sub(a, cb));
}

// dehomogenisation with canonical (or p) outputs: all but one of its twiddles deliver that for free (pow2_canonical)
SR_HD void dehomogenize_c(u64 (&o)[D], const u64 (&c)[D]) {
#define MULW(k, x) ::sr::gl::mul_pow2c<::sr::gl::root_exp(k)>(x)
#define NEG ::sr::gl::neg
    SR_GL_DEHOMOG(o, c)
#undef MULW
#undef NEG
}
// first inverse stage (ntt.rs:250-270) on the values of dehomogenize_c (all canonical or p)
SR_HD void icrt_stage1(u64 (&o)[D]) {
    ibfly1<23, true, true>(o[0], o[3]);
    ibfly1<23, true, true>(o[1], o[4]);
    ibfly1<23, true, true>(o[2], o[5]);
    ibfly1<17, true, true>(o[6], o[9]);
    ibfly1<17, true, true>(o[7], o[10]);
    ibfly1<17, true, true>(o[8], o[11]);
    ibfly1<19, true, true>(o[12], o[15]);
    ibfly1<19, true, true>(o[13], o[16]);
    ibfly1<19, true, true>(o[14], o[17]);
    ibfly1<13, true, true>(o[18], o[21]);
    ibfly1<13, true, true>(o[19], o[22]);
    ibfly1<13, true, true>(o[20], o[23]);
}
// ntt.rs:240-319 with the slot isomorphism in front (ntt.rs:385-437).  Canonical in, canonical out.
SR_HD void icrt(u64 (&c)[D]) {
    u64 o[D];
    dehomogenize_c(o, c);
    icrt_stage1(o);
    typedef TailK<189, 190> K;
#pragma unroll
    for (int i = 0; i < 6; i++)
        tail_dot2<K>(c[i], c[6 + i], c[12 + i], c[18 + i], o[i], canon(o[6 + i]), o[12 + i], canon(o[18 + i]));
}

// z = x * y in F_p[u]/(u^3 - 2^RHO_EXP), then times 2^POST_EXP.  Weak in, weak out.
// y1, y2 are pre-multiplied by rho = 2^RHO_EXP (two shift-reductions), after which every output coefficient
// is one lazy sum of three 64 x 64 products: 9 accumulations and 3 reductions per slot.
template <int RHO_EXP, int POST_EXP>
SR_HD void slot_mul(u64* z, const u64* x, const u64* y) {
    const u64 x0 = x[0], x1 = x[1], x2 = x[2], y0 = y[0], y1 = y[1], y2 = y[2];
    u64 c0, c1, c2;
    const u64 r1 = mul_pow2<RHO_EXP>(y1), r2 = mul_pow2<RHO_EXP>(y2);
    Acc d0, d1, d2;
    acc_mul(d0, x0, y0); acc_mad(d0, x1, r2); acc_mad(d0, x2, r1);
    acc_mul(d1, x0, y1); acc_mad(d1, x1, y0); acc_mad(d1, x2, r2);
    acc_mul(d2, x0, y2); acc_mad(d2, x1, y1); acc_mad(d2, x2, y0);
    if (POST_EXP == 128) {  // the Montgomery factor of a raw-limb product rides on the reduction
        z[0] = acc_reduce_m128(d0);
        z[1] = acc_reduce_m128(d1);
        z[2] = acc_reduce_m128(d2);
        return;
    }
    c0 = acc_reduce(d0);
    c1 = acc_reduce(d1);
    c2 = acc_reduce(d2);
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
    for (int s = 0; s < 8; s++) {
        slot_mul<root_exp(1), 128>(&a[3 * s], &a[3 * s], &b[3 * s]);
    }
#pragma unroll
    for (int i = 0; i < D; i++) a[i] = canon(a[i]);
}

// Fused unit of the metric without the slot isomorphisms: slot s multiplied directly modulo
// X^3 - r^k_s; the Montgomery 2^-64 rides on the final scalings (EXTRA = 128).
template <int S>
SR_HD void fused_slot(u64 (&bs)[D], const u64* as) {
    constexpr int KS[8] = {1, 13, 7, 19, 5, 17, 11, 23};
    slot_mul<root_exp(KS[S]), 0>(&bs[3 * S], &as[3 * S], &bs[3 * S]);
}
// bs <- ring product (coefficient form, canonical); as = crt_stages(a), bs = crt_stages(b)
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
#pragma unroll
    for (int i = 0; i < D; i++) bs[i] = canon(bs[i]);
}

}  // namespace gl
}  // namespace sr
