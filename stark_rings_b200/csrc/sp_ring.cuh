// Starknet-prime ring F_p[X]/(X^16 + 1), p = 2^251 + 17 * 2^192 + 1: per-element transforms.
//
// Mirrors (reference, crates/ring/src/cyclotomic_ring/models/stark_prime/):
//   ntt.rs:121-235  serial_stark_prime_crt_in_place   -> sp::crt
//   ntt.rs:245-346  serial_stark_prime_icrt_in_place  -> sp::icrt
//   ntt_form.rs:159-175 with BaseCRTField = Fq        -> sp::mont_mul per slot
//
// A field element is ark-ff Fp256<MontBackend<_,4>>: 4 little-endian u64 limbs of x * 2^256 mod p,
// handled here as 8 little-endian 32-bit limbs.  Multiplication is Montgomery (CIOS, 32-bit
// limbs).  p = 1 (mod 2^32), so -p^-1 mod 2^32 = 0xFFFFFFFF and the reduction multiplier is just
// m = -t0; p has three non-zero 32-bit limbs (limb 0 = 1, limb 6 = 0x11, limb 7 = 0x08000000), so
// adding m * p costs two multiply-adds per round instead of eight.
#pragma once
#include "sr_common.cuh"
#include "sr_consts_gen.cuh"

namespace sr {
namespace sp {

constexpr int D = 16;   // coefficients per element
constexpr int L = 8;    // 32-bit limbs per coefficient
constexpr u32 P6 = 0x00000011u, P7 = 0x08000000u;  // p = 1 + P6 * 2^192 + P7 * 2^224

struct Fe {
    u32 v[L];
};

constexpr u32 p_limb(int i) { return i == 0 ? 1u : i == 6 ? P6 : i == 7 ? P7 : 0u; }

struct RootTable {
    u32 w[32][8];
};
constexpr RootTable ROOTS_MONT = {SR_SP_ROOTS_MONT};
constexpr u32 root_limb(int k, int i) { return ROOTS_MONT.w[k][i]; }
// WHICH = 0: 1/16; WHICH = 1: ROOTS_OF_UNITY_32[24] / 16 (ntt.rs:51-55), Montgomery form
constexpr u32 scale_limb(int which, int i) {
    constexpr u32 a[8] = SR_SP_SIXTEEN_INV_MONT;
    constexpr u32 b[8] = SR_SP_SIXTEEN_INV_W24_MONT;
    return which == 0 ? a[i] : b[i];
}

// r = a + b mod p (inputs canonical)
SR_HD void add(Fe& r, const Fe& a, const Fe& b) {
#if defined(__CUDA_ARCH__)
    u32 s[L], d[L], br;
    asm volatile(
        "add.cc.u32   %0, %8,  %16;\n\t"
        "addc.cc.u32  %1, %9,  %17;\n\t"
        "addc.cc.u32  %2, %10, %18;\n\t"
        "addc.cc.u32  %3, %11, %19;\n\t"
        "addc.cc.u32  %4, %12, %20;\n\t"
        "addc.cc.u32  %5, %13, %21;\n\t"
        "addc.cc.u32  %6, %14, %22;\n\t"
        "addc.u32     %7, %15, %23;\n\t"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    asm volatile(
        "sub.cc.u32   %0, %9,  1;\n\t"
        "subc.cc.u32  %1, %10, 0;\n\t"
        "subc.cc.u32  %2, %11, 0;\n\t"
        "subc.cc.u32  %3, %12, 0;\n\t"
        "subc.cc.u32  %4, %13, 0;\n\t"
        "subc.cc.u32  %5, %14, 0;\n\t"
        "subc.cc.u32  %6, %15, 0x11;\n\t"
        "subc.cc.u32  %7, %16, 0x08000000;\n\t"
        "subc.u32     %8, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(br)
        : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]));
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = br ? s[i] : d[i];
    return;
#else
    u32 s[L], d[L];
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        c += (u64)a.v[i] + b.v[i];
        s[i] = (u32)c;
        c >>= 32;
    }
    // d = s - p; keep d if no borrow (s >= p).  a + b < 2p < 2^253: no carry out of s.
    u64 br = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        u64 t = (u64)s[i] - p_limb(i) - br;
        d[i] = (u32)t;
        br = (t >> 63) & 1;
    }
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = br ? s[i] : d[i];
#endif
}
// r = a - b mod p
SR_HD void sub(Fe& r, const Fe& a, const Fe& b) {
#if defined(__CUDA_ARCH__)
    u32 d[L], br;
    asm volatile(
        "sub.cc.u32   %0, %9,  %17;\n\t"
        "subc.cc.u32  %1, %10, %18;\n\t"
        "subc.cc.u32  %2, %11, %19;\n\t"
        "subc.cc.u32  %3, %12, %20;\n\t"
        "subc.cc.u32  %4, %13, %21;\n\t"
        "subc.cc.u32  %5, %14, %22;\n\t"
        "subc.cc.u32  %6, %15, %23;\n\t"
        "subc.cc.u32  %7, %16, %24;\n\t"
        "subc.u32     %8, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(br)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // br = 0 or 0xFFFFFFFF: add p back when the subtraction borrowed
    asm volatile(
        "add.cc.u32   %0, %0, %8;\n\t"
        "addc.cc.u32  %1, %1, 0;\n\t"
        "addc.cc.u32  %2, %2, 0;\n\t"
        "addc.cc.u32  %3, %3, 0;\n\t"
        "addc.cc.u32  %4, %4, 0;\n\t"
        "addc.cc.u32  %5, %5, 0;\n\t"
        "addc.cc.u32  %6, %6, %9;\n\t"
        "addc.u32     %7, %7, %10;\n\t"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7])
        : "r"(br & 1u), "r"(br & P6), "r"(br & P7));
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = d[i];
    return;
#else
    u32 d[L];
    u64 br = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        u64 t = (u64)a.v[i] - b.v[i] - br;
        d[i] = (u32)t;
        br = (t >> 63) & 1;
    }
    // if borrowed add p back
    u32 mask = (u32)0 - (u32)br;
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        c += (u64)d[i] + (p_limb(i) & mask);
        r.v[i] = (u32)c;
        c >>= 32;
    }
#endif
}

// One CIOS round: t += a * bi; then t = (t + m p) / 2^32 with m = -t0.
SR_HD void mont_round(u32 (&t)[L + 2], const u32 (&a)[L], u32 bi) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < L; j++) {
        c += (u64)a[j] * bi + t[j];
        t[j] = (u32)c;
        c >>= 32;
    }
    c += t[L];
    t[L] = (u32)c;
    t[L + 1] = (u32)(c >> 32);
    u32 m = 0u - t[0];
    c = (t[0] != 0) ? 1 : 0;  // t0 + m = 2^32 (or 0 when t0 = 0)
#pragma unroll
    for (int j = 1; j < L; j++) {
        if (j == 6) c += (u64)m * P6;
        if (j == 7) c += (u64)m * P7;
        c += t[j];
        t[j - 1] = (u32)c;
        c >>= 32;
    }
    c += t[L];
    t[L - 1] = (u32)c;
    t[L] = t[L + 1] + (u32)(c >> 32);
}
#if defined(__CUDA_ARCH__)
// Even/odd carry-save form of the same CIOS round (no register moves): the running value is
//   T = X + Y * 2^32 + z,   X = limbs 0..8 in even-aligned pairs (0,1)(2,3)(4,5)(6,7),
//                           Y = limbs 1..8 in odd-aligned pairs  (1,2)(3,4)(5,6)(7,8)   (Y[j] = limb j+1),
//                           z = a pending addend at limb 0 (zlo + 2^32 zhi, zhi <= 1).
// a_even * b_i goes into X, a_odd * b_i into Y, each as one chain of four IMAD.WIDE with carry.  With
// m = -(X0 + zlo), m * p adds m at limb 0 (cancelling it, carry k2), 0x11 m at the X pair (6,7) and
// 2^27 m at the Y pair (7,8).  Dividing by 2^32 then swaps the roles: new X = old Y, new Y = old X >> 64,
// and old X1 plus the limb-0 carries becomes the new pending addend.
SR_D void mont_round_eo(u32 (&X)[L + 1], u32 (&Y)[L + 1], u32& zlo, u32& zhi, const u32 (&a)[L], u32 bi) {
    // One carry-flag-linked instruction sequence per round (every chain starts with the carry-out,
    // always 0, of the previous one) so that ptxas keeps a single live carry per multiplication.
    u32 nzlo, nzhi;
    asm volatile(
        "{\n\t"
        ".reg .u32 t0, tmp, m, k;\n\t"
        "mad.lo.cc.u32   %0, %21, %29, %0;\n\t"
        "madc.hi.cc.u32  %1, %21, %29, %1;\n\t"
        "madc.lo.cc.u32  %2, %23, %29, %2;\n\t"
        "madc.hi.cc.u32  %3, %23, %29, %3;\n\t"
        "madc.lo.cc.u32  %4, %25, %29, %4;\n\t"
        "madc.hi.cc.u32  %5, %25, %29, %5;\n\t"
        "madc.lo.cc.u32  %6, %27, %29, %6;\n\t"
        "madc.hi.cc.u32  %7, %27, %29, %7;\n\t"
        "addc.cc.u32     %8, %8, 0;\n\t"
        "madc.lo.cc.u32  %9,  %22, %29, %9;\n\t"
        "madc.hi.cc.u32  %10, %22, %29, %10;\n\t"
        "madc.lo.cc.u32  %11, %24, %29, %11;\n\t"
        "madc.hi.cc.u32  %12, %24, %29, %12;\n\t"
        "madc.lo.cc.u32  %13, %26, %29, %13;\n\t"
        "madc.hi.cc.u32  %14, %26, %29, %14;\n\t"
        "madc.lo.cc.u32  %15, %28, %29, %15;\n\t"
        "madc.hi.cc.u32  %16, %28, %29, %16;\n\t"
        "addc.cc.u32     t0, %0, %17;\n\t"      // limb 0: X0 + zlo
        "addc.u32        k, %18, 0;\n\t"        // k = zhi + carry
        "sub.u32         m, 0, t0;\n\t"         // m = -t0
        "add.cc.u32      tmp, t0, m;\n\t"       // carry = [t0 != 0]
        "addc.u32        k, k, 0;\n\t"
        "add.cc.u32      %19, %1, k;\n\t"       // new pending addend: X1 + k
        "addc.cc.u32     %20, 0, 0;\n\t"
        "madc.lo.cc.u32  %6, m, 0x11, %6;\n\t"  // + 0x11 m at limbs (6,7)
        "madc.hi.cc.u32  %7, m, 0x11, %7;\n\t"
        "addc.cc.u32     %8, %8, 0;\n\t"
        "madc.lo.cc.u32  %15, m, 0x08000000, %15;\n\t"  // + 2^27 m at limbs (7,8) = Y[6], Y[7]
        "madc.hi.u32     %16, m, 0x08000000, %16;\n\t"
        "}"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]),
          "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]),
          "+r"(zlo), "+r"(zhi), "=&r"(nzlo), "=&r"(nzhi)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
    zlo = nzlo;
    zhi = nzhi;
    // divide by 2^32: new X = old Y (limbs 1..8 -> 0..7), new Y = old X limbs 2..8 (-> 1..7)
    u32 nX[L + 1], nY[L + 1];
#pragma unroll
    for (int j = 0; j < L; j++) nX[j] = Y[j];
    nX[L] = 0;
#pragma unroll
    for (int j = 0; j < L - 1; j++) nY[j] = X[j + 2];
    nY[L - 1] = 0;
    nY[L] = 0;
#pragma unroll
    for (int j = 0; j <= L; j++) {
        X[j] = nX[j];
        Y[j] = nY[j];
    }
}
#endif

// r = a * b * 2^-256 mod p.  REDUCE: canonical output (for a b < p 2^256, e.g. canonical inputs).  !REDUCE: the
// final conditional subtraction is left out and the output is only < 2p -- for a b < p 2^256 still, which unreduced
// inputs below 32p times a canonical constant, or two inputs below 5.6p, satisfy (p < 2^256 / 31.99); the slot
// product of two inputs below 9p comes out below 3.6p.
template <bool REDUCE = true>
SR_HD void mont_mul_limbs(Fe& r, const u32 (&a)[L], const u32 (&b)[L]) {
#if defined(__CUDA_ARCH__)  // even/odd carry-save rounds
    u32 X[L + 1], Y[L + 1], zlo = 0, zhi = 0, t[L + 1];
#pragma unroll
    for (int i = 0; i < L + 1; i++) X[i] = Y[i] = 0;
#pragma unroll
    for (int i = 0; i < L; i++) mont_round_eo(X, Y, zlo, zhi, a, b[i]);
    {   // merge T = X + Y * 2^32 + z
        u64 c = (u64)X[0] + zlo;
        t[0] = (u32)c;
        c = (c >> 32) + (u64)X[1] + Y[0] + zhi;
        t[1] = (u32)c;
#pragma unroll
        for (int j = 2; j <= L; j++) {
            c = (c >> 32) + (u64)X[j] + Y[j - 1];
            t[j] = (u32)c;
        }
    }
#endif
#if defined(__CUDA_ARCH__)
    if (!REDUCE) {
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = t[i];
        return;
    }
    // t < 2p: d = t - p, keep t if that borrowed
    u32 d[L], br;
    asm volatile(
        "sub.cc.u32   %0, %9,  1;\n\t"
        "subc.cc.u32  %1, %10, 0;\n\t"
        "subc.cc.u32  %2, %11, 0;\n\t"
        "subc.cc.u32  %3, %12, 0;\n\t"
        "subc.cc.u32  %4, %13, 0;\n\t"
        "subc.cc.u32  %5, %14, 0;\n\t"
        "subc.cc.u32  %6, %15, 0x11;\n\t"
        "subc.cc.u32  %7, %16, 0x08000000;\n\t"
        "subc.u32     %8, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(br)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]));
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = br ? t[i] : d[i];
    return;
#else
    u32 t[L + 2];
#pragma unroll
    for (int i = 0; i < L + 2; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < L; i++) mont_round(t, a, b[i]);
    if (!REDUCE) {
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = t[i];
        return;
    }
    // t < 2p: conditional subtraction
    u32 d[L];
    u64 br = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        u64 x = (u64)t[i] - p_limb(i) - br;
        d[i] = (u32)x;
        br = (x >> 63) & 1;
    }
    bool keep_t = br && (t[L] == 0);
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = keep_t ? t[i] : d[i];
#endif
}
SR_HD void mont_mul(Fe& r, const Fe& a, const Fe& b) { mont_mul_limbs<true>(r, a.v, b.v); }
SR_HD void mont_mul_nr(Fe& r, const Fe& a, const Fe& b) { mont_mul_limbs<false>(r, a.v, b.v); }

// ---- unreduced ("lazy") arithmetic of the fused ring product ------------------------------------------------
// Values are only kept below 2^256 (31.99 p): sums are plain 256-bit additions, differences get a multiple of p
// added instead of a conditional correction, products skip the final subtraction.  The bounds are tracked statically
// in sp_quad.cuh; the host build (tests/hostcheck) checks every one of them at run time.
// r = a + b (the caller guarantees a + b < 2^256)
SR_HD void add_nr(Fe& r, const Fe& a, const Fe& b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32   %0, %8,  %16;\n\t"
        "addc.cc.u32  %1, %9,  %17;\n\t"
        "addc.cc.u32  %2, %10, %18;\n\t"
        "addc.cc.u32  %3, %11, %19;\n\t"
        "addc.cc.u32  %4, %12, %20;\n\t"
        "addc.cc.u32  %5, %13, %21;\n\t"
        "addc.cc.u32  %6, %14, %22;\n\t"
        "addc.u32     %7, %15, %23;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
    u64 c = 0;
    for (int i = 0; i < L; i++) {
        c += (u64)a.v[i] + b.v[i];
        r.v[i] = (u32)c;
        c >>= 32;
    }
    if (c) __builtin_trap();  // bound violated
#endif
}
// r = a - b + K p  (K p >= b and a - b + K p < 2^256 guaranteed by the caller; K <= 16)
SR_HD void sub_kp(Fe& r, const Fe& a, const Fe& b, u32 K) {
    const u32 k6 = K * P6, k7 = K * P7;  // K p = K + k6 2^192 + k7 2^224
#if defined(__CUDA_ARCH__)
    u32 d[L];
    asm("sub.cc.u32   %0, %8,  %16;\n\t"
        "subc.cc.u32  %1, %9,  %17;\n\t"
        "subc.cc.u32  %2, %10, %18;\n\t"
        "subc.cc.u32  %3, %11, %19;\n\t"
        "subc.cc.u32  %4, %12, %20;\n\t"
        "subc.cc.u32  %5, %13, %21;\n\t"
        "subc.cc.u32  %6, %14, %22;\n\t"
        "subc.u32     %7, %15, %23;\n\t"
        : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    asm("add.cc.u32   %0, %8,  %16;\n\t"   // the wrap of a - b modulo 2^256 cancels against this carry-out
        "addc.cc.u32  %1, %9,  0;\n\t"
        "addc.cc.u32  %2, %10, 0;\n\t"
        "addc.cc.u32  %3, %11, 0;\n\t"
        "addc.cc.u32  %4, %12, 0;\n\t"
        "addc.cc.u32  %5, %13, 0;\n\t"
        "addc.cc.u32  %6, %14, %17;\n\t"
        "addc.u32     %7, %15, %18;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7])
        : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]), "r"(K), "r"(k6), "r"(k7));
#else
    // exact integer arithmetic with the bound checks
    u32 kp[L] = {K, 0, 0, 0, 0, 0, k6, k7};
    u32 t[L];
    u64 c = 0;
    for (int i = 0; i < L; i++) {
        c += (u64)a.v[i] + kp[i];
        t[i] = (u32)c;
        c >>= 32;
    }
    if (c) __builtin_trap();  // a + K p < 2^256 is implied by the documented bounds (a - b + K p < 2^256, b small)
    u64 br = 0;
    for (int i = 0; i < L; i++) {
        u64 x = (u64)t[i] - b.v[i] - br;
        r.v[i] = (u32)x;
        br = (x >> 63) & 1;
    }
    if (br) __builtin_trap();  // K p < b
#endif
}
// x < 2^256  ->  r = x mod p + {2p or 3p-ish}: r = x + 2p - q 2p with q = floor(x / 2^252), which lies in
// [2p - 2^200, 4p): since 2p = 2^252 + delta (delta < 2^198), x - q 2p = (x mod 2^252) - q delta > -2^202.
SR_HD void partial_reduce(Fe& r, const Fe& x) {
    const u32 q = x.v[7] >> 28;  // 0 .. 15
    // 2p = 2 + 0x22 2^192 + 2^28 2^224
    Fe two_p;
    for (int i = 0; i < L; i++) two_p.v[i] = 0;
    two_p.v[0] = 2; two_p.v[6] = 2 * P6; two_p.v[7] = 2 * P7;
    Fe q2p;
    for (int i = 0; i < L; i++) q2p.v[i] = 0;
    q2p.v[0] = 2 * q; q2p.v[6] = 2 * P6 * q; q2p.v[7] = (2 * P7) * q;  // q 2^28 <= 15 2^28 < 2^32
    Fe t;
#if defined(__CUDA_ARCH__)
    // t = x - q 2p (may wrap), r = t + 2p: the same two chains as sub_kp
    asm("sub.cc.u32   %0, %8,  %16;\n\t"
        "subc.cc.u32  %1, %9,  0;\n\t"
        "subc.cc.u32  %2, %10, 0;\n\t"
        "subc.cc.u32  %3, %11, 0;\n\t"
        "subc.cc.u32  %4, %12, 0;\n\t"
        "subc.cc.u32  %5, %13, 0;\n\t"
        "subc.cc.u32  %6, %14, %17;\n\t"
        "subc.u32     %7, %15, %18;\n\t"
        : "=&r"(t.v[0]), "=&r"(t.v[1]), "=&r"(t.v[2]), "=&r"(t.v[3]), "=&r"(t.v[4]), "=&r"(t.v[5]), "=&r"(t.v[6]), "=&r"(t.v[7])
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(q2p.v[0]), "r"(q2p.v[6]), "r"(q2p.v[7]));
    asm("add.cc.u32   %0, %8,  2;\n\t"
        "addc.cc.u32  %1, %9,  0;\n\t"
        "addc.cc.u32  %2, %10, 0;\n\t"
        "addc.cc.u32  %3, %11, 0;\n\t"
        "addc.cc.u32  %4, %12, 0;\n\t"
        "addc.cc.u32  %5, %13, 0;\n\t"
        "addc.cc.u32  %6, %14, 0x22;\n\t"
        "addc.u32     %7, %15, 0x10000000;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7])
        : "r"(t.v[0]), "r"(t.v[1]), "r"(t.v[2]), "r"(t.v[3]), "r"(t.v[4]), "r"(t.v[5]), "r"(t.v[6]), "r"(t.v[7]));
#else
    u64 c = 0;
    for (int i = 0; i < L; i++) {  // x + 2p: below 2^256 + 2p, so one extra bit
        c += (u64)x.v[i] + two_p.v[i];
        t.v[i] = (u32)c;
        c >>= 32;
    }
    u64 top = c, br = 0;
    for (int i = 0; i < L; i++) {
        u64 y = (u64)t.v[i] - q2p.v[i] - br;
        r.v[i] = (u32)y;
        br = (y >> 63) & 1;
    }
    if (top != br) __builtin_trap();  // the result must be in [0, 2^256)
    if ((r.v[7] >> 28) >= 4 + 1) __builtin_trap();  // and below 4p < 5 2^252 (loose check)
#endif
}

// x < 2^256 (in practice the < 9p outputs of the unreduced forward transform)  ->  r = x mod p, canonical, in ONE step:
// q = floor(x / 2^251) (top limb >> 27, at most 31); since p = 2^251 + delta with delta = 17 2^192 + 1 < 2^197,
// x - q p = (x mod 2^251) - q delta lies in (-2^202, 2^251): it is the canonical residue, or negative, in which case adding p
// once gives a value in (p - 2^202, p).  A sparse subtraction chain, the borrow as a mask, a sparse masked addition
// chain: about the cost of ONE conditional subtraction of the reduced arithmetic, paid once per stored value
// instead of after every multiplication, addition and subtraction.
SR_HD void canon_small(Fe& r, const Fe& x) {
    const u32 q = x.v[7] >> 27;
    const u32 q6 = q * P6, q7 = q << 27;  // q p = q + q6 2^192 + q7 2^224
#if defined(__CUDA_ARCH__)
    u32 t[L], m;
    asm("sub.cc.u32   %0, %9,  %17;\n\t"
        "subc.cc.u32  %1, %10, 0;\n\t"
        "subc.cc.u32  %2, %11, 0;\n\t"
        "subc.cc.u32  %3, %12, 0;\n\t"
        "subc.cc.u32  %4, %13, 0;\n\t"
        "subc.cc.u32  %5, %14, 0;\n\t"
        "subc.cc.u32  %6, %15, %18;\n\t"
        "subc.cc.u32  %7, %16, %19;\n\t"
        "subc.u32     %8, 0, 0;\n\t"          // all ones when x - q p < 0
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(m)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(q), "r"(q6), "r"(q7));
    const u32 m0 = m & 1u, m6 = m & P6, m7 = m & P7;
    asm("add.cc.u32   %0, %8,  %16;\n\t"
        "addc.cc.u32  %1, %9,  0;\n\t"
        "addc.cc.u32  %2, %10, 0;\n\t"
        "addc.cc.u32  %3, %11, 0;\n\t"
        "addc.cc.u32  %4, %12, 0;\n\t"
        "addc.cc.u32  %5, %13, 0;\n\t"
        "addc.cc.u32  %6, %14, %17;\n\t"
        "addc.u32     %7, %15, %18;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(m0), "r"(m6), "r"(m7));
#else
    const u32 qp[L] = {q, 0, 0, 0, 0, 0, q6, q7};
    u32 t[L];
    u64 br = 0;
    for (int i = 0; i < L; i++) {
        const u64 y = (u64)x.v[i] - qp[i] - br;
        t[i] = (u32)y;
        br = (y >> 63) & 1;
    }
    u64 c = 0;
    for (int i = 0; i < L; i++) {
        c += (u64)t[i] + (br ? p_limb(i) : 0u);
        r.v[i] = (u32)c;
        c >>= 32;
    }
    if (c != br) __builtin_trap();  // the correction must bring a negative difference back into [0, 2^256)
    // the result must be canonical (< p): compare from the top limb down
    for (int i = L - 1; i >= 0; i--) {
        if (r.v[i] < p_limb(i)) break;
        if (r.v[i] > p_limb(i) || i == 0) __builtin_trap();
    }
#endif
}

// ---- sums of products (the mat-vec family): unreduced ------------------------------------------------------------------
// sum_c a_c x_c with canonical factors: every product skips its final subtraction (mont_mul_nr, < 2p), the running sum
// is a plain 256-bit addition, and it is brought back below p (canon_small) after every 15 products: p + 15 * 2p = 31p
// stays below 2^256 = 31.99p.  Saves the conditional subtraction and the modular addition (8 + 9 + 8 selects) of
// every product.
struct DotAcc {
    Fe v;
    int n;  // products added since v was last canonical
};
constexpr int DOT_BURST = 15;
SR_HD void dot_zero(DotAcc& A) {
    for (int i = 0; i < L; i++) A.v.v[i] = 0;
    A.n = 0;
}
SR_HD void dot_mad(DotAcc& A, const Fe& a, const Fe& x) {
    Fe t, s;
    mont_mul_nr(t, a, x);
    add_nr(s, A.v, t);
    A.v = s;
    if (++A.n == DOT_BURST) {
        canon_small(s, A.v);
        A.v = s;
        A.n = 0;
    }
}
SR_HD void dot_result(Fe& r, const DotAcc& A) { canon_small(r, A.v); }

// r = a * ROOTS_OF_UNITY_32[K] (constant in Montgomery form, limbs as immediates)
template <int K>
SR_HD void mulw(Fe& r, const Fe& a) {
    constexpr u32 w[L] = {root_limb(K, 0), root_limb(K, 1), root_limb(K, 2), root_limb(K, 3),
                          root_limb(K, 4), root_limb(K, 5), root_limb(K, 6), root_limb(K, 7)};
    mont_mul_limbs(r, a.v, w);
}
template <int WHICH>
SR_HD void mul_scale(Fe& r, const Fe& a) {
    constexpr u32 w[L] = {scale_limb(WHICH, 0), scale_limb(WHICH, 1), scale_limb(WHICH, 2), scale_limb(WHICH, 3),
                          scale_limb(WHICH, 4), scale_limb(WHICH, 5), scale_limb(WHICH, 6), scale_limb(WHICH, 7)};
    mont_mul_limbs(r, a.v, w);
}

template <int I, int SPAN, int K>
SR_HD void bfly(Fe (&c)[D]) {  // (a, b) <- (a + w b, a - w b)
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        Fe t, a = c[I + i];
        mulw<K>(t, c[I + SPAN + i]);
        add(c[I + i], a, t);
        sub(c[I + SPAN + i], a, t);
    }
}
template <int I, int SPAN, int K>
SR_HD void ibfly(Fe (&c)[D]) {  // (a, b) <- (a + b, w (a - b))
#pragma unroll
    for (int i = 0; i < SPAN; i++) {
        Fe a = c[I + i], b = c[I + SPAN + i], d;
        add(c[I + i], a, b);
        sub(d, a, b);
        mulw<K>(c[I + SPAN + i], d);
    }
}

// ntt.rs:121-235
SR_HD void crt(Fe (&c)[D]) {
    bfly<0, 8, 8>(c);
    bfly<0, 4, 4>(c);
    bfly<8, 4, 12>(c);
    bfly<0, 2, 2>(c);
    bfly<4, 2, 10>(c);
    bfly<8, 2, 6>(c);
    bfly<12, 2, 14>(c);
    bfly<0, 1, 1>(c);
    bfly<2, 1, 9>(c);
    bfly<4, 1, 5>(c);
    bfly<6, 1, 13>(c);
    bfly<8, 1, 3>(c);
    bfly<10, 1, 11>(c);
    bfly<12, 1, 7>(c);
    bfly<14, 1, 15>(c);
}
// ntt.rs:245-346
SR_HD void icrt(Fe (&c)[D]) {
    ibfly<0, 1, 31>(c);
    ibfly<2, 1, 23>(c);
    ibfly<4, 1, 27>(c);
    ibfly<6, 1, 19>(c);
    ibfly<8, 1, 29>(c);
    ibfly<10, 1, 21>(c);
    ibfly<12, 1, 25>(c);
    ibfly<14, 1, 17>(c);
    ibfly<0, 2, 30>(c);
    ibfly<4, 2, 22>(c);
    ibfly<8, 2, 26>(c);
    ibfly<12, 2, 18>(c);
    ibfly<0, 4, 28>(c);
    ibfly<8, 4, 20>(c);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        Fe a = c[i], b = c[8 + i], s, d;
        add(s, a, b);
        sub(d, a, b);
        mul_scale<0>(c[i], s);
        mul_scale<1>(c[8 + i], d);
    }
}

}  // namespace sp
}  // namespace sr
