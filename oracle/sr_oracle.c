/* CPU restatement (plain C) of the stark-rings hot path: crt / icrt / NTT-form multiply /
 * fused ring multiply / ring matrix-vector product for the Goldilocks, BabyBear and
 * Starknet-prime ring models, on the reference's raw memory layout (ark-ff MontBackend:
 * little-endian u64 limbs of x * 2^(64 N) mod p).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.  The product
 * (stark_rings_b200/, libstarkrings_cuda.so) never links or calls it.
 *
 * Parity: pinned.  oracle/ref_py.py reproduces every known-answer vector of the reference's
 * tests (tests/golden/, tests/test_oracle.py); this file is checked bit-for-bit against
 * ref_py.py on those vectors and on random inputs (tests/test_c_oracle.py).
 * The reference itself (Rust + un-vendored ark-ff 0.4.2) cannot be built here: no cargo/rustc.
 *
 * It follows the reference's schedule statement by statement:
 *   goldilocks/ntt.rs:135-228 (crt), :240-319 (icrt), :326-437 (slot isomorphisms)
 *   babybear/ntt.rs:143-236, :238-317, :324-588
 *   stark_prime/ntt.rs:121-235, :245-346
 *   ntt_form.rs:159-189 (slot-wise Mul), :588-601,640-654 (Add, Sum)
 *   crt.rs:10-25,34-49 (elementwise_crt / elementwise_icrt: serial loop over the batch)
 *   linear_algebra/src/matrix.rs:168-178 (checked_mul_vec)
 * Field arithmetic restates ark-ff 0.4.2 (pinned in the reference's Cargo.lock): Montgomery
 * multiplication (CIOS) on N limbs, add/sub with conditional correction, and CubicExtField's
 * Karatsuba multiplication (6 base multiplications + 2 by the non-residue); Fq9 is the cubic
 * extension of Fq3 with non-residue u (babybear/fq9.rs:19-56).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sr_oracle_consts.h"

typedef unsigned __int128 u128;
typedef uint64_t u64;

enum { SRO_GL = 0, SRO_BB = 1, SRO_SP = 2 };

/* ------------------------------------------------------------------ 1-limb fields ---- */
typedef struct {
    u64 p, ninv; /* ninv = -p^-1 mod 2^64 */
    u64 r2;      /* 2^128 mod p */
    u64 W[24];   /* Montgomery form roots */
    u64 kappa, eight_inv, four_inv, one;
} f1_ctx;

static f1_ctx GL, BB;

static inline u64 f1_add(const f1_ctx* F, u64 a, u64 b) {
    u128 s = (u128)a + b;
    if (s >= F->p) s -= F->p;
    return (u64)s;
}
static inline u64 f1_sub(const f1_ctx* F, u64 a, u64 b) { return a >= b ? a - b : a + (F->p - b); }
static inline u64 f1_neg(const f1_ctx* F, u64 a) { return a ? F->p - a : 0; }
static inline u64 f1_mul(const f1_ctx* F, u64 a, u64 b) {
    u128 t = (u128)a * b;
    u64 m = (u64)t * F->ninv;
    u128 mp = (u128)m * F->p;
    u128 r = (t >> 64) + (mp >> 64) + ((u64)t != 0); /* low halves cancel to 2^64 or 0 */
    if (r >= F->p) r -= F->p;
    return (u64)r;
}

static void f1_init(f1_ctx* F, u64 p, const u64* roots, u64 kappa, u64 e8, u64 e4) {
    F->p = p;
    u64 inv = 1; /* Newton: inv = p^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) inv *= 2 - p * inv;
    F->ninv = (u64)0 - inv;
    u64 x = 1 % p;
    for (int i = 0; i < 128; i++) x = f1_add(F, x, x);
    F->r2 = x;
    for (int i = 0; i < 24; i++) F->W[i] = f1_mul(F, roots[i], F->r2);
    F->kappa = f1_mul(F, kappa, F->r2);
    F->eight_inv = f1_mul(F, e8, F->r2);
    F->four_inv = f1_mul(F, e4, F->r2);
    F->one = f1_mul(F, 1, F->r2);
}

/* (a,b) <- (a + w b, a - w b) on c[lo+i], c[lo+span+i] */
static void f1_bfly(const f1_ctx* F, u64* c, int lo, int span, u64 w) {
    for (int i = 0; i < span; i++) {
        u64 a = c[lo + i], t = f1_mul(F, w, c[lo + span + i]);
        c[lo + i] = f1_add(F, a, t);
        c[lo + span + i] = f1_sub(F, a, t);
    }
}
/* (a,b) <- (a + b, w (a - b)) */
static void f1_ibfly(const f1_ctx* F, u64* c, int lo, int span, u64 w) {
    for (int i = 0; i < span; i++) {
        u64 a = c[lo + i], b = c[lo + span + i];
        c[lo + i] = f1_add(F, a, b);
        c[lo + span + i] = f1_mul(F, w, f1_sub(F, a, b));
    }
}

/* goldilocks/ntt.rs:146-225, babybear/ntt.rs:154-233 */
static void f1_crt_stages(const f1_ctx* F, u64* c, int D) {
    int h = D / 2, q = D / 4, e = D / 8;
    for (int i = 0; i < h; i++) {
        u64 a = c[i], b = c[h + i], z = f1_mul(F, F->W[4], b);
        c[i] = f1_add(F, a, z);
        c[h + i] = f1_sub(F, f1_add(F, a, b), z);
    }
    f1_bfly(F, c, 0, q, F->W[2]);
    f1_bfly(F, c, h, q, F->W[10]);
    f1_bfly(F, c, 0, e, F->W[1]);
    f1_bfly(F, c, q, e, F->W[7]);
    f1_bfly(F, c, h, e, F->W[5]);
    f1_bfly(F, c, 3 * q, e, F->W[11]);
}
/* goldilocks/ntt.rs:250-318, babybear/ntt.rs:249-316 */
static void f1_icrt_stages(const f1_ctx* F, u64* c, int D) {
    int h = D / 2, q = D / 4, e = D / 8;
    f1_ibfly(F, c, 0, e, F->W[23]);
    f1_ibfly(F, c, q, e, F->W[17]);
    f1_ibfly(F, c, h, e, F->W[19]);
    f1_ibfly(F, c, 3 * q, e, F->W[13]);
    f1_ibfly(F, c, 0, q, F->W[22]);
    f1_ibfly(F, c, h, q, F->W[14]);
    for (int i = 0; i < h; i++) {
        u64 a = c[i], b = c[h + i];
        u64 kd = f1_mul(F, F->kappa, f1_sub(F, a, b));
        c[i] = f1_mul(F, F->eight_inv, f1_sub(F, f1_add(F, a, b), kd));
        c[h + i] = f1_mul(F, F->four_inv, kd);
    }
}

#define MULW(F, x, k) f1_mul(F, x, (F)->W[k])

/* ---- Goldilocks slot isomorphisms, goldilocks/ntt.rs:350-437 ---- */
static void gl_homogenize(u64* c) {
    const f1_ctx* F = &GL;
    u64 t;
    c[4] = f1_neg(F, c[4]);                                                 /* 13 */
    c[7] = MULW(F, c[7], 2);   c[8] = MULW(F, c[8], 4);                     /* 7  */
    c[10] = MULW(F, c[10], 6); c[11] = MULW(F, c[11], 12);                  /* 19 */
    t = c[13]; c[13] = MULW(F, c[14], 3);  c[14] = MULW(F, t, 1);           /* 5  */
    t = c[16]; c[16] = MULW(F, c[17], 11); c[17] = MULW(F, t, 5);           /* 17 */
    t = c[19]; c[19] = MULW(F, c[20], 7);  c[20] = MULW(F, t, 3);           /* 11 */
    t = c[22]; c[22] = MULW(F, c[23], 15); c[23] = MULW(F, t, 7);           /* 23 */
}
static void gl_dehomogenize(u64* c) {
    const f1_ctx* F = &GL;
    u64 t;
    c[4] = f1_neg(F, c[4]);
    c[7] = MULW(F, c[7], 22);  c[8] = MULW(F, c[8], 20);
    c[10] = MULW(F, c[10], 18); c[11] = MULW(F, c[11], 12);
    t = c[13]; c[13] = MULW(F, c[14], 23); c[14] = MULW(F, t, 21);
    t = c[16]; c[16] = MULW(F, c[17], 19); c[17] = MULW(F, t, 13);
    t = c[19]; c[19] = MULW(F, c[20], 21); c[20] = MULW(F, t, 17);
    t = c[22]; c[22] = MULW(F, c[23], 17); c[23] = MULW(F, t, 9);
}

/* ---- BabyBear slot isomorphisms, babybear/ntt.rs:351-588 ----
 * Each map is "dst[i] = src[j] * W[k]" (k = 0: plain copy, k = 12: negation), followed
 * (homogenize) or preceded (dehomogenize) by the (1 3)(2 6)(5 7) transpose.  Tables give,
 * per destination index i (before the transpose), the source index and root exponent. */
typedef struct { int8_t src[9], k[9]; } bb_map;
static const bb_map BB_H[8] = {
    /* 1  */ {{0, 1, 2, 3, 4, 5, 6, 7, 8}, {0, 0, 0, 0, 0, 0, 0, 0, 0}},
    /* 13 */ {{0, 7, 5, 3, 1, 8, 6, 4, 2}, {0, 10, 7, 4, 1, 11, 8, 5, 2}},
    /* 7  */ {{0, 4, 8, 3, 7, 2, 6, 1, 5}, {0, 3, 6, 2, 5, 1, 4, 0, 3}},
    /* 19 */ {{0, 1, 2, 3, 4, 5, 6, 7, 8}, {0, 2, 4, 6, 8, 10, 12, 14, 16}},
    /* 5  */ {{0, 2, 4, 6, 8, 1, 3, 5, 7}, {0, 1, 2, 3, 4, 0, 1, 2, 3}},
    /* 17 */ {{0, 8, 7, 6, 5, 4, 3, 2, 1}, {0, 15, 13, 11, 9, 7, 5, 3, 1}},
    /* 11 */ {{0, 5, 1, 6, 2, 7, 3, 8, 4}, {0, 6, 1, 7, 2, 8, 3, 9, 4}},
    /* 23 */ {{0, 2, 4, 6, 8, 1, 3, 5, 7}, {0, 5, 10, 15, 20, 2, 7, 12, 17}},
};
static const bb_map BB_DH[8] = {
    /* 1  */ {{0, 1, 2, 3, 4, 5, 6, 7, 8}, {0, 0, 0, 0, 0, 0, 0, 0, 0}},
    /* 13 */ {{0, 4, 8, 3, 7, 2, 6, 1, 5}, {0, 23, 22, 20, 19, 17, 16, 14, 13}},
    /* 7  */ {{0, 7, 5, 3, 1, 8, 6, 4, 2}, {0, 0, 23, 22, 21, 21, 20, 19, 18}},
    /* 19 */ {{0, 1, 2, 3, 4, 5, 6, 7, 8}, {0, 22, 20, 18, 16, 14, 12, 10, 8}},
    /* 5  */ {{0, 5, 1, 6, 2, 7, 3, 8, 4}, {0, 0, 23, 23, 22, 22, 21, 21, 20}},
    /* 17 */ {{0, 8, 7, 6, 5, 4, 3, 2, 1}, {0, 23, 21, 19, 17, 15, 13, 11, 9}},
    /* 11 */ {{0, 2, 4, 6, 8, 1, 3, 5, 7}, {0, 23, 22, 21, 20, 18, 17, 16, 15}},
    /* 23 */ {{0, 5, 1, 6, 2, 7, 3, 8, 4}, {0, 22, 19, 17, 14, 12, 9, 7, 4}},
};
static void bb_transpose(u64* c) {
    u64 t;
    t = c[1]; c[1] = c[3]; c[3] = t;
    t = c[2]; c[2] = c[6]; c[6] = t;
    t = c[5]; c[5] = c[7]; c[7] = t;
}
static void bb_apply(const bb_map* m, u64* c) {
    u64 s[9];
    memcpy(s, c, sizeof s);
    for (int i = 0; i < 9; i++) c[i] = m->k[i] ? f1_mul(&BB, s[m->src[i]], BB.W[m->k[i]]) : s[m->src[i]];
}
static void bb_homogenize(u64* c) {
    for (int s = 0; s < 8; s++) { bb_apply(&BB_H[s], c + 9 * s); bb_transpose(c + 9 * s); }
}
static void bb_dehomogenize(u64* c) {
    for (int s = 0; s < 8; s++) { bb_transpose(c + 9 * s); bb_apply(&BB_DH[s], c + 9 * s); }
}

/* ---- cubic extensions (ark-ff CubicExtField::mul_assign, Karatsuba) ---- */
static void f3_mul(const f1_ctx* F, u64* x, const u64* y) { /* x *= y in Fq[u]/(u^3 - W[1]) */
    u64 d = x[0], e = x[1], f = x[2], a = y[0], b = y[1], c = y[2];
    u64 ad = f1_mul(F, d, a), be = f1_mul(F, e, b), cf = f1_mul(F, f, c);
    u64 X = f1_sub(F, f1_sub(F, f1_mul(F, f1_add(F, e, f), f1_add(F, b, c)), be), cf);
    u64 Y = f1_sub(F, f1_sub(F, f1_mul(F, f1_add(F, d, e), f1_add(F, a, b)), ad), be);
    u64 Z = f1_sub(F, f1_add(F, f1_sub(F, f1_mul(F, f1_add(F, d, f), f1_add(F, a, c)), ad), be), cf);
    x[0] = f1_add(F, ad, f1_mul(F, X, F->W[1]));
    x[1] = f1_add(F, Y, f1_mul(F, cf, F->W[1]));
    x[2] = Z;
}
static void f3_addv(const f1_ctx* F, u64* r, const u64* a, const u64* b) { for (int i = 0; i < 3; i++) r[i] = f1_add(F, a[i], b[i]); }
static void f3_subv(const f1_ctx* F, u64* r, const u64* a, const u64* b) { for (int i = 0; i < 3; i++) r[i] = f1_sub(F, a[i], b[i]); }
static void f3_mul_by_u(const f1_ctx* F, u64* x) { /* babybear/fq9.rs:19-26 */
    u64 c2 = x[2];
    x[2] = x[1]; x[1] = x[0]; x[0] = f1_mul(F, c2, F->W[1]);
}
static void f9_mul(u64* x, const u64* y) { /* x *= y, x = (c0,c1,c2) of Fq3 each */
    const f1_ctx* F = &BB;
    u64 ad[3], be[3], cf[3], s1[3], s2[3], X[3], Y[3], Z[3];
    memcpy(ad, x, 24);     f3_mul(F, ad, y);
    memcpy(be, x + 3, 24); f3_mul(F, be, y + 3);
    memcpy(cf, x + 6, 24); f3_mul(F, cf, y + 6);
    f3_addv(F, s1, x + 3, x + 6); f3_addv(F, s2, y + 3, y + 6); f3_mul(F, s1, s2);
    f3_subv(F, X, s1, be); f3_subv(F, X, X, cf);
    f3_addv(F, s1, x, x + 3); f3_addv(F, s2, y, y + 3); f3_mul(F, s1, s2);
    f3_subv(F, Y, s1, ad); f3_subv(F, Y, Y, be);
    f3_addv(F, s1, x, x + 6); f3_addv(F, s2, y, y + 6); f3_mul(F, s1, s2);
    f3_subv(F, Z, s1, ad); f3_addv(F, Z, Z, be); f3_subv(F, Z, Z, cf);
    f3_mul_by_u(F, X);
    f3_addv(F, x, ad, X);
    f3_mul_by_u(F, cf);
    f3_addv(F, x + 3, Y, cf);
    memcpy(x + 6, Z, 24);
}

/* ------------------------------------------------------------------ Starknet prime ---- */
typedef struct { u64 v[4]; } fp4;
static fp4 SP_P, SP_R2, SP_W[32], SP_16INV, SP_16INV_W24;
static u64 SP_NINV;

static inline int fp4_geq(const fp4* a, const fp4* b) {
    for (int i = 3; i >= 0; i--) if (a->v[i] != b->v[i]) return a->v[i] > b->v[i];
    return 1;
}
static inline u64 fp4_add_raw(fp4* r, const fp4* a, const fp4* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 fp4_sub_raw(fp4* r, const fp4* a, const fp4* b) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->v[i] - b->v[i] - br;
        r->v[i] = (u64)d; br = (u64)(d >> 64) & 1;
    }
    return br;
}
static inline void sp_add(fp4* r, const fp4* a, const fp4* b) {
    fp4_add_raw(r, a, b); /* p < 2^252: no carry out */
    if (fp4_geq(r, &SP_P)) fp4_sub_raw(r, r, &SP_P);
}
static inline void sp_sub(fp4* r, const fp4* a, const fp4* b) {
    if (fp4_sub_raw(r, a, b)) fp4_add_raw(r, r, &SP_P);
}
static void sp_mul(fp4* r, const fp4* a, const fp4* b) { /* CIOS Montgomery, R = 2^256 */
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
        u64 m = t[0] * SP_NINV;
        c = (u128)m * SP_P.v[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * SP_P.v[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    fp4 x = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fp4_geq(&x, &SP_P)) fp4_sub_raw(&x, &x, &SP_P);
    *r = x;
}
static void sp_bfly(fp4* c, int lo, int span, const fp4* w) {
    for (int i = 0; i < span; i++) {
        fp4 a = c[lo + i], t;
        sp_mul(&t, w, &c[lo + span + i]);
        sp_add(&c[lo + i], &a, &t);
        sp_sub(&c[lo + span + i], &a, &t);
    }
}
static void sp_ibfly(fp4* c, int lo, int span, const fp4* w) {
    for (int i = 0; i < span; i++) {
        fp4 a = c[lo + i], b = c[lo + span + i], d;
        sp_add(&c[lo + i], &a, &b);
        sp_sub(&d, &a, &b);
        sp_mul(&c[lo + span + i], w, &d);
    }
}
/* stark_prime/ntt.rs:121-235 */
static void sp_crt(fp4* c) {
    static const int s2[2] = {4, 12}, s3[4] = {2, 10, 6, 14}, s4[8] = {1, 9, 5, 13, 3, 11, 7, 15};
    sp_bfly(c, 0, 8, &SP_W[8]);
    for (int b = 0; b < 2; b++) sp_bfly(c, 8 * b, 4, &SP_W[s2[b]]);
    for (int b = 0; b < 4; b++) sp_bfly(c, 4 * b, 2, &SP_W[s3[b]]);
    for (int b = 0; b < 8; b++) sp_bfly(c, 2 * b, 1, &SP_W[s4[b]]);
}
/* stark_prime/ntt.rs:245-346 */
static void sp_icrt(fp4* c) {
    static const int s1[8] = {31, 23, 27, 19, 29, 21, 25, 17}, s2[4] = {30, 22, 26, 18}, s3[2] = {28, 20};
    for (int b = 0; b < 8; b++) sp_ibfly(c, 2 * b, 1, &SP_W[s1[b]]);
    for (int b = 0; b < 4; b++) sp_ibfly(c, 4 * b, 2, &SP_W[s2[b]]);
    for (int b = 0; b < 2; b++) sp_ibfly(c, 8 * b, 4, &SP_W[s3[b]]);
    for (int i = 0; i < 8; i++) {
        fp4 a = c[i], b = c[8 + i], s, d;
        sp_add(&s, &a, &b);
        sp_sub(&d, &a, &b);
        sp_mul(&c[i], &SP_16INV, &s);
        sp_mul(&c[8 + i], &SP_16INV_W24, &d);
    }
}

static void sp_init(void) {
    memcpy(SP_P.v, SRO_SP_P, 32);
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - SP_P.v[0] * inv;
    SP_NINV = (u64)0 - inv;
    fp4 x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) sp_add(&x, &x, &x);
    SP_R2 = x;
    for (int i = 0; i < 32; i++) { fp4 w; memcpy(w.v, SRO_SP_ROOTS[i], 32); sp_mul(&SP_W[i], &w, &SP_R2); }
    fp4 w;
    memcpy(w.v, SRO_SP_SIXTEEN_INV, 32); sp_mul(&SP_16INV, &w, &SP_R2);
    memcpy(w.v, SRO_SP_SIXTEEN_INV_W24, 32); sp_mul(&SP_16INV_W24, &w, &SP_R2);
}

/* ------------------------------------------------------------------ per-element API ---- */
static pthread_once_t once = PTHREAD_ONCE_INIT;
static void init_all(void) {
    f1_init(&GL, SRO_GL_P, SRO_GL_ROOTS, SRO_GL_KAPPA, SRO_GL_EIGHT_INV, SRO_GL_FOUR_INV);
    f1_init(&BB, SRO_BB_P, SRO_BB_ROOTS, SRO_BB_KAPPA, SRO_BB_EIGHT_INV, SRO_BB_FOUR_INV);
    sp_init();
}

size_t sro_elem_words(int ring) { return ring == SRO_GL ? 24 : ring == SRO_BB ? 72 : 64; }

static void crt1(int ring, u64* e) {
    switch (ring) {
    case SRO_GL: f1_crt_stages(&GL, e, 24); gl_homogenize(e); break;
    case SRO_BB: f1_crt_stages(&BB, e, 72); bb_homogenize(e); break;
    default: sp_crt((fp4*)e);
    }
}
static void icrt1(int ring, u64* e) {
    switch (ring) {
    case SRO_GL: gl_dehomogenize(e); f1_icrt_stages(&GL, e, 24); break;
    case SRO_BB: bb_dehomogenize(e); f1_icrt_stages(&BB, e, 72); break;
    default: sp_icrt((fp4*)e);
    }
}
/* ntt_form.rs:159-175: a *= b slot-wise */
static void nttmul1(int ring, u64* a, const u64* b) {
    switch (ring) {
    case SRO_GL: for (int s = 0; s < 8; s++) f3_mul(&GL, a + 3 * s, b + 3 * s); break;
    case SRO_BB: for (int s = 0; s < 8; s++) f9_mul(a + 9 * s, b + 9 * s); break;
    default: for (int s = 0; s < 16; s++) sp_mul((fp4*)a + s, (fp4*)a + s, (const fp4*)b + s);
    }
}
/* ntt_form.rs:588-601: a += b */
static void nttadd1(int ring, u64* a, const u64* b) {
    switch (ring) {
    case SRO_GL: for (int i = 0; i < 24; i++) a[i] = f1_add(&GL, a[i], b[i]); break;
    case SRO_BB: for (int i = 0; i < 72; i++) a[i] = f1_add(&BB, a[i], b[i]); break;
    default: for (int s = 0; s < 16; s++) sp_add((fp4*)a + s, (fp4*)a + s, (const fp4*)b + s);
    }
}

/* ------------------------------------------------------------------ batch API (threads) ---- */
typedef struct {
    int op, ring;
    u64 *a, *out;
    const u64* b;
    size_t lo, hi;
    const u64* const* rows; size_t m; /* matvec */
} job;

static void* worker(void* arg) {
    job* j = (job*)arg;
    size_t w = sro_elem_words(j->ring);
    u64 tmp[72];
    switch (j->op) {
    case 0: for (size_t i = j->lo; i < j->hi; i++) crt1(j->ring, j->a + i * w); break;
    case 1: for (size_t i = j->lo; i < j->hi; i++) icrt1(j->ring, j->a + i * w); break;
    case 2: for (size_t i = j->lo; i < j->hi; i++) nttmul1(j->ring, j->a + i * w, j->b + i * w); break;
    case 3: /* fused unit of the metric: icrt(crt(a) * crt(b)) */
        for (size_t i = j->lo; i < j->hi; i++) {
            u64* o = j->out + i * w;
            memcpy(tmp, j->b + i * w, w * 8);
            if (o != j->a + i * w) memcpy(o, j->a + i * w, w * 8);
            crt1(j->ring, o); crt1(j->ring, tmp); nttmul1(j->ring, o, tmp); icrt1(j->ring, o);
        }
        break;
    case 4: /* rows lo..hi of y = A v (matrix.rs:174: rayon parallelises over rows) */
        for (size_t r = j->lo; r < j->hi; r++) {
            u64* acc = j->out + r * w;
            memset(acc, 0, w * 8);
            for (size_t c = 0; c < j->m; c++) {
                memcpy(tmp, j->rows[r] + c * w, w * 8);
                nttmul1(j->ring, tmp, j->b + c * w);
                nttadd1(j->ring, acc, tmp);
            }
        }
        break;
    }
    return 0;
}

static void run(job base, size_t n, int threads) {
    pthread_once(&once, init_all);
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { base.lo = 0; base.hi = n; worker(&base); return; }
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    job* jobs = malloc(sizeof(job) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = base;
        jobs[t].lo = n * t / threads;
        jobs[t].hi = n * (t + 1) / threads;
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    free(th); free(jobs);
}

/* In-place batched conversions (crt.rs:10-25, 34-49).  n = number of ring elements. */
void sro_crt(int ring, u64* buf, size_t n, int threads) { job j = {0, ring, buf, 0, 0, 0, 0, 0, 0}; run(j, n, threads); }
void sro_icrt(int ring, u64* buf, size_t n, int threads) { job j = {1, ring, buf, 0, 0, 0, 0, 0, 0}; run(j, n, threads); }
void sro_ntt_mul(int ring, u64* a, const u64* b, size_t n, int threads) { job j = {2, ring, a, 0, b, 0, 0, 0, 0}; run(j, n, threads); }
void sro_ring_mul(int ring, const u64* a, const u64* b, u64* out, size_t n, int threads) {
    job j = {3, ring, (u64*)a, out, b, 0, 0, 0, 0}; run(j, n, threads);
}
/* y = A v; returns 1 (and writes nothing) when ncols != vlen (matrix.rs:169-171). */
int sro_matvec(int ring, const u64* const* rows, size_t kappa, size_t ncols, const u64* v, size_t vlen,
               u64* out, int threads) {
    if (ncols != vlen) return 1;
    job j = {4, ring, 0, out, v, 0, 0, rows, ncols};
    run(j, kappa, threads);
    return 0;
}

/* The callers' other linear maps.
 * Sparse mat-vec (sparse_matrix.rs:201-212) on the CSR image of coeffs: out[i] = sum_e vals[e] * v[col_idx[e]] for
 * e in [row_ptr[i], row_ptr[i+1]).  Returns 1 when ncols != vlen (None), 2 on a column index >= ncols (panic). */
int sro_sparse_matvec(int ring, size_t nrows, size_t ncols, const u64* row_ptr, const u64* col_idx, const u64* vals,
                      const u64* v, size_t vlen, u64* out) {
    pthread_once(&once, init_all);
    if (ncols != vlen) return 1;
    size_t w = sro_elem_words(ring);
    u64 tmp[72];
    for (size_t i = 0; i < nrows; i++) {
        u64* acc = out + i * w;
        memset(acc, 0, w * 8);
        for (u64 e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
            if (col_idx[e] >= ncols) return 2;
            memcpy(tmp, vals + e * w, w * 8);
            nttmul1(ring, tmp, v + col_idx[e] * w);
            nttadd1(ring, acc, tmp);
        }
    }
    return 0;
}
/* Dense mat-mat (matrix.rs:148-166): out[i][j] = sum_k a[i][k] * m[k][j]; returns 1 when a_ncols != m_nrows. */
int sro_matmat(int ring, const u64* const* a_rows, size_t a_nrows, size_t a_ncols, const u64* const* m_rows,
               size_t m_nrows, size_t m_ncols, u64* const* out_rows) {
    pthread_once(&once, init_all);
    if (a_ncols != m_nrows) return 1;
    size_t w = sro_elem_words(ring);
    u64 tmp[72];
    for (size_t i = 0; i < a_nrows; i++)
        for (size_t j = 0; j < m_ncols; j++) {
            u64* acc = out_rows[i] + j * w;
            memset(acc, 0, w * 8);
            for (size_t k = 0; k < a_ncols; k++) {
                memcpy(tmp, a_rows[i] + k * w, w * 8);
                nttmul1(ring, tmp, m_rows[k] + j * w);
                nttadd1(ring, acc, tmp);
            }
        }
    return 0;
}
/* MulAssign<&R> on a batch (matrix.rs:207-211, sparse_matrix.rs:298-302): a[e] *= r */
/* Element-wise ring addition / subtraction / negation over n elements (ntt_form.rs:588-626 and the same operators of
 * coeff_form.rs: field element by field element in either form); op 0 = add, 1 = sub, 2 = neg; in place on a. */
void sro_addsub(int ring, int op, u64* a, const u64* b, size_t n) {
    pthread_once(&once, init_all);
    size_t w = sro_elem_words(ring);
    if (ring == SRO_SP) {
        fp4 zero = {{0, 0, 0, 0}};
        for (size_t i = 0; i < n * 16; i++) {
            fp4* x = (fp4*)a + i;
            if (op == 0) sp_add(x, x, (const fp4*)b + i);
            else if (op == 1) sp_sub(x, x, (const fp4*)b + i);
            else sp_sub(x, &zero, x);
        }
        return;
    }
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    for (size_t i = 0; i < n * w; i++)
        a[i] = op == 0 ? f1_add(F, a[i], b[i]) : op == 1 ? f1_sub(F, a[i], b[i]) : f1_neg(F, a[i]);
}
/* Sum of n elements folded from ZERO (ntt_form.rs:640-654) */
void sro_sum(int ring, const u64* in, size_t n, u64* out) {
    pthread_once(&once, init_all);
    size_t w = sro_elem_words(ring);
    memset(out, 0, w * 8);
    for (size_t e = 0; e < n; e++) nttadd1(ring, out, in + e * w);
}

void sro_scale(int ring, u64* a, size_t n, const u64* r) {
    pthread_once(&once, init_all);
    size_t w = sro_elem_words(ring);
    for (size_t e = 0; e < n; e++) nttmul1(ring, a + e * w, r);
}

/* Canonical (de)serialization of n ring elements (coeff_form.rs:154-189, ntt_form.rs:24 through
 * ark-serialize 0.4, restated: each field element as the little-endian bytes of its standard-form integer, 8 / 4 / 32
 * bytes for Goldilocks / BabyBear / Starknet prime; no length prefix).  The reference holds no serialized vector:
 * parity unpinned. */
size_t sro_fe_bytes(int ring) { return ring == SRO_GL ? 8 : ring == SRO_BB ? 4 : 32; }
void sro_serialize(int ring, const u64* in, size_t n, unsigned char* out) {
    pthread_once(&once, init_all);
    size_t w = sro_elem_words(ring);
    if (ring == SRO_SP) {
        fp4 one = {{1, 0, 0, 0}};
        for (size_t i = 0; i < n * 16; i++) {
            fp4 r;
            sp_mul(&r, (const fp4*)in + i, &one); /* x R * 1 / R */
            memcpy(out + i * 32, r.v, 32);
        }
        return;
    }
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    size_t fb = sro_fe_bytes(ring);
    for (size_t i = 0; i < n * w; i++) {
        u64 x = f1_mul(F, in[i], 1);
        memcpy(out + i * fb, &x, fb); /* little-endian host */
    }
}
/* returns 1 (InvalidData) when an integer is not below the modulus */
int sro_deserialize(int ring, const unsigned char* in, size_t n, u64* out) {
    pthread_once(&once, init_all);
    size_t w = sro_elem_words(ring);
    if (ring == SRO_SP) {
        for (size_t i = 0; i < n * 16; i++) {
            fp4 x;
            memcpy(x.v, in + i * 32, 32);
            if (fp4_geq(&x, &SP_P)) return 1;
            sp_mul((fp4*)out + i, &x, &SP_R2);
        }
        return 0;
    }
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    size_t fb = sro_fe_bytes(ring);
    for (size_t i = 0; i < n * w; i++) {
        u64 x = 0;
        memcpy(&x, in + i * fb, fb);
        if (x >= F->p) return 1;
        out[i] = f1_mul(F, x, F->r2);
    }
    return 0;
}

/* Coefficient-form helpers.  reduce: n polynomials of len field elements (D <= len <= 2D) -> n elements
 * (goldilocks/mod.rs:75-98, babybear/mod.rs:87-110, stark_prime/mod.rs:40-47). */
void sro_reduce(int ring, const u64* in, size_t n, size_t len, u64* out) {
    pthread_once(&once, init_all);
    if (ring == SRO_SP) {
        for (size_t e = 0; e < n; e++)
            for (size_t i = 0; i < 16; i++) {
                fp4 r, z = {{0, 0, 0, 0}};
                memcpy(r.v, in + (e * len + i) * 4, 32);
                if (16 + i < len) memcpy(z.v, in + (e * len + 16 + i) * 4, 32);
                sp_sub(&r, &r, &z);
                memcpy(out + (e * 16 + i) * 4, r.v, 32);
            }
        return;
    }
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    const size_t D = ring == SRO_GL ? 24 : 72, H = D / 2;
    for (size_t e = 0; e < n; e++) {
        const u64* c = in + e * len;
        for (size_t i = 0; i < H; i++) {
            u64 r = c[i];
            if (D + i < len) r = f1_sub(F, r, c[D + i]);
            if (D + H + i < len) r = f1_sub(F, r, c[D + H + i]);
            out[e * D + i] = r;
        }
        for (size_t i = H; i < D; i++) out[e * D + i] = (H + i < len) ? f1_add(F, c[i], c[H + i]) : c[i];
    }
}
/* rot: multiplication by X (goldilocks/mod.rs:138-149, babybear/mod.rs:150-161, stark_prime/mod.rs:87-95) */
void sro_rot(int ring, const u64* in, size_t n, u64* out) {
    pthread_once(&once, init_all);
    if (ring == SRO_SP) {
        for (size_t e = 0; e < n; e++) {
            fp4 last, z = {{0, 0, 0, 0}}, neg;
            memcpy(last.v, in + (e * 16 + 15) * 4, 32);
            sp_sub(&neg, &z, &last);
            memcpy(out + e * 64, neg.v, 32);
            memcpy(out + e * 64 + 4, in + e * 64, 15 * 32);
        }
        return;
    }
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    const size_t D = ring == SRO_GL ? 24 : 72;
    for (size_t e = 0; e < n; e++) {
        const u64 last = in[e * D + D - 1];
        out[e * D] = f1_neg(F, last);
        for (size_t i = 1; i < D; i++) out[e * D + i] = in[e * D + i - 1];
        out[e * D + D / 2] = f1_add(F, out[e * D + D / 2], last);
    }
}

/* Balanced gadget decomposition (balanced_decomposition/mod.rs:62-103,163-175, coeff_form.rs:588-606),
 * Fp64 rings only.  in: n elements; out: n * pad elements.  Returns 1 if a decomposition does not fit in pad digits
 * (the reference panics), 2 for an unsupported ring / basis. */
int sro_gadget_decompose(int ring, const u64* in, size_t n, u64 b, size_t pad, u64* out) {
    pthread_once(&once, init_all);
    if (ring == SRO_SP || b < 2 || (b & 1) || b >= (1ull << 62)) return 2;
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    const size_t D = ring == SRO_GL ? 24 : 72;
    const long long B = (long long)b, BH = B / 2;
    int overflow = 0;
    for (size_t j = 0; j < n; j++)
        for (size_t i = 0; i < D; i++) {
            u64 x = f1_mul(F, in[j * D + i], 1);  /* out of Montgomery form */
            long long cur = x > (F->p - 1) / 2 ? (long long)(x - F->p) : (long long)x;
            size_t t = 0;
            for (; t < pad; t++) {
                long long rem = cur % B, q = cur / B, digit;
                if ((rem < 0 ? -rem : rem) <= BH) { digit = rem; cur = q; }
                else { digit = rem < 0 ? rem + B : rem - B; cur = q + (rem < 0 ? -1 : 1); }
                u64 mag = (u64)(digit < 0 ? -digit : digit) % F->p;
                u64 dstd = (digit < 0 && mag) ? F->p - mag : mag;
                out[((j * pad + t) * D) + i] = f1_mul(F, dstd, F->r2);
                if (cur == 0) { t++; break; }
            }
            if (cur != 0) overflow = 1;
            for (; t < pad; t++) out[((j * pad + t) * D) + i] = 0;
        }
    return overflow;
}
/* GadgetRecompose (mod.rs:177-190): in holds n * pad digit elements, out n elements. */
int sro_gadget_recompose(int ring, const u64* in, size_t n, u64 b, size_t pad, u64* out) {
    pthread_once(&once, init_all);
    if (ring == SRO_SP) return 2;
    const f1_ctx* F = ring == SRO_GL ? &GL : &BB;
    const size_t D = ring == SRO_GL ? 24 : 72;
    const u64 bm = f1_mul(F, b % F->p, F->r2);
    for (size_t j = 0; j < n; j++)
        for (size_t i = 0; i < D; i++) {
            u64 acc = 0;
            for (size_t t = pad; t-- > 0;) acc = f1_add(F, f1_mul(F, acc, bm), in[((j * pad + t) * D) + i]);
            out[j * D + i] = acc;
        }
    return 0;
}

/* stage-only variants for tests (crt without homogenize) */
void sro_crt_stages(int ring, u64* e) {
    pthread_once(&once, init_all);
    if (ring == SRO_GL) f1_crt_stages(&GL, e, 24); else if (ring == SRO_BB) f1_crt_stages(&BB, e, 72); else sp_crt((fp4*)e);
}
