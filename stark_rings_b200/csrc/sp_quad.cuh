// Starknet-prime ring with FOUR threads per ring element, one butterfly stage at a time over the element's
// shared-memory row.
//
// Why (profiles/r01b_gl_ncu.md, Starknet section): two-thread kernels (round 1, removed) hold eight 256-bit values per
// thread, need 168 registers and keep 12 warps per SM; `wait` on the carry chains of the wide multiply-adds is the
// dominant stall, fewer warps are slower (8 warps: -25%), and more warps are impossible both by registers and by
// shared memory (two 528-byte rows per element at two threads per element = 14 warps).  Here an element belongs to
// four threads of one warp (thread t of the warp's eight elements sits in lanes 8t .. 8t + 7, so that a quarter-warp
// touches the same coefficient of eight consecutive rows: conflict-free with the padded rows); in every stage thread t
// owns butterflies 2t and 2t + 1, reads its four operands from the
// row, writes them back, and a __syncwarp() separates the stages.  A thread never holds more than four values, the
// stage body exists once (the stage, the operand and the twiddle are runtime values), so the kernel needs about half
// the registers and a sixth of the code of the two-thread version and 24 warps fit on an SM.
//
// Same schedule and constants as sp::crt / sp::icrt (reference stark_prime/ntt.rs:121-235, :245-346):
// stage s of the forward transform has span S = 8 >> s and uses, for block b (pairs (i, i + S), i = 2 S b + o),
// the root W[2^(3-s) + 2^(4-s) rev_s(b)]: 8 | 4, 12 | 2, 10, 6, 14 | 1, 9, 5, 13, 3, 11, 7, 15; the inverse uses
// W[32 - that] with the stages in the opposite order and ends with the two scaled outputs (1/16, W[24]/16).
// All functions are __host__ __device__; tests/hostcheck runs the four "threads" of an element in turn.
//
// LAZY: values stay UNREDUCED below 2^256 = 31.99 p (sp_ring.cuh: add_nr / sub_kp / mont_mul_nr) until they are stored as
// results.  Used by the fused ring product (bounds below), by the stand-alone CRT (the same forward bounds; the sixteen
// outputs < 9p are canonicalised in one step each, canon_small) and by the stand-alone ICRT (canonical inputs: every
// bound of the inverse stages below holds with room to spare).  Bounds, in multiples of p:
//   forward stage s = 0..3   in < 2s + 1   t = w b < 2   out = a + t, a - t + 2p < 2s + 3        (9 after the last)
//   slot product             x y 2^-256 < 81 / 31.99 + 1 < 3.54
//   inverse stage 0 (K = 4)  a + b < 7.08,  a - b + 4p < 7.54 -> times w < 2
//   inverse stage 1 (K = 8)  a + b < 14.16, a - b + 8p < 15.08 -> < 2
//   inverse stage 2 (K = 16) a + b < 28.32, a - b + 16p < 30.16 -> < 2
//   last stage               both inputs partially reduced below 4, a + b, a - b + 4p < 8, scaled and canonicalised
// Per ring product this removes 104 of the 120 conditional subtractions after a Montgomery multiplication and turns
// the 128 modular additions / subtractions (8 + 9 + 8 selects each) into 8 / 16 plain carry-chain instructions.
#pragma once
#include "sp_ring.cuh"

namespace sr {
namespace sp {

SR_HD void quad_ld(Fe& f, const u32* row, int i) {
#if defined(__CUDA_ARCH__)
    const uint4 lo = *reinterpret_cast<const uint4*>(row + 8 * i), hi = *reinterpret_cast<const uint4*>(row + 8 * i + 4);
    f.v[0] = lo.x; f.v[1] = lo.y; f.v[2] = lo.z; f.v[3] = lo.w;
    f.v[4] = hi.x; f.v[5] = hi.y; f.v[6] = hi.z; f.v[7] = hi.w;
#else
    for (int k = 0; k < L; k++) f.v[k] = row[8 * i + k];
#endif
}
SR_HD void quad_st(u32* row, int i, const Fe& f) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(row + 8 * i) = make_uint4(f.v[0], f.v[1], f.v[2], f.v[3]);
    *reinterpret_cast<uint4*>(row + 8 * i + 4) = make_uint4(f.v[4], f.v[5], f.v[6], f.v[7]);
#else
    for (int k = 0; k < L; k++) row[8 * i + k] = f.v[k];
#endif
}
// reversal of the low s bits of b (s <= 3)
SR_HD int rev_bits(int b, int s) {
    const int r3 = ((b & 1) << 2) | (b & 2) | ((b >> 2) & 1);
    return r3 >> (3 - s);
}
// index into ROOTS_OF_UNITY_32 of forward stage s (span 8 >> s), block b
SR_HD int fwd_root(int s, int b) { return (8 >> s) + (16 >> s) * rev_bits(b, s); }

// wtab: the 32 roots in Montgomery form, 8 limbs each (shared memory on the device)
SR_HD void quad_root(Fe& w, const u32* wtab, int k) { quad_ld(w, wtab, k); }

// forward stage s (0..3) of the element in `row`, thread t (0..3): butterflies 2t, 2t + 1
// CANON_LAST (with LAZY): the outputs of the last stage (< 9p) are canonicalised in one step before they are stored
// (the stand-alone CRT, whose results leave the kernel)
template <bool LAZY = false, bool CANON_LAST = false>
SR_HD void quad_fwd_stage(u32* row, const u32* wtab, int s, int t) {
    const int S = 8 >> s;
    Fe a[2], b[2], w[2], m[2];
    int idx[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int bf = 2 * t + u, blk = bf >> (3 - s), o = bf & (S - 1);
        idx[u] = (blk << (4 - s)) + o;
        quad_ld(a[u], row, idx[u]);
        quad_ld(b[u], row, idx[u] + S);
        quad_root(w[u], wtab, fwd_root(s, blk));
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        if (LAZY) mont_mul_nr(m[u], b[u], w[u]);
        else mont_mul(m[u], b[u], w[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        Fe x, y;
        if (LAZY) {
            add_nr(x, a[u], m[u]);
            sub_kp(y, a[u], m[u], 2);
            if (CANON_LAST && s == 3) {
                canon_small(x, x);
                canon_small(y, y);
            }
        } else {
            add(x, a[u], m[u]);
            sub(y, a[u], m[u]);
        }
        quad_st(row, idx[u], x);
        quad_st(row, idx[u] + S, y);
    }
}
// inverse stage s (0..2: span 1 << s): (a, b) <- (a + b, w (a - b)) with w = W[32 - forward root of that span]
template <bool LAZY = false>
SR_HD void quad_inv_stage(u32* row, const u32* wtab, int s, int t) {
    const int S = 1 << s;
    Fe a[2], b[2], w[2], d[2];
    int idx[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int bf = 2 * t + u, blk = bf >> s, o = bf & (S - 1);
        idx[u] = (blk << (s + 1)) + o;
        quad_ld(a[u], row, idx[u]);
        quad_ld(b[u], row, idx[u] + S);
        quad_root(w[u], wtab, 32 - fwd_root(3 - s, blk));
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        Fe x;
        if (LAZY) {
            add_nr(x, a[u], b[u]);
            sub_kp(d[u], a[u], b[u], 4u << s);
        } else {
            add(x, a[u], b[u]);
            sub(d[u], a[u], b[u]);
        }
        quad_st(row, idx[u], x);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        Fe y;
        if (LAZY) mont_mul_nr(y, d[u], w[u]);
        else mont_mul(y, d[u], w[u]);
        quad_st(row, idx[u] + S, y);
    }
}
// last inverse stage (span 8) with the scalings 1/16 and W[24]/16 (ntt.rs:332-345); output canonical
// SMALL_IN (with LAZY): the inputs are below 8p (the stand-alone ICRT, whose inverse stages start from canonical values:
// a + b < 2, 4, 8 after stages 0, 1, 2), so that a + b and a - b + 8p stay below 16p without a partial reduction
template <bool LAZY = false, bool SMALL_IN = false>
SR_HD void quad_inv_last(u32* row, int t) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int i = 2 * t + u;
        Fe a, b, s, d, x, y;
        quad_ld(a, row, i);
        quad_ld(b, row, i + 8);
        if (LAZY && SMALL_IN) {
            add_nr(s, a, b);
            sub_kp(d, a, b, 8);
        } else if (LAZY) {
            Fe ar, br;
            partial_reduce(ar, a);
            partial_reduce(br, b);
            add_nr(s, ar, br);
            sub_kp(d, ar, br, 4);
        } else {
            add(s, a, b);
            sub(d, a, b);
        }
        mul_scale<0>(x, s);
        mul_scale<1>(y, d);
        quad_st(row, i, x);
        quad_st(row, i + 8, y);
    }
}
// slot products 4t .. 4t + 3: rowA[k] <- rowA[k] * rowB[k] (ntt_form.rs:159-175 with BaseCRTField = Fq)
template <bool LAZY = false>
SR_HD void quad_slots(u32* rowA, const u32* rowB, int t) {
#pragma unroll 2
    for (int u = 0; u < 4; u++) {
        Fe x, y, z;
        quad_ld(x, rowA, 4 * t + u);
        quad_ld(y, rowB, 4 * t + u);
        if (LAZY) mont_mul_nr(z, x, y);
        else mont_mul(z, x, y);
        quad_st(rowA, 4 * t + u, z);
    }
}

}  // namespace sp
}  // namespace sr
