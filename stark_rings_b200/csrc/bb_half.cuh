// BabyBear fused ring multiplication with TWO threads per ring element.
//
// The thread-per-element mapping needs ~255 registers for the fused CRT -> slot product -> ICRT
// and leaves 8 warps per SM, which is too few to hide the integer-pipe latencies.  Here thread h
// of an element (h = lane / 16, partner = lane ^ 16) owns the four CRT slots 4h..4h+3, i.e.
// coefficients [36 h, 36 h + 36) once the first (zeta) stage is done:
//   stage 1   out_i = a_i + c_h * a_{i+36}          c_0 = zeta, c_1 = zeta^5 = 1 - zeta
//             (both threads read the whole row from shared memory; each computes its own half)
//   stages 2,3, slot products, inverse stages 1,2    thread-local (36 values)
//   last inverse stage  out = cx_h * own + cy_h * partner     (values exchanged by warp shuffle)
//             h = 0: ((1 - kappa)/8, (1 + kappa)/8)   h = 1: (-kappa/4, kappa/4)
// All twiddles are per-lane runtime constants, so both halves run the same instruction stream
// (no divergence) and the code is half as long.  Same algebra as bb_ring.cuh, same results.
//
// Reference: babybear/ntt.rs:143-236 (CRT), :238-317 (ICRT); the slot isomorphisms (:324-588) are
// skipped in the fused product (see bb::fused_mul_icrt).
#pragma once
#include "bb_ring.cuh"

namespace sr {
namespace bb {

// Twiddles in standard form with their Shoup companions (bb::Tw); only the last stage's coefficients, which also carry
// the Montgomery 2^-32 of the slot products, stay in Montgomery operand form.
struct HalfConsts {
    Tw c1;         // stage 1 multiplier
    Tw w2;         // stage 2 twiddle
    Tw w3[2];      // stage 3 twiddles (first / second 18-block of the half)
    Tw rho[4];     // slot moduli r^k_s
    Tw iw1[2];     // inverse stage 1 twiddles
    Tw iw2;        // inverse stage 2 twiddle
    u32 cx, cy;    // last stage: own / partner coefficients (include the Montgomery 2^-32), operand form (x 2^32)
};

constexpr u32 csub(u32 a, u32 b) { return a >= b ? a - b : a + P - b; }
constexpr u32 cadd(u32 a, u32 b) { return (u32)(((u64)a + b) % P); }
constexpr Tw tw(int k) { return shoup(w_std(k)); }

constexpr HalfConsts half_consts(int h) {
    constexpr u32 kappa = SR_BB_KAPPA_STD, e8 = SR_BB_EIGHT_INV_STD, e4 = SR_BB_FOUR_INV_STD;
    // extra 2^-32: second half of the Montgomery-64 factor of the slot product
    const u32 s8 = cmulmod(e8, R32_INV), s4 = cmulmod(e4, R32_INV);
    if (h == 0)
        return HalfConsts{tw(4), tw(2), {tw(1), tw(7)},
                          {tw(1), tw(13), tw(7), tw(19)},
                          {tw(23), tw(17)}, tw(22),
                          to_m32(cmulmod(csub(1, kappa), s8)), to_m32(cmulmod(cadd(1, kappa), s8))};
    return HalfConsts{shoup(csub(1, w_std(4))), tw(10), {tw(5), tw(11)},
                      {tw(5), tw(17), tw(11), tw(23)},
                      {tw(19), tw(13)}, tw(14),
                      to_m32(cmulmod(csub(0, kappa), s4)), to_m32(cmulmod(kappa, s4))};
}

// x <- this thread's half of crt_stages(row); `row` holds all 72 coefficients.
SR_HD void half_crt(u32 (&x)[36], const u32* row, const HalfConsts& K) {
#pragma unroll
    for (int j = 0; j < 9; j++) {
#if defined(__CUDA_ARCH__)
        uint4 lo = *reinterpret_cast<const uint4*>(row + 4 * j);
        uint4 hi = *reinterpret_cast<const uint4*>(row + 36 + 4 * j);
        u32 a[4] = {lo.x, lo.y, lo.z, lo.w}, b[4] = {hi.x, hi.y, hi.z, hi.w};
#else
        u32 a[4], b[4];
        for (int t = 0; t < 4; t++) { a[t] = row[4 * j + t]; b[t] = row[36 + 4 * j + t]; }
#endif
#pragma unroll
        for (int t = 0; t < 4; t++) {
            x[4 * j + t] = add(a[t], muls(b[t], K.c1));
        }
    }
#pragma unroll
    for (int i = 0; i < 18; i++) {
        u32 a = x[i], t = muls(x[18 + i], K.w2);
        x[i] = add(a, t);
        x[18 + i] = sub(a, t);
    }
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int i = 0; i < 9; i++) {
            u32 a = x[18 * q + i], t = muls(x[18 * q + 9 + i], K.w3[q]);
            x[18 * q + i] = add(a, t);
            x[18 * q + 9 + i] = sub(a, t);
        }
}

// Slot product with a runtime modulus (operand form); see slot_mul_pow.
SR_HD void slot_mul_rt(u32* z, const u32* x, const u32* y, const Tw& rho) {
    u32 yr[SLOT], xv[SLOT], yv[SLOT];
#pragma unroll
    for (int j = 0; j < SLOT; j++) {
        xv[j] = x[j];
        yv[j] = y[j];
    }
#pragma unroll
    for (int j = 1; j < SLOT; j++) yr[j] = muls(yv[j], rho);
#pragma unroll
    for (int k = 0; k < SLOT; k++) {
        u64 acc = 0;
#pragma unroll
        for (int i = 0; i < SLOT; i++) {
            u32 f = (i <= k) ? yv[(i <= k) ? k - i : 0] : yr[(i <= k) ? 1 : k + SLOT - i];
            acc += (u64)xv[i] * f;
            if (i == 3 || i == 5 || i == 7 || i == 8) {
                u32 hi = (u32)(acc >> 32);
                hi = umin32(hi, hi - P);
                acc = ((u64)hi << 32) | (u32)acc;
            }
        }
        z[k] = red(acc);
    }
}

// b <- a * b slot-wise on this thread's four slots (natural order, moduli K.rho)
SR_HD void half_slots(u32 (&b)[36], const u32 (&a)[36], const HalfConsts& K) {
#pragma unroll
    for (int s = 0; s < 4; s++) slot_mul_rt(&b[9 * s], &a[9 * s], &b[9 * s], K.rho[s]);
}

// the two thread-local inverse stages
SR_HD void half_icrt_local(u32 (&x)[36], const HalfConsts& K) {
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int i = 0; i < 9; i++) {
            u32 a = x[18 * q + i], b = x[18 * q + 9 + i];
            x[18 * q + i] = add(a, b);
            x[18 * q + 9 + i] = muls(a - b + P, K.iw1[q]);
        }
#pragma unroll
    for (int i = 0; i < 18; i++) {
        u32 a = x[i], b = x[18 + i];
        x[i] = add(a, b);
        x[18 + i] = muls(a - b + P, K.iw2);
    }
}

// last inverse stage: own half of the coefficient-form product from own (x) and partner (y) values
SR_HD void half_final(u32 (&out)[36], const u32 (&x)[36], const u32 (&y)[36], const HalfConsts& K) {
#pragma unroll
    for (int i = 0; i < 36; i++) out[i] = red((u64)x[i] * K.cx + (u64)y[i] * K.cy);
}

}  // namespace bb
}  // namespace sr
