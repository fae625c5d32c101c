// Starknet-prime ring with TWO threads per ring element.
//
// One polynomial is 16 x 8 = 128 registers, so the thread-per-element kernels sit at 255 registers and
// 6 warps per SM.  Here thread h (h = lane / 16, partner = lane ^ 16) holds the eight coefficients of
// parity h, c[j] = coefficient h + 2 j.  The butterfly stages of span 8, 4 and 2 pair coefficients of
// equal parity, so they are thread-local; only the span-1 stage pairs (2k, 2k+1) across the two threads.
// That stage is balanced by exchanging half of the operands: thread 0 finishes pairs k = 0..3, thread 1
// pairs k = 4..7, after which thread h holds the contiguous positions 8h .. 8h+7 of the CRT form.
// The inverse transform runs the mirror image (span-1 stage local on positions 8h + (2j, 2j+1), one
// exchange back to the parity layout, then three local stages).
//
// Reference: stark_prime/ntt.rs:121-235 (CRT) and :245-346 (ICRT); same schedule, same constants.
#pragma once
#include "sp_ring.cuh"

namespace sr {
namespace sp {

// twiddle indices of the span-1 stages, by pair k (ntt.rs:194-234 and :248-280)
constexpr int CRT_LAST_K(int k) {
    constexpr int t[8] = {1, 9, 5, 13, 3, 11, 7, 15};
    return t[k];
}
constexpr int ICRT_FIRST_K(int k) {
    constexpr int t[8] = {31, 23, 27, 19, 29, 21, 25, 17};
    return t[k];
}

struct RootTableRT {
    u32 w[32][8];
};
#if defined(__CUDACC__)
__constant__ RootTableRT SP_WTAB = {SR_SP_ROOTS_MONT};
#endif
SR_HD void load_root(Fe& w, int k) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < L; i++) w.v[i] = SP_WTAB.w[k][i];
#else
    for (int i = 0; i < L; i++) w.v[i] = ROOTS_MONT.w[k][i];
#endif
}

template <int K>
SR_HD void bfly2(Fe& a, Fe& b) {  // (a, b) <- (a + w b, a - w b)
    Fe t, x = a;
    mulw<K>(t, b);
    add(a, x, t);
    sub(b, x, t);
}
template <int K>
SR_HD void ibfly2(Fe& a, Fe& b) {  // (a, b) <- (a + b, w (a - b))
    Fe x = a, y = b, d;
    add(a, x, y);
    sub(d, x, y);
    mulw<K>(b, d);
}

// c[j] = coefficient h + 2j.  Stages of span 8, 4, 2 (ntt.rs:124-192), identical for both parities.
SR_HD void half_crt_local(Fe (&c)[8]) {
    bfly2<8>(c[0], c[4]); bfly2<8>(c[1], c[5]); bfly2<8>(c[2], c[6]); bfly2<8>(c[3], c[7]);
    bfly2<4>(c[0], c[2]); bfly2<4>(c[1], c[3]); bfly2<12>(c[4], c[6]); bfly2<12>(c[5], c[7]);
    bfly2<2>(c[0], c[1]); bfly2<10>(c[2], c[3]); bfly2<6>(c[4], c[5]); bfly2<14>(c[6], c[7]);
}
// what this thread hands to its partner before the span-1 stage: thread 0 gives a_4..a_7, thread 1 b_0..b_3
SR_HD void half_crt_send(Fe (&send)[4], const Fe (&c)[8], int h) {
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < L; i++) send[j].v[i] = h ? c[j].v[i] : c[4 + j].v[i];
}
// span-1 stage on pairs k = 4h + j: out[2j], out[2j+1] = positions 8h + 2j, 8h + 2j + 1 of the CRT form
SR_HD void half_crt_cross(Fe (&out)[8], const Fe (&c)[8], const Fe (&recv)[4], int h) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        Fe a, b, w, t;
#pragma unroll
        for (int i = 0; i < L; i++) {
            a.v[i] = h ? recv[j].v[i] : c[j].v[i];
            b.v[i] = h ? c[4 + j].v[i] : recv[j].v[i];
        }
        load_root(w, h ? CRT_LAST_K(4 + j) : CRT_LAST_K(j));
        mont_mul(t, b, w);
        add(out[2 * j], a, t);
        sub(out[2 * j + 1], a, t);
    }
}

// Inverse: p[q] = position 8h + q.  Span-1 stage (ntt.rs:248-280), local.
SR_HD void half_icrt_first(Fe (&p)[8], int h) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        Fe a = p[2 * j], b = p[2 * j + 1], d, w;
        add(p[2 * j], a, b);
        sub(d, a, b);
        load_root(w, h ? ICRT_FIRST_K(4 + j) : ICRT_FIRST_K(j));
        mont_mul(p[2 * j + 1], d, w);
    }
}
// back to the parity layout: thread 0 keeps its even positions and hands over the odd ones, thread 1
// keeps its odd positions and hands over the even ones
SR_HD void half_icrt_send(Fe (&send)[4], const Fe (&p)[8], int h) {
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < L; i++) send[j].v[i] = h ? p[2 * j].v[i] : p[2 * j + 1].v[i];
}
SR_HD void half_icrt_gather(Fe (&c)[8], const Fe (&p)[8], const Fe (&recv)[4], int h) {
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < L; i++) {
            const u32 keep = h ? p[2 * j + 1].v[i] : p[2 * j].v[i];
            c[j].v[i] = h ? recv[j].v[i] : keep;
            c[4 + j].v[i] = h ? keep : recv[j].v[i];
        }
}
// remaining stages (spans 2, 4, 8 with the final 1/16 scalings, ntt.rs:282-345), thread-local
SR_HD void half_icrt_local(Fe (&c)[8]) {
    ibfly2<30>(c[0], c[1]); ibfly2<22>(c[2], c[3]); ibfly2<26>(c[4], c[5]); ibfly2<18>(c[6], c[7]);
    ibfly2<28>(c[0], c[2]); ibfly2<28>(c[1], c[3]); ibfly2<20>(c[4], c[6]); ibfly2<20>(c[5], c[7]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        Fe a = c[j], b = c[4 + j], s, d;
        add(s, a, b);
        sub(d, a, b);
        mul_scale<0>(c[j], s);
        mul_scale<1>(c[4 + j], d);
    }
}

}  // namespace sp
}  // namespace sr
