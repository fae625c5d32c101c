// BabyBear row policy: how a ring element sits in a shared-memory row and the per-row operators.
// HBM layout: 72 u64 limbs (576 B), each < 2^31.  Shared-memory row: the 72 low words (288 B) + 16 B pad.
#pragma once
#include "bb_ring.cuh"
#include "sr_tile.cuh"

namespace sr {

struct BBPolicy {
    static constexpr int RING = RING_BB;
    static constexpr int WORDS64 = 72;
    static constexpr int CHUNKS = 36;  // 16-byte chunks per element in HBM
    static constexpr int STAGE_UNROLL = 18;  // loads in flight per thread while staging
    static constexpr int ROW = 76;     // words per shared-memory row
    static constexpr int NREG = 72;

    SR_D static void put(u32* row, int j, uint4 v) {  // two limbs -> two words
        *reinterpret_cast<uint2*>(row + 2 * j) = make_uint2(v.x, v.z);
    }
    SR_D static uint4 get(const u32* row, int j) {
        uint2 t = *reinterpret_cast<const uint2*>(row + 2 * j);
        return make_uint4(t.x, 0u, t.y, 0u);
    }

    // words [9 S, 9 S + 9) of a row through aligned 128-bit loads
    template <int S>
    SR_D static void load_slot(u32 (&x)[9], const u32* row) {
        constexpr int lo = (9 * S) / 4 * 4, hi = (9 * S + 9 + 3) / 4 * 4;
        u32 t[hi - lo];
#pragma unroll
        for (int i = 0; i < (hi - lo) / 4; i++) {
            uint4 v = *reinterpret_cast<const uint4*>(row + lo + 4 * i);
            t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 9; j++) x[j] = t[9 * S - lo + j];
    }

    SR_D static void op_crt(u32* rowA) {
        u32 c[72];
        row_load(c, rowA);
        bb::crt(c);
        row_store(rowA, c);
    }
    SR_D static void op_icrt(u32* rowA) {
        u32 c[72];
        row_load(c, rowA);
        bb::icrt(c);
        row_store(rowA, c);
    }
    // rowA <- rowA * rowB (NTT form, slot-wise)
    SR_D static void op_ntt_mul(u32* rowA, const u32* rowB) {
        u32 a[72], b[72];
        row_load(a, rowA);
        row_load(b, rowB);
        bb::ntt_mul(a, b);
        row_store(rowA, a);
    }

    template <int S>
    SR_D static void fused_slot(u32 (&bs)[72], const u32* rowA) {
        constexpr int KS[8] = {1, 13, 7, 19, 5, 17, 11, 23};
        u32 x[9], y[9], z[9];
        load_slot<S>(x, rowA);
#pragma unroll
        for (int j = 0; j < 9; j++) y[j] = bs[9 * S + j];
        bb::slot_mul_pow<bb::w_std(KS[S])>(z, x, y);
#pragma unroll
        for (int j = 0; j < 9; j++) bs[9 * S + j] = z[j];
    }
    // rowA <- icrt(crt(rowA) * crt(rowB)); crt(a) is parked in the thread's own row meanwhile
    SR_D static void op_ring_mul(u32* rowA, const u32* rowB) {
        u32 c[72];
        row_load(c, rowA);
        bb::crt_stages(c);
        row_store(rowA, c);
        row_load(c, rowB);
        bb::crt_stages(c);
        fused_slot<0>(c, rowA);
        fused_slot<1>(c, rowA);
        fused_slot<2>(c, rowA);
        fused_slot<3>(c, rowA);
        fused_slot<4>(c, rowA);
        fused_slot<5>(c, rowA);
        fused_slot<6>(c, rowA);
        fused_slot<7>(c, rowA);
        bb::icrt_stages<bb::R32_INV>(c);
        row_store(rowA, c);
    }
};

}  // namespace sr
