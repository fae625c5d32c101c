// Starknet-prime row policy.  HBM layout: 16 x 4 u64 limbs (512 B).  Shared-memory row: the same
// 128 words + 16 B pad.  One polynomial fills 128 registers, so the fused product parks crt(a) in
// the thread's own row and streams it back one coefficient at a time.
#pragma once
#include "sp_ring.cuh"
#include "sr_tile.cuh"

namespace sr {

struct SPPolicy {
    static constexpr int RING = RING_SP;
    static constexpr int WORDS64 = 64;
    static constexpr int CHUNKS = 32;
    static constexpr int STAGE_UNROLL = 16;  // loads in flight per thread while staging
    static constexpr int ROW = 132;

    SR_D static void put(u32* row, int j, uint4 v) { *reinterpret_cast<uint4*>(row + 4 * j) = v; }
    SR_D static uint4 get(const u32* row, int j) { return *reinterpret_cast<const uint4*>(row + 4 * j); }

    SR_D static void load_fe(sp::Fe& f, const u32* row, int i) {
        uint4 lo = *reinterpret_cast<const uint4*>(row + 8 * i);
        uint4 hi = *reinterpret_cast<const uint4*>(row + 8 * i + 4);
        f.v[0] = lo.x; f.v[1] = lo.y; f.v[2] = lo.z; f.v[3] = lo.w;
        f.v[4] = hi.x; f.v[5] = hi.y; f.v[6] = hi.z; f.v[7] = hi.w;
    }
    SR_D static void store_fe(u32* row, int i, const sp::Fe& f) {
        *reinterpret_cast<uint4*>(row + 8 * i) = make_uint4(f.v[0], f.v[1], f.v[2], f.v[3]);
        *reinterpret_cast<uint4*>(row + 8 * i + 4) = make_uint4(f.v[4], f.v[5], f.v[6], f.v[7]);
    }
    SR_D static void load(sp::Fe (&c)[16], const u32* row) {
#pragma unroll
        for (int i = 0; i < 16; i++) load_fe(c[i], row, i);
    }
    SR_D static void store(u32* row, const sp::Fe (&c)[16]) {
#pragma unroll
        for (int i = 0; i < 16; i++) store_fe(row, i, c[i]);
    }

    SR_D static void op_crt(u32* rowA) {
        sp::Fe c[16];
        load(c, rowA);
        sp::crt(c);
        store(rowA, c);
    }
    SR_D static void op_icrt(u32* rowA) {
        sp::Fe c[16];
        load(c, rowA);
        sp::icrt(c);
        store(rowA, c);
    }
    SR_D static void op_ntt_mul(u32* rowA, const u32* rowB) {
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            sp::Fe x, y, z;
            load_fe(x, rowA, i);
            load_fe(y, rowB, i);
            sp::mont_mul(z, x, y);
            store_fe(rowA, i, z);
        }
    }
    SR_D static void op_ring_mul(u32* rowA, const u32* rowB) {
        sp::Fe c[16];
        load(c, rowA);
        sp::crt(c);
        store(rowA, c);
        load(c, rowB);
        sp::crt(c);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            sp::Fe x;
            load_fe(x, rowA, i);
            sp::mont_mul(c[i], c[i], x);
        }
        sp::icrt(c);
        store(rowA, c);
    }
};

}  // namespace sr
