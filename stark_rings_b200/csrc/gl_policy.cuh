// Goldilocks row policy.  HBM layout: 24 u64 limbs (192 B).  Shared-memory row: the same 48 words + 16 B pad.
#pragma once
#include "gl_ring.cuh"
#include "gl_fused6.cuh"
#include "sr_tile.cuh"

namespace sr {

struct GLPolicy {
    static constexpr int RING = RING_GL;
    static constexpr int WORDS64 = 24;
    static constexpr int CHUNKS = 12;
    static constexpr int STAGE_UNROLL = 12;  // loads in flight per thread while staging
    static constexpr int ROW = 52;

    SR_D static void put(u32* row, int j, uint4 v) { *reinterpret_cast<uint4*>(row + 4 * j) = v; }
    SR_D static uint4 get(const u32* row, int j) { return *reinterpret_cast<const uint4*>(row + 4 * j); }

    SR_D static void load(u64 (&c)[24], const u32* row) {
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint4 v = *reinterpret_cast<const uint4*>(row + 4 * i);
            c[2 * i] = (u64)v.x | ((u64)v.y << 32);
            c[2 * i + 1] = (u64)v.z | ((u64)v.w << 32);
        }
    }
    SR_D static void store(u32* row, const u64 (&c)[24]) {
#pragma unroll
        for (int i = 0; i < 12; i++)
            *reinterpret_cast<uint4*>(row + 4 * i) = make_uint4((u32)c[2 * i], (u32)(c[2 * i] >> 32),
                                                                (u32)c[2 * i + 1], (u32)(c[2 * i + 1] >> 32));
    }

    SR_D static void op_crt(u32* rowA) {
        u64 c[24];
        load(c, rowA);
        gl::crt(c);
        store(rowA, c);
    }
    SR_D static void op_icrt(u32* rowA) {
        gl::icrt_row(reinterpret_cast<u64*>(rowA));
    }
    SR_D static void op_ntt_mul(u32* rowA, const u32* rowB) {
        gl::ntt_mul_rolled(reinterpret_cast<u64*>(rowA), reinterpret_cast<const u64*>(rowB));
    }
    SR_D static void op_ring_mul(u32* rowA, const u32* rowB) {
        int trips = 2;
        asm volatile("" : "+r"(trips));  // opaque trip count: one copy of the forward transform
        gl::ring_mul_fused6(reinterpret_cast<u64*>(rowA), reinterpret_cast<u64*>(const_cast<u32*>(rowB)), trips);
    }
};

}  // namespace sr
