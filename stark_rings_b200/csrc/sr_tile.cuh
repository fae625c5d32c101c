// Tile staging between HBM and shared memory, and the per-ring row policies.
//
// A batch is a dense array of ring elements in the reference's layout (S = 192 / 576 / 512 bytes
// per element).  A CTA of T threads owns a tile of T consecutive elements = one contiguous
// T*S-byte span of HBM.  The span is moved with fully coalesced 128-bit accesses (thread i moves
// 16-byte chunks i, i+T, ...) into shared-memory rows, one row per element, padded by 16 bytes so
// that the per-thread 128-bit row accesses (thread t owns row t) are bank-conflict free:
//   row stride 52 / 76 / 132 words  ->  8 consecutive rows start 20 / 12 / 4 words apart mod 32.
#pragma once
#include "sr_common.cuh"

namespace sr {

SR_D uint4 ld_stream(const uint4* p) { return __ldcs(p); }
SR_D void st_stream(uint4* p, uint4 v) { __stcs(p, v); }

// R::CHUNKS   16-byte chunks per element in HBM
// R::ROW      u32 words per shared-memory row
// R::put / R::get  move one chunk into / out of a row
template <class R, int T>
SR_D void stage_in(u32* __restrict__ s, const u64* __restrict__ g, int ne) {
    const uint4* g4 = reinterpret_cast<const uint4*>(g);
    const int total = ne * R::CHUNKS;
    // U independent 128-bit loads in flight per thread (registers are free during staging)
    constexpr int U = R::STAGE_UNROLL;
    int c = threadIdx.x;
    for (; c + (U - 1) * T < total; c += U * T) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld_stream(g4 + c + u * T);
#pragma unroll
        for (int u = 0; u < U; u++) {
            int cc = c + u * T, e = cc / R::CHUNKS, j = cc - e * R::CHUNKS;
            R::put(s + e * R::ROW, j, v[u]);
        }
    }
    for (; c < total; c += T) {
        int e = c / R::CHUNKS, j = c - e * R::CHUNKS;
        R::put(s + e * R::ROW, j, ld_stream(g4 + c));
    }
}

template <class R, int T>
SR_D void stage_out(u64* __restrict__ g, const u32* __restrict__ s, int ne) {
    uint4* g4 = reinterpret_cast<uint4*>(g);
    const int total = ne * R::CHUNKS;
#pragma unroll 4
    for (int c = threadIdx.x; c < total; c += T) {
        int e = c / R::CHUNKS, j = c - e * R::CHUNKS;
        st_stream(g4 + c, R::get(s + e * R::ROW, j));
    }
}

// Row <-> registers with 128-bit shared-memory accesses
template <int N>
SR_D void row_load(u32 (&r)[N], const u32* row) {
    static_assert(N % 4 == 0, "row length");
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
        uint4 v = *reinterpret_cast<const uint4*>(row + 4 * i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
}
template <int N>
SR_D void row_store(u32* row, const u32 (&r)[N]) {
#pragma unroll
    for (int i = 0; i < N / 4; i++)
        *reinterpret_cast<uint4*>(row + 4 * i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}

}  // namespace sr
