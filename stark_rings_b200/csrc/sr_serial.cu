// Canonical (de)serialization of batches on the device.
//
// Replaces, for a batch, CanonicalSerialize / CanonicalDeserialize of the ring types
// (cyclotomic_ring/coeff_form.rs:154-189: the element is its [Fp; D] array; ntt_form.rs:24: derived, the extension-field
// coefficients in memory order), i.e. ark-serialize 0.4 applied to each field element in turn.  ark-serialize is not
// vendored in the reference tree; its published format, restated here: a prime-field element is written as the
// little-endian bytes of its STANDARD-form integer (not the Montgomery limbs) truncated to ceil(MODULUS_BIT_SIZE / 8)
// bytes = 8 (Goldilocks), 4 (BabyBear, 31-bit modulus), 32 (Starknet prime, 252 bits); arrays carry no length prefix
// (a Vec adds a u64 little-endian length, left to the caller); reading rejects an integer >= p
// (SerializationError::InvalidData).  PARITY UNPINNED: the reference holds no serialized test vector, so the byte
// format rests on the restated ark-serialize rules above.
//
// One thread per field element; pure streaming kernels.  The conversions out of / into Montgomery form are
// x * 2^-64 / x * 2^64 (Goldilocks: shifts, 2^64 = 2^32 - 1; BabyBear: two 32-bit reductions / one multiplication by
// 2^96 mod p) and one Montgomery multiplication by 1 / by R^2 mod p (Starknet prime).
#include <cuda_runtime.h>

#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "sp_ring.cuh"

namespace sr {

struct GLSer {
    static constexpr int LIMBS = 1, BYTES = 8;
    SR_D static void ser(const u64* in, unsigned char* out) {
        *reinterpret_cast<u64*>(out) = gl::canon(gl::mul_pow2<128>(in[0]));  // 2^-64 = 2^128
    }
    SR_D static bool de(const unsigned char* in, u64* out) {
        const u64 x = *reinterpret_cast<const u64*>(in);
        out[0] = gl::canon(gl::mul_pow2<64>(x));  // any u64 is a valid weak input; the range check is separate
        return x < gl::P;
    }
};
struct BBSer {
    static constexpr int LIMBS = 1, BYTES = 4;
    static constexpr u32 C96 = 0x12f37bfbu;  // 2^96 mod p
    static_assert(bb::cmulmod(bb::cmulmod(bb::R32, bb::R32), bb::R32) == C96, "2^96 mod p");
    SR_D static void ser(const u64* in, unsigned char* out) {
        const u32 m = (u32)in[0];  // x * 2^64 mod p
        *reinterpret_cast<u32*>(out) = bb::red((u64)bb::red((u64)m));
    }
    SR_D static bool de(const unsigned char* in, u64* out) {
        const u32 x = *reinterpret_cast<const u32*>(in);
        out[0] = x < bb::P ? bb::mulc(x, C96) : 0u;  // x * 2^96 * 2^-32
        return x < bb::P;
    }
};
struct SPSer {
    static constexpr int LIMBS = 4, BYTES = 32;
    SR_D static sp::Fe load8(const void* p) {
        const uint4 lo = reinterpret_cast<const uint4*>(p)[0], hi = reinterpret_cast<const uint4*>(p)[1];
        sp::Fe v;
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    SR_D static void store8(void* p, const sp::Fe& v) {
        reinterpret_cast<uint4*>(p)[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
        reinterpret_cast<uint4*>(p)[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
    SR_D static void ser(const u64* in, unsigned char* out) {
        constexpr u32 one[8] = {1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        sp::Fe r;
        sp::mont_mul_limbs(r, load8(in).v, one);  // x R * 1 / R
        store8(out, r);
    }
    SR_D static bool de(const unsigned char* in, u64* out) {
        constexpr u32 r2[8] = {0x7e000401u, 0xfffffd73u, 0x330fffffu, 0x00000001u,
                               0xff6f8000u, 0xffffffffu, 0x5e008810u, 0x07ffd4abu};  // 2^512 mod p
        const sp::Fe x = load8(in);
        bool lt = false;  // x < p, most significant limb first
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            const u32 pl = sp::p_limb(i);
            if (x.v[i] != pl) {
                lt = x.v[i] < pl;
                break;
            }
        }
        sp::Fe r;
        if (lt) sp::mont_mul_limbs(r, x.v, r2);  // x * R^2 / R
        else {
#pragma unroll
            for (int i = 0; i < 8; i++) r.v[i] = 0;
        }
        store8(out, r);
        return lt;
    }
};

constexpr int FE_PER_THREAD = 4;

template <class F>
__global__ void __launch_bounds__(256)
serialize_kernel(const u64* __restrict__ in, unsigned char* __restrict__ out, size_t nfe) {
#pragma unroll
    for (int rep = 0; rep < FE_PER_THREAD; rep++) {  // one block width apart: coalesced, independent loads in flight
        const size_t i = ((size_t)blockIdx.x * FE_PER_THREAD + rep) * 256 + threadIdx.x;
        if (i < nfe) F::ser(in + i * F::LIMBS, out + i * F::BYTES);
    }
}
template <class F>
__global__ void __launch_bounds__(256)
deserialize_kernel(const unsigned char* __restrict__ in, u64* __restrict__ out, size_t nfe, int* __restrict__ bad) {
#pragma unroll
    for (int rep = 0; rep < FE_PER_THREAD; rep++) {
        const size_t i = ((size_t)blockIdx.x * FE_PER_THREAD + rep) * 256 + threadIdx.x;
        if (i < nfe && !F::de(in + i * F::BYTES, out + i * F::LIMBS)) *bad = 1;
    }
}

template <class F>
static cudaError_t serial_t(int op, const void* in, void* out, size_t nfe, int* bad, cudaStream_t st) {
    if (nfe == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((nfe + 256 * FE_PER_THREAD - 1) / (256 * FE_PER_THREAD));
    if (op == 0) serialize_kernel<F><<<grid, 256, 0, st>>>((const u64*)in, (unsigned char*)out, nfe);
    else deserialize_kernel<F><<<grid, 256, 0, st>>>((const unsigned char*)in, (u64*)out, nfe, bad);
    return cudaGetLastError();
}
// op 0: limbs -> bytes, op 1: bytes -> limbs (sets *bad when an integer >= p is met); nfe = number of field elements
cudaError_t serial_launch(int ring, int op, const void* in, void* out, size_t nfe, int* bad, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return serial_t<GLSer>(op, in, out, nfe, bad, st);
    case RING_BB: return serial_t<BBSer>(op, in, out, nfe, bad, st);
    case RING_SP: return serial_t<SPSer>(op, in, out, nfe, bad, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
