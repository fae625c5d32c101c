// Balanced gadget decomposition / recomposition of coefficient-form ring elements: the step
// right before an Ajtai commit (decompose -> CRT -> mat-vec), kept on the device.
//   decompose  decompose_balanced_in_place (balanced_decomposition/mod.rs:62-103) applied per coefficient
//              (coeff_form.rs:588-606) and per element (GadgetDecompose for &[R], mod.rs:163-175):
//              out[j * k + t] = t-th digit element of in[j], digits in [-b/2, b/2] of the signed representative
//              in [-(q-1)/2, (q-1)/2] (fq_convertible.rs:22-35), zero padded to k.
//   recompose  GadgetRecompose for &[R] (mod.rs:177-190): Horner sum_t b^t d_t in the ring.
// Fp64 fields only (Goldilocks, BabyBear): one thread per coefficient, i64 digit loop.  The Starknet prime uses the
// reference's BigInt path (stark_prime/decomposition.rs), which SURVEY marks out of scope.
#include <cuda_runtime.h>

#include "bb_ring.cuh"
#include "gl_ring.cuh"

namespace sr {

struct GLD {
    static constexpr int D = 24;
    static constexpr u64 P = gl::P;
    SR_D static u64 to_std(u64 raw) { return gl::canon(gl::mul_pow2<128>(raw)); }    // raw = x 2^64
    SR_D static u64 from_std(u64 x) { return gl::canon(gl::mul_pow2<64>(x)); }
    SR_D static u64 mul_std(u64 raw, u64 b_std) { return gl::canon(gl::mul(raw, b_std)); }
    SR_D static u64 add(u64 a, u64 b) { return gl::canon(gl::add(a, b)); }
};
struct BBD {
    static constexpr int D = 72;
    static constexpr u64 P = bb::P;
    // 2^64 mod p in operand form (x 2^32) = 2^96 mod p
    static constexpr u32 R64_M32 = bb::cmulmod(bb::cmulmod(bb::R32, bb::R32), bb::R32);
    SR_D static u64 to_std(u64 raw) { return bb::red((u64)bb::red((u64)(u32)raw)); }  // raw = x 2^64
    SR_D static u64 from_std(u64 x) { return bb::mulc((u32)x, R64_M32); }
    SR_D static u64 mul_std(u64 raw, u64 b_std) {  // b_std < p; operand form computed on the fly
        const u32 bm = bb::mulc((u32)b_std, bb::cmulmod(bb::R32, bb::R32));          // b 2^32
        return bb::mulc((u32)raw, bm);
    }
    SR_D static u64 add(u64 a, u64 b) { return bb::add((u32)a, (u32)b); }
};

template <class F>
__global__ void __launch_bounds__(256)
decompose_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n, long long b, int shift, int pad,
                 int* overflow) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * F::D) return;
    const size_t j = idx / F::D;
    const int i = (int)(idx - j * F::D);
    const u64 x = F::to_std(in[idx]);
    // signed representative in [-(q-1)/2, (q-1)/2]
    long long cur = (x > (F::P - 1) / 2) ? (long long)(x - F::P) : (long long)x;
    const long long bh = b / 2;
    u64* o = out + (j * (size_t)pad) * F::D + i;
    int t = 0;
    for (; t < pad; t++) {
        long long rem, q;  // truncating division, like Rust's i128 `%` and `/`
        if (shift >= 0) {  // b = 2^shift (the usual gadget bases): shift and mask on the magnitude, sign restored
            const unsigned long long mag = cur < 0 ? (unsigned long long)(-cur) : (unsigned long long)cur;
            const long long r = (long long)(mag & (unsigned long long)(b - 1)), qq = (long long)(mag >> shift);
            rem = cur < 0 ? -r : r;
            q = cur < 0 ? -qq : qq;
        } else {
            rem = cur % b;
            q = cur / b;
        }
        long long digit;
        if ((rem < 0 ? -rem : rem) <= bh) {
            digit = rem;
            cur = q;
        } else {
            digit = rem < 0 ? rem + b : rem - b;
            cur = q + (rem < 0 ? -1 : 1);  // rounded_div(rem, b) = +-1 here
        }
        const u64 mag = (u64)(digit < 0 ? -digit : digit) % F::P;
        const u64 dstd = (digit < 0 && mag) ? F::P - mag : mag;
        o[(size_t)t * F::D] = F::from_std(dstd);
        if (cur == 0) { t++; break; }
    }
    if (cur != 0) atomicOr(overflow, 1);  // the reference indexes past `out` and panics
    for (; t < pad; t++) o[(size_t)t * F::D] = 0;
}

template <class F>
__global__ void __launch_bounds__(256)
recompose_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n, u64 b_std, int pad) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * F::D) return;
    const size_t j = idx / F::D;
    const int i = (int)(idx - j * F::D);
    const u64* d = in + (j * (size_t)pad) * F::D + i;
    u64 acc = 0;
    for (int t = pad - 1; t >= 0; t--) acc = F::add(F::mul_std(acc, b_std), d[(size_t)t * F::D]);
    out[idx] = acc;
}

// op 0: decompose (n input elements -> n * pad), op 1: recompose (n output elements from n * pad)
cudaError_t decomp_launch(int ring, int op, const u64* in, u64* out, size_t n, unsigned long long b, u64 b_std, int pad,
                          int* overflow, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t D = ring == RING_GL ? 24 : 72;
    const unsigned grid = (unsigned)((n * D + 255) / 256);
    int shift = -1;  // log2(b) when b is a power of two
    if ((b & (b - 1)) == 0)
        for (shift = 0; (1ull << shift) != b; shift++) {}
    if (ring == RING_GL) {
        if (op == 0) decompose_kernel<GLD><<<grid, 256, 0, st>>>(in, out, n, (long long)b, shift, pad, overflow);
        else recompose_kernel<GLD><<<grid, 256, 0, st>>>(in, out, n, b_std, pad);
    } else if (ring == RING_BB) {
        if (op == 0) decompose_kernel<BBD><<<grid, 256, 0, st>>>(in, out, n, (long long)b, shift, pad, overflow);
        else recompose_kernel<BBD><<<grid, 256, 0, st>>>(in, out, n, b_std, pad);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace sr
