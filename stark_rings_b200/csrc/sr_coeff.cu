// Coefficient-form helpers next to the hot path, batched so that data stays on the device
// between CRT / multiplication calls:
//   reduce   CyclotomicConfig::reduce_in_place   goldilocks/mod.rs:75-98, babybear/mod.rs:87-110 (X^D = X^(D/2) - 1),
//                                                stark_prime/mod.rs:40-47 (X^16 = -1); input of up to 2D coefficients
//   rot      Cyclotomic::rot (multiply by X)     goldilocks/mod.rs:138-149, babybear/mod.rs:150-161, stark_prime/mod.rs:87-95
// One thread per output coefficient; pure streaming kernels (HBM-bound, coalesced 8 / 32-byte accesses).
#include <cuda_runtime.h>

#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "sp_ring.cuh"

namespace sr {

struct GLF {
    static constexpr int N = 1, D = 24;
    static constexpr bool PHI3 = true;
    typedef u64 V;
    SR_D static V load(const u64* p) { return p[0]; }
    SR_D static void store(u64* p, V v) { p[0] = v; }
    SR_D static V zero() { return 0; }
    SR_D static V add(V a, V b) { return gl::canon(gl::add(a, b)); }
    SR_D static V sub(V a, V b) { return gl::canon(gl::sub(a, b)); }
};
struct BBF {
    static constexpr int N = 1, D = 72;
    static constexpr bool PHI3 = true;
    typedef u32 V;
    SR_D static V load(const u64* p) { return (u32)p[0]; }
    SR_D static void store(u64* p, V v) { p[0] = v; }
    SR_D static V zero() { return 0; }
    SR_D static V add(V a, V b) { return bb::add(a, b); }
    SR_D static V sub(V a, V b) { return bb::sub(a, b); }
};
struct SPF {
    static constexpr int N = 4, D = 16;
    static constexpr bool PHI3 = false;
    typedef sp::Fe V;
    SR_D static V load(const u64* p) {
        V v;
        const uint4 lo = reinterpret_cast<const uint4*>(p)[0], hi = reinterpret_cast<const uint4*>(p)[1];
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    SR_D static void store(u64* p, const V& v) {
        reinterpret_cast<uint4*>(p)[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
        reinterpret_cast<uint4*>(p)[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
    SR_D static V zero() {
        V v;
#pragma unroll
        for (int i = 0; i < 8; i++) v.v[i] = 0;
        return v;
    }
    SR_D static V add(const V& a, const V& b) { V r; sp::add(r, a, b); return r; }
    SR_D static V sub(const V& a, const V& b) { V r; sp::sub(r, a, b); return r; }
};

constexpr int COEFF_PER_THREAD = 4;

// out[e][i] for i < D from in[e][0 .. len), D <= len <= 2D
template <class F>
__global__ void __launch_bounds__(256)
reduce_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n, int len) {
  // COEFF_PER_THREAD outputs per thread, one block width apart: coalesced, several independent loads in flight
#pragma unroll
  for (int rep = 0; rep < COEFF_PER_THREAD; rep++) {
    const size_t idx = ((size_t)blockIdx.x * COEFF_PER_THREAD + rep) * 256 + threadIdx.x;
    if (idx >= n * F::D) return;
    const size_t e = idx / F::D;
    const int i = (int)(idx - e * F::D);
    const u64* c = in + e * (size_t)len * F::N;
    auto get = [&](int k) { return k < len ? F::load(c + (size_t)k * F::N) : F::zero(); };
    typename F::V r = get(i);
    if (F::PHI3) {
        constexpr int H = F::D / 2;
        if (i < H) {
            r = F::sub(r, get(F::D + i));
            r = F::sub(r, get(F::D + H + i));
        } else {
            r = F::add(r, get(H + i));
        }
    } else {
        r = F::sub(r, get(F::D + i));
    }
    F::store(out + idx * F::N, r);
  }
}

// out = X * in (mod Phi), per element
template <class F>
__global__ void __launch_bounds__(256)
rot_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n) {
#pragma unroll
  for (int rep = 0; rep < COEFF_PER_THREAD; rep++) {
    const size_t idx = ((size_t)blockIdx.x * COEFF_PER_THREAD + rep) * 256 + threadIdx.x;
    if (idx >= n * F::D) return;
    const size_t e = idx / F::D;
    const int i = (int)(idx - e * F::D);
    const u64* c = in + e * (size_t)F::D * F::N;
    typename F::V r;
    if (i == 0) {
        r = F::sub(F::zero(), F::load(c + (size_t)(F::D - 1) * F::N));
    } else {
        r = F::load(c + (size_t)(i - 1) * F::N);
        if (F::PHI3 && i == F::D / 2) r = F::add(r, F::load(c + (size_t)(F::D - 1) * F::N));
    }
    F::store(out + idx * F::N, r);
  }
}

// Element-wise ring addition / subtraction / negation (ntt_form.rs:588-601 Add, :603-626 Sub / Neg, and the same
// operator impls of coeff_form.rs: both forms add field element by field element, so one kernel serves RqPoly and
// RqNTT): a[i] <- a[i] + b[i] (op 0), a[i] - b[i] (op 1), -a[i] (op 2).  One thread per field element, FE_PER_THREAD
// of them a block width apart: pure streaming, 3 S (add / sub) or 2 S (neg) bytes per element.
template <class F>
__global__ void __launch_bounds__(256)
addsub_kernel(u64* __restrict__ a, const u64* __restrict__ b, size_t nfe, int op) {
#pragma unroll
    for (int rep = 0; rep < COEFF_PER_THREAD; rep++) {
        const size_t idx = ((size_t)blockIdx.x * COEFF_PER_THREAD + rep) * 256 + threadIdx.x;
        if (idx >= nfe) return;
        const typename F::V x = F::load(a + idx * F::N);
        typename F::V r;
        if (op == 0) r = F::add(x, F::load(b + idx * F::N));
        else if (op == 1) r = F::sub(x, F::load(b + idx * F::N));
        else r = F::sub(F::zero(), x);
        F::store(a + idx * F::N, r);
    }
}
template <class F>
static cudaError_t addsub_launch_t(int op, u64* a, const u64* b, size_t n, cudaStream_t st) {
    const size_t total = n * F::D;
    if (total == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 256 * COEFF_PER_THREAD - 1) / (256 * COEFF_PER_THREAD));
    addsub_kernel<F><<<grid, 256, 0, st>>>(a, b, total, op);
    return cudaGetLastError();
}
// op 0 = add, 1 = sub, 2 = neg (b unused); n ring elements, in place on a
cudaError_t addsub_launch(int ring, int op, u64* a, const u64* b, size_t n, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return addsub_launch_t<GLF>(op, a, b, n, st);
    case RING_BB: return addsub_launch_t<BBF>(op, a, b, n, st);
    case RING_SP: return addsub_launch_t<SPF>(op, a, b, n, st);
    }
    return cudaErrorInvalidValue;
}

template <class F>
static cudaError_t coeff_launch_t(int op, const u64* in, u64* out, size_t n, int len, cudaStream_t st) {
    const size_t total = n * F::D;
    if (total == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 256 * COEFF_PER_THREAD - 1) / (256 * COEFF_PER_THREAD));
    if (op == 0) reduce_kernel<F><<<grid, 256, 0, st>>>(in, out, n, len);
    else rot_kernel<F><<<grid, 256, 0, st>>>(in, out, n);
    return cudaGetLastError();
}

// op 0 = reduce (len coefficients per input polynomial), op 1 = rot
cudaError_t coeff_launch(int ring, int op, const u64* in, u64* out, size_t n, int len, cudaStream_t st) {
    switch (ring) {
    case RING_GL: return coeff_launch_t<GLF>(op, in, out, n, len, st);
    case RING_BB: return coeff_launch_t<BBF>(op, in, out, n, len, st);
    case RING_SP: return coeff_launch_t<SPF>(op, in, out, n, len, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
