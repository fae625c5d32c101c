// Common definitions for the stark-rings sm_100a kernels.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define SR_HD __host__ __device__ __forceinline__
#define SR_D __device__ __forceinline__
#else
#define SR_HD inline
#define SR_D inline
#endif

namespace sr {

typedef uint32_t u32;
typedef uint64_t u64;

SR_HD u32 umin32(u32 a, u32 b) { return a < b ? a : b; }

#if defined(__CUDA_ARCH__)
SR_D u32 mulhi32(u32 a, u32 b) { return __umulhi(a, b); }
SR_D u64 mul64hi(u64 a, u64 b) { return __umul64hi(a, b); }
#else
SR_HD u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
SR_HD u64 mul64hi(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
#endif

// Ring ids of the C ABI (include/stark_rings_cuda.h)
enum { RING_GL = 0, RING_BB = 1, RING_SP = 2 };
// Batch operations
enum { OP_CRT = 0, OP_ICRT = 1, OP_NTT_MUL = 2, OP_RING_MUL = 3 };

}  // namespace sr
