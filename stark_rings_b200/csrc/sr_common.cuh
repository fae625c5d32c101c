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

// Hand-off of a partial commitment through a peer-memory mailbox (sr_matvec.cu).  role 0 = plain kernel.
struct PeerSync {
    int role;               // 1: writer (the tail of a rank's product kernel), 2: the root's separate reduction kernel,
                            // 3: root, fused (the tail of the root's own product kernel also sums all ranks)
    int nranks, rank;
    u64* slots;             // mailbox slots [MAILBOX_DEPTH][nranks][slot_stride] (the root's memory)
    size_t slot_stride;     // u64 per (epoch, rank) slot
    u64* flags;             // [nranks] last epoch each rank has delivered (the root's memory)
    u64* consumed;          // last epoch the root has summed (the root's memory)
    u64 epoch;              // epoch of this call, or 0: take *epoch_ctr + 1 (device-resident, CUDA-graph friendly)
    u64* epoch_ctr;         // device-local count of the epochs this rank has delivered (role 2: == consumed)
    unsigned* counter;      // device-local block-arrival counter (zero between launches)
    int* err;               // set to 1 when a wait times out
    unsigned long long timeout_ns;  // budget of one wait (a lost peer must not hang the GPU)
};
constexpr int MAILBOX_DEPTH = 4;

}  // namespace sr
