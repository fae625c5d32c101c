// BabyBear batch kernels (instantiations).
#include "bb_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

#ifndef SR_BB_T
#define SR_BB_T 128
#endif

cudaError_t bb_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_batch_op<BBPolicy, OP_CRT, SR_BB_T, 3>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<BBPolicy, OP_ICRT, SR_BB_T, 3>(a, b, out, n, st, sms);
    case OP_NTT_MUL: return launch_batch_op<BBPolicy, OP_NTT_MUL, SR_BB_T, 2>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_batch_op<BBPolicy, OP_RING_MUL, SR_BB_T, 2>(a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
