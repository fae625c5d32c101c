// Starknet-prime batch kernels (instantiations).
#include "sp_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

#ifndef SR_SP_T
#define SR_SP_T 64
#endif

cudaError_t sp_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_batch_op<SPPolicy, OP_CRT, SR_SP_T, 4>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<SPPolicy, OP_ICRT, SR_SP_T, 4>(a, b, out, n, st, sms);
    case OP_NTT_MUL: return launch_batch_op<SPPolicy, OP_NTT_MUL, SR_SP_T, 4>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_batch_op<SPPolicy, OP_RING_MUL, SR_SP_T, 3>(a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
