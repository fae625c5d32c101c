// Goldilocks batch kernels (instantiations).
#include "gl_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

// One warp per CTA, 16 CTAs per SM: the load / transform / store phases of sixteen independent tiles overlap, where
// four 4-warp CTAs left the SM waiting on global loads (ncu: long_scoreboard 27% of the stall samples in ntt_mul).
// n = 2^22, T = 128 x 4 -> 64 x 8 -> 32 x 16: ring_mul 0.324 -> 0.334 -> 0.340, ntt_mul 0.648 -> 0.692 -> 0.732,
// crt 0.763 -> 0.805 -> 0.824, icrt 0.512 -> 0.520 -> 0.519 of the HBM roofline.
// (Round 2, measured and NOT kept: fetching the next tile's operands with cp.async into the rows that are dead after the
// sextic products / after the result has been read out -- 3.55 ms against 3.51 ms without at n = 2^24: the staging
// code holds 19% of the stall samples, but the kernel is bound by the multiply-add pipe (72% busy), so warps parked
// on their loads cost nothing that the other fifteen do not fill.)
cudaError_t gl_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_batch_op<GLPolicy, OP_CRT, 32, 16>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<GLPolicy, OP_ICRT, 32, 16>(a, b, out, n, st, sms);
    case OP_NTT_MUL: return launch_batch_op<GLPolicy, OP_NTT_MUL, 32, 16>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_batch_op<GLPolicy, OP_RING_MUL, 32, 16>(a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
