// Goldilocks batch kernels (instantiations).
#include "gl_policy.cuh"
#include "sr_batch_kernel.cuh"

namespace sr {

#ifndef SR_GL_T
#define SR_GL_T 128
#endif
#ifndef SR_GL_RM_T   // fused ring mul: threads per CTA / resident CTAs per SM
#define SR_GL_RM_T 128
#endif
#ifndef SR_GL_RM_MINB
#define SR_GL_RM_MINB 4
#endif
#ifndef SR_GL_MINB
#define SR_GL_MINB 4
#endif

cudaError_t gl_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (op) {
    case OP_CRT: return launch_batch_op<GLPolicy, OP_CRT, SR_GL_T, SR_GL_MINB>(a, b, out, n, st, sms);
    case OP_ICRT: return launch_batch_op<GLPolicy, OP_ICRT, SR_GL_T, SR_GL_MINB>(a, b, out, n, st, sms);
    case OP_NTT_MUL: return launch_batch_op<GLPolicy, OP_NTT_MUL, SR_GL_T, SR_GL_MINB>(a, b, out, n, st, sms);
    case OP_RING_MUL: return launch_batch_op<GLPolicy, OP_RING_MUL, SR_GL_RM_T, SR_GL_RM_MINB>(a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
