// C ABI of libstarkrings_cuda.so (include/stark_rings_cuda.h): context, memory, dispatch, the
// host-buffer pipeline.  No CPU fallback exists: every entry point either runs the sm_100a
// kernels or returns an error.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/stark_rings_cuda.h"
#include "sr_common.cuh"

namespace sr {
// per-ring launchers (bb_kernels.cu, gl_kernels.cu, sp_kernels.cu)
cudaError_t bb_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms);
cudaError_t gl_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms);
cudaError_t sp_launch(int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms);
// mat-vec (sr_matvec.cu): partial products of `ncols` columns -> nrows elements in out; uses
// scratch (>= matvec_scratch_bytes).  Returns the number of kernels launched via *launches.
size_t matvec_scratch_bytes(int ring, size_t nrows, int sms);
cudaError_t matvec_launch(int ring, const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v,
                          u64* out, void* scratch, unsigned* counters, unsigned* seq, bool pdl, cudaStream_t st, int sms,
                          int* launches, const PeerSync* ps);
cudaError_t modsum_launch(int ring, const u64* gathered, size_t nranks, size_t stride_rows, size_t nrows, u64* out,
                          cudaStream_t st, const PeerSync* ps);
// coefficient-form helpers (sr_coeff.cu): op 0 = reduce, op 1 = rot
cudaError_t coeff_launch(int ring, int op, const u64* in, u64* out, size_t n, int len, cudaStream_t st);
// element-wise add / sub / neg (sr_coeff.cu): op 0 = add, 1 = sub, 2 = neg
cudaError_t addsub_launch(int ring, int op, u64* a, const u64* b, size_t n, cudaStream_t st);
// balanced gadget decomposition / recomposition (sr_decomp.cu): op 0 = decompose, op 1 = recompose
cudaError_t decomp_launch(int ring, int op, const u64* in, u64* out, size_t n, unsigned long long b, u64 b_std, int pad,
                          int* overflow, cudaStream_t st);
// sparse mat-vec, dense mat-mat, scalar scaling (sr_sparse.cu)
cudaError_t sparse_matvec_launch(int ring, const u64* row_ptr, const u64* col_idx, const u64* vals, const u64* v,
                                 size_t nrows, size_t ncols, size_t nnz, u64* out, int* bad, cudaStream_t st);
cudaError_t matmat_launch(int ring, const u64* const* a_rows, const u64* const* m_rows, u64* const* out_rows,
                          size_t a_nrows, size_t inner, size_t m_ncols, cudaStream_t st);
cudaError_t scale_launch(int ring, u64* a, const u64* r, size_t n, cudaStream_t st);
cudaError_t sparse_pairs_launch(int ring, const u64* a_vals, const u64* m_vals, const u64* pair_ptr, const u64* pair_a,
                                const u64* pair_m, size_t ncand, u64* out, int* nonzero, cudaStream_t st);
// canonical (de)serialization (sr_serial.cu): op 0 limbs -> bytes, op 1 bytes -> limbs
cudaError_t serial_launch(int ring, int op, const void* in, void* out, size_t nfe, int* bad, cudaStream_t st);
// integer multiply-add peaks of the device (sr_peak.cu)
cudaError_t imad_peak_measure(int sms, cudaStream_t st, double* tops);
}  // namespace sr

using sr::u64;

struct sr_ctx {
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;      // stream used for device-resident calls
    cudaStream_t own_stream = nullptr;  // created by sr_init
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host pipeline
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    // host pipeline staging (device side), grown on demand
    static constexpr int NBUF = 3;
    void* stage[NBUF][3] = {};
    size_t stage_bytes = 0;
    cudaEvent_t ev_in[NBUF] = {}, ev_k[NBUF] = {}, ev_out[NBUF] = {};
    // mat-vec scratch
    void* mv_scratch = nullptr;
    size_t mv_scratch_bytes = 0;
    void* mv_rows = nullptr;  // device copy of the row-pointer table
    size_t mv_rows_cap = 0;
    std::vector<const void*> mv_rows_cached;  // host copy of what mv_rows holds (skip re-upload if unchanged)
    unsigned* commit_counters = nullptr;      // [4]: [0] mat-vec tail ticket, [1] mailbox-reduction ticket, [2..3] chunk counters
    unsigned mv_seq = 0;                      // mat-vec launches so far (alternates the chunk counter)
    bool pipelined = false;                   // sr_set_pipelined: successive products may overlap (PDL)
    cudaEvent_t ev_stream = nullptr;          // orders a newly selected stream after the previous one (sr_set_stream)
    std::vector<std::pair<void*, size_t>> pool;  // grow-only device scratch of the host-buffer paths (DevTemps)
    int* dflag = nullptr;                     // device error flag of the checking kernels (allocated once)
    int* hflag = nullptr;                     // its pinned host mirror
};

// Mailbox of the column-sharded commitment: device memory of ONE rank (the root) that every rank maps
// through CUDA IPC.  Layout: header { flags[MAX_RANKS], consumed, err } in the first 4 KiB, then
// MAILBOX_DEPTH x nranks x nrows_max partial elements.
struct sr_mailbox {
    static constexpr int MAX_RANKS = 64;
    static constexpr size_t HEADER = 4096;
    int ring = 0;
    size_t nrows_max = 0;
    int nranks = 0;
    bool owner = false;  // created here (cudaMalloc) or opened from a peer's IPC handle
    void* base = nullptr;
    size_t bytes = 0;
    u64* sent = nullptr;  // device-LOCAL count of the epochs this process has delivered into this mailbox
    unsigned long long timeout_ns = 4000000000ull;  // budget of one in-kernel wait (sr_mailbox_set_timeout)
    u64* flags() const { return reinterpret_cast<u64*>(base); }
    u64* consumed() const { return reinterpret_cast<u64*>(base) + MAX_RANKS; }
    int* err() const { return reinterpret_cast<int*>(reinterpret_cast<u64*>(base) + MAX_RANKS + 1); }
    u64* slots() const { return reinterpret_cast<u64*>(reinterpret_cast<char*>(base) + HEADER); }
};

namespace {

int fail(sr_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
int cuda_fail(sr_ctx* c, cudaError_t e, const char* what) {
    return fail(c, SR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                               \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); \
    } while (0)

size_t elem_limbs(int ring) {
    switch (ring) {
    case SR_GOLDILOCKS: return 24;
    case SR_BABYBEAR: return 72;
    case SR_STARK: return 64;
    }
    return 0;
}

cudaError_t launch(int ring, int op, const u64* a, const u64* b, u64* out, size_t n, cudaStream_t st, int sms) {
    switch (ring) {
    case SR_GOLDILOCKS: return sr::gl_launch(op, a, b, out, n, st, sms);
    case SR_BABYBEAR: return sr::bb_launch(op, a, b, out, n, st, sms);
    case SR_STARK: return sr::sp_launch(op, a, b, out, n, st, sms);
    }
    return cudaErrorInvalidValue;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int ensure_stage(sr_ctx* ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return SR_OK;
    for (int i = 0; i < sr_ctx::NBUF; i++)
        for (int j = 0; j < 3; j++) {
            if (ctx->stage[i][j]) cudaFree(ctx->stage[i][j]);
            ctx->stage[i][j] = nullptr;
        }
    ctx->stage_bytes = 0;
    for (int i = 0; i < sr_ctx::NBUF; i++)
        for (int j = 0; j < 2; j++) CU(cudaMalloc(&ctx->stage[i][j], bytes));
    ctx->stage_bytes = bytes;
    return SR_OK;
}

// Host-buffer path: the batch is cut into chunks that flow through a 3-deep ring of device
// buffers; H2D of chunk i+1, the kernel of chunk i and D2H of chunk i-1 overlap on three streams.
int batch_host(sr_ctx* ctx, int ring, int op, const u64* a, const u64* b, u64* out, size_t n) {
    const size_t w = elem_limbs(ring), ebytes = w * 8;
    const bool two = (op == sr::OP_NTT_MUL || op == sr::OP_RING_MUL);
    size_t chunk_elems = ((size_t)64 << 20) / ebytes;  // 64 MiB per operand per chunk
    if (chunk_elems > n) chunk_elems = n;
    if (chunk_elems == 0) return SR_OK;
    int rc = ensure_stage(ctx, chunk_elems * ebytes);
    if (rc) return rc;
    const size_t nchunks = (n + chunk_elems - 1) / chunk_elems;
    for (size_t c = 0; c < nchunks; c++) {
        const int s = (int)(c % sr_ctx::NBUF);
        const size_t e0 = c * chunk_elems, ne = (n - e0 < chunk_elems) ? n - e0 : chunk_elems;
        u64* dA = (u64*)ctx->stage[s][0];
        u64* dB = (u64*)ctx->stage[s][1];
        // the slot is free once the D2H of the chunk that used it last has finished
        if (c >= (size_t)sr_ctx::NBUF) CU(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_out[s], 0));
        CU(cudaMemcpyAsync(dA, a + e0 * w, ne * ebytes, cudaMemcpyHostToDevice, ctx->copy_in));
        if (two) CU(cudaMemcpyAsync(dB, b + e0 * w, ne * ebytes, cudaMemcpyHostToDevice, ctx->copy_in));
        CU(cudaEventRecord(ctx->ev_in[s], ctx->copy_in));
        CU(cudaStreamWaitEvent(ctx->own_stream, ctx->ev_in[s], 0));
        CU(launch(ring, op, dA, dB, dA, ne, ctx->own_stream, ctx->sms));
        ctx->launches++;
        CU(cudaEventRecord(ctx->ev_k[s], ctx->own_stream));
        CU(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_k[s], 0));
        CU(cudaMemcpyAsync(out + e0 * w, dA, ne * ebytes, cudaMemcpyDeviceToHost, ctx->copy_out));
        CU(cudaEventRecord(ctx->ev_out[s], ctx->copy_out));
    }
    CU(cudaStreamSynchronize(ctx->copy_out));
    return SR_OK;
}

int batch(sr_ctx* ctx, int ring, int op, const u64* a, const u64* b, u64* out, size_t n_limbs, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (n_limbs % w != 0)
        return fail(ctx, SR_ERR_BAD_LENGTH,
                    "slice length " + std::to_string(n_limbs) + " is not a multiple of " + std::to_string(w));
    if (n_limbs == 0) return SR_OK;
    const bool two = (op == sr::OP_NTT_MUL || op == sr::OP_RING_MUL);
    if (!a || !out || (two && !b)) return fail(ctx, SR_ERR_INVALID, "null buffer");
    CU(cudaSetDevice(ctx->device));
    const size_t n = n_limbs / w;
    if (loc == SR_DEVICE) {
        if (!aligned16(a) || !aligned16(out) || (two && !aligned16(b)))
            return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
        CU(launch(ring, op, a, b, out, n, ctx->stream, ctx->sms));
        ctx->launches++;
        return SR_OK;
    }
    if (loc == SR_HOST) return batch_host(ctx, ring, op, a, b, out, n);
    return fail(ctx, SR_ERR_INVALID, "unknown loc");
}

int ensure_counters(sr_ctx* ctx) {
    if (ctx->commit_counters) return SR_OK;
    CU(cudaMalloc((void**)&ctx->commit_counters, 4 * sizeof(unsigned)));
    CU(cudaMemset(ctx->commit_counters, 0, 4 * sizeof(unsigned)));
    return SR_OK;
}

// Grow-only device scratch of the host-buffer paths: slot k serves the k-th temporary of a call.  Calls hold the
// context mutex and synchronise before they return, so the slots are free again at the next call; no cudaMalloc /
// cudaFree (which synchronise the whole device) in steady state.
cudaError_t pool_get(sr_ctx* ctx, size_t k, size_t bytes, void** p) {
    if (ctx->pool.size() <= k) ctx->pool.resize(k + 1, std::make_pair((void*)nullptr, (size_t)0));
    auto& slot = ctx->pool[k];
    if (slot.second < bytes || !slot.first) {
        if (slot.first) cudaFree(slot.first);
        slot.first = nullptr;
        slot.second = 0;
        const size_t want = bytes < 256 ? 256 : bytes;
        cudaError_t e = cudaMalloc(&slot.first, want);
        if (e != cudaSuccess) return e;
        slot.second = want;
    }
    *p = slot.first;
    return cudaSuccess;
}

int matvec_impl(sr_ctx* ctx, int ring, const u64* const* rows, size_t nrows, size_t ncols, const u64* v,
                size_t v_limbs, u64* out, int loc, const sr::PeerSync* ps = nullptr) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (v_limbs % w != 0 || v_limbs / w != ncols)
        return fail(ctx, SR_ERR_BAD_LENGTH,
                    "DifferentLengths(" + std::to_string(ncols) + ", " + std::to_string(v_limbs / (w ? w : 1)) + ")");
    if (nrows == 0) return SR_OK;
    if (!rows || !out || (ncols && !v)) return fail(ctx, SR_ERR_INVALID, "null buffer");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (loc == SR_DEVICE) ? ctx->stream : ctx->own_stream;
    // scratch + row table
    const size_t need = sr::matvec_scratch_bytes(ring, nrows, ctx->sms);
    if (ctx->mv_scratch_bytes < need) {
        if (ctx->mv_scratch) cudaFree(ctx->mv_scratch);
        ctx->mv_scratch = nullptr;
        ctx->mv_scratch_bytes = 0;
        CU(cudaMalloc(&ctx->mv_scratch, need));
        ctx->mv_scratch_bytes = need;
    }
    if (ctx->mv_rows_cap < nrows) {
        if (ctx->mv_rows) cudaFree(ctx->mv_rows);
        ctx->mv_rows = nullptr;
        ctx->mv_rows_cap = 0;
        ctx->mv_rows_cached.clear();
        CU(cudaMalloc(&ctx->mv_rows, nrows * sizeof(void*)));
        ctx->mv_rows_cap = nrows;
    }
    int rcc = ensure_counters(ctx);
    if (rcc) return rcc;
    int launches = 0;
    if (loc == SR_DEVICE) {
        for (size_t i = 0; i < nrows; i++)
            if (!rows[i] || !aligned16(rows[i])) return fail(ctx, SR_ERR_INVALID, "row pointer null or misaligned");
        // the row table is uploaded only when it changed, so a repeated commit with the same matrix issues
        // kernels only (and can be captured in a CUDA graph; a context whose calls were captured must keep serving
        // the same matrix, because the captured kernels read this context's row table)
        bool same = ctx->mv_rows_cached.size() == nrows;
        for (size_t i = 0; same && i < nrows; i++) same = (ctx->mv_rows_cached[i] == (const void*)rows[i]);
        if (!same) {
            ctx->mv_rows_cached.assign(rows, rows + nrows);
            CU(cudaMemcpyAsync(ctx->mv_rows, ctx->mv_rows_cached.data(), nrows * sizeof(void*),
                               cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        }
        CU(sr::matvec_launch(ring, (const u64* const*)ctx->mv_rows, nrows, ncols, v, out, ctx->mv_scratch,
                             ctx->commit_counters, &ctx->mv_seq, ctx->pipelined, st, ctx->sms, &launches, ps));
        ctx->launches += launches;
        return SR_OK;
    }
    if (loc != SR_HOST || ps) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    // Host path: whole operands are copied to the device (a commitment matrix is normally kept resident with
    // SR_DEVICE; this path exists for drop-in completeness).  The device copies live in the context's grow-only
    // pool: no per-call cudaMalloc / cudaFree.
    const size_t row_bytes = ncols * w * 8;
    std::vector<void*> drows(nrows, nullptr);
    void* dv = nullptr;
    void* dout = nullptr;
    CU(pool_get(ctx, 0, nrows * w * 8, &dout));
    if (ncols) {
        CU(pool_get(ctx, 1, row_bytes, &dv));
        CU(cudaMemcpyAsync(dv, v, row_bytes, cudaMemcpyHostToDevice, st));
        for (size_t i = 0; i < nrows; i++) {
            if (!rows[i]) return fail(ctx, SR_ERR_INVALID, "null row pointer");
            CU(pool_get(ctx, 2 + i, row_bytes, &drows[i]));
            CU(cudaMemcpyAsync(drows[i], rows[i], row_bytes, cudaMemcpyHostToDevice, st));
        }
    }
    ctx->mv_rows_cached.clear();
    CU(cudaMemcpyAsync(ctx->mv_rows, drows.data(), nrows * sizeof(void*), cudaMemcpyHostToDevice, st));
    CU(sr::matvec_launch(ring, (const u64* const*)ctx->mv_rows, nrows, ncols, (const u64*)dv, (u64*)dout,
                         ctx->mv_scratch, ctx->commit_counters, &ctx->mv_seq, false, st, ctx->sms, &launches, nullptr));
    ctx->launches += launches;
    CU(cudaMemcpyAsync(out, dout, nrows * w * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SR_OK;
}

// Device error flag shared by the kernels that report a data-dependent failure (decomposition overflow, column
// index out of range, integer not below the modulus): allocated once per context, so that such calls cost one small
// async copy and one stream synchronisation instead of a cudaMalloc / cudaFree pair (which synchronises the device).
int flag_begin(sr_ctx* ctx, cudaStream_t st) {
    if (!ctx->dflag) {
        CU(cudaMalloc((void**)&ctx->dflag, sizeof(int)));
        CU(cudaMallocHost((void**)&ctx->hflag, sizeof(int)));
    }
    CU(cudaMemsetAsync(ctx->dflag, 0, sizeof(int), st));
    return SR_OK;
}
// copies the flag back and waits for the stream; *flag receives its value
int flag_end(sr_ctx* ctx, cudaStream_t st, int* flag) {
    CU(cudaMemcpyAsync(ctx->hflag, ctx->dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *flag = *ctx->hflag;
    return SR_OK;
}

// Temporary device buffers of one call (host-buffer paths of the linear-algebra entry points), drawn from the
// context's grow-only pool (pool_get): slot k = the k-th temporary of the call.
struct DevTemps {
    sr_ctx* ctx;
    size_t next = 0;
    explicit DevTemps(sr_ctx* c) : ctx(c) {}
    cudaError_t alloc(void** p, size_t bytes) { return pool_get(ctx, next++, bytes ? bytes : 16, p); }
    // device copy of a host array (async on st)
    cudaError_t upload(void** p, const void* host, size_t bytes, cudaStream_t st) {
        cudaError_t e = alloc(p, bytes);
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*p, host, bytes, cudaMemcpyHostToDevice, st);
        return e;
    }
};

int sparse_matvec_impl(sr_ctx* ctx, int ring, size_t nrows, size_t ncols, const u64* row_ptr, const u64* col_idx,
                       const u64* vals, const u64* v, size_t v_limbs, u64* out, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (v_limbs % w != 0 || v_limbs / w != ncols)
        return fail(ctx, SR_ERR_BAD_LENGTH,
                    "DifferentLengths(" + std::to_string(ncols) + ", " + std::to_string(v_limbs / w) + ")");
    if (nrows == 0) return SR_OK;
    if (!row_ptr || !out) return fail(ctx, SR_ERR_INVALID, "null buffer");
    if (loc != SR_DEVICE && loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (loc == SR_DEVICE) ? ctx->stream : ctx->own_stream;
    // nnz = row_ptr[nrows]
    u64 ends[2] = {0, 0};
    if (loc == SR_DEVICE) {
        CU(cudaMemcpyAsync(&ends[0], row_ptr, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&ends[1], row_ptr + nrows, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    } else {
        ends[0] = row_ptr[0];
        ends[1] = row_ptr[nrows];
    }
    if (ends[0] != 0) return fail(ctx, SR_ERR_INVALID, "row_ptr[0] must be 0");
    const size_t nnz = (size_t)ends[1];
    if (nnz && (!col_idx || !vals || !v)) return fail(ctx, SR_ERR_INVALID, "null buffer");
    DevTemps tmp(ctx);
    int rcf = flag_begin(ctx, st);
    if (rcf) return rcf;
    int* dbad = ctx->dflag;
    const u64 *k_rp = row_ptr, *k_ci = col_idx, *k_vals = vals, *k_v = v;
    u64* k_out = out;
    if (loc == SR_HOST) {
        for (size_t i = 0; i < nrows; i++)
            if (row_ptr[i] > row_ptr[i + 1]) return fail(ctx, SR_ERR_INVALID, "row_ptr must be non-decreasing");
        CU(tmp.upload((void**)&k_rp, row_ptr, (nrows + 1) * 8, st));
        CU(tmp.upload((void**)&k_ci, col_idx, nnz * 8, st));
        CU(tmp.upload((void**)&k_vals, vals, nnz * w * 8, st));
        CU(tmp.upload((void**)&k_v, v, ncols * w * 8, st));
        CU(tmp.alloc((void**)&k_out, nrows * w * 8));
    } else if (!aligned16(vals) || !aligned16(v) || !aligned16(out)) {
        return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
    }
    CU(sr::sparse_matvec_launch(ring, k_rp, k_ci, k_vals, k_v, nrows, ncols, nnz, k_out, dbad, st));
    ctx->launches++;
    if (loc == SR_HOST) CU(cudaMemcpyAsync(out, k_out, nrows * w * 8, cudaMemcpyDeviceToHost, st));
    int bad = 0;
    rcf = flag_end(ctx, st, &bad);
    if (rcf) return rcf;
    if (bad == 2) return fail(ctx, SR_ERR_INVALID, "row_ptr must be non-decreasing and end at nnz");
    if (bad) return fail(ctx, SR_ERR_INVALID, "column index out of range (the reference panics on v[i])");
    return SR_OK;
}

int matmat_impl(sr_ctx* ctx, int ring, const u64* const* a_rows, size_t a_nrows, size_t a_ncols,
                const u64* const* m_rows, size_t m_nrows, size_t m_ncols, u64* const* out_rows, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (a_ncols != m_nrows)
        return fail(ctx, SR_ERR_BAD_LENGTH,
                    "DifferentLengths(" + std::to_string(a_ncols) + ", " + std::to_string(m_nrows) + ")");
    if (a_nrows == 0 || m_ncols == 0) return SR_OK;
    if (a_ncols >= ((size_t)1 << 30))  // the unreduced 160-bit Goldilocks sums hold 2^32 products
        return fail(ctx, SR_ERR_INVALID, "inner dimension must be below 2^30");
    if (!a_rows || !out_rows || (m_nrows && !m_rows)) return fail(ctx, SR_ERR_INVALID, "null row table");
    if (loc != SR_DEVICE && loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    for (size_t i = 0; i < a_nrows; i++)
        if ((a_ncols && !a_rows[i]) || !out_rows[i]) return fail(ctx, SR_ERR_INVALID, "null row pointer");
    for (size_t k = 0; k < m_nrows; k++)
        if (!m_rows[k]) return fail(ctx, SR_ERR_INVALID, "null row pointer");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (loc == SR_DEVICE) ? ctx->stream : ctx->own_stream;
    DevTemps tmp(ctx);
    std::vector<const u64*> ha(a_rows, a_rows + a_nrows), hm(m_rows, m_rows + m_nrows);
    std::vector<u64*> ho(out_rows, out_rows + a_nrows);
    if (loc == SR_HOST) {
        for (size_t i = 0; i < a_nrows; i++) {
            CU(tmp.upload((void**)&ha[i], a_rows[i], a_ncols * w * 8, st));
            CU(tmp.alloc((void**)&ho[i], m_ncols * w * 8));
        }
        for (size_t k = 0; k < m_nrows; k++) CU(tmp.upload((void**)&hm[k], m_rows[k], m_ncols * w * 8, st));
    } else {
        for (size_t i = 0; i < a_nrows; i++)
            if (!aligned16(ha[i]) || !aligned16(ho[i])) return fail(ctx, SR_ERR_INVALID, "row pointer misaligned");
        for (size_t k = 0; k < m_nrows; k++)
            if (!aligned16(hm[k])) return fail(ctx, SR_ERR_INVALID, "row pointer misaligned");
    }
    void *da = nullptr, *dm = nullptr, *dout = nullptr;
    CU(tmp.upload(&da, ha.data(), a_nrows * sizeof(void*), st));
    CU(tmp.upload(&dm, hm.data(), m_nrows * sizeof(void*), st));
    CU(tmp.upload(&dout, ho.data(), a_nrows * sizeof(void*), st));
    CU(sr::matmat_launch(ring, (const u64* const*)da, (const u64* const*)dm, (u64* const*)dout, a_nrows, a_ncols,
                         m_ncols, st));
    ctx->launches += (a_nrows + 65534) / 65535;
    if (loc == SR_HOST)
        for (size_t i = 0; i < a_nrows; i++)
            CU(cudaMemcpyAsync(out_rows[i], ho[i], m_ncols * w * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));  // the pointer tables are temporaries of this call
    return SR_OK;
}

int scale_impl(sr_ctx* ctx, int ring, u64* a, size_t n_limbs, const u64* r, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (n_limbs % w != 0) return fail(ctx, SR_ERR_BAD_LENGTH, "slice length is not a whole number of elements");
    const size_t n = n_limbs / w;
    if (n == 0) return SR_OK;
    if (!a || !r) return fail(ctx, SR_ERR_INVALID, "null buffer");
    CU(cudaSetDevice(ctx->device));
    if (loc == SR_DEVICE) {
        if (!aligned16(a) || !aligned16(r)) return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
        CU(sr::scale_launch(ring, a, r, n, ctx->stream));
        ctx->launches++;
        return SR_OK;
    }
    if (loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    cudaStream_t st = ctx->own_stream;
    DevTemps tmp(ctx);
    u64 *da = nullptr, *dr = nullptr;
    CU(tmp.upload((void**)&da, a, n_limbs * 8, st));
    CU(tmp.upload((void**)&dr, r, w * 8, st));
    CU(sr::scale_launch(ring, da, dr, n, st));
    ctx->launches++;
    CU(cudaMemcpyAsync(a, da, n_limbs * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SR_OK;
}

size_t fe_bytes(int ring) { return ring == SR_GOLDILOCKS ? 8 : ring == SR_BABYBEAR ? 4 : ring == SR_STARK ? 32 : 0; }

// op 0: serialize (in = limbs, out = bytes); op 1: deserialize (in = bytes, out = limbs)
int serial_impl(sr_ctx* ctx, int ring, int op, const void* in, size_t in_len, void* out, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring), fb = fe_bytes(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    const size_t N = (ring == SR_STARK) ? 4 : 1, D = w / N;
    const size_t per_elem = (op == 0) ? w : D * fb;  // limbs in / bytes in
    if (in_len % per_elem != 0)
        return fail(ctx, SR_ERR_BAD_LENGTH, "input length is not a whole number of ring elements");
    const size_t n = in_len / per_elem, nfe = n * D;
    if (n == 0) return SR_OK;
    if (!in || !out || in == out) return fail(ctx, SR_ERR_INVALID, "null or aliased buffer");
    if (loc != SR_DEVICE && loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (loc == SR_DEVICE) ? ctx->stream : ctx->own_stream;
    const size_t in_bytes = (op == 0) ? in_len * 8 : in_len, out_bytes = (op == 0) ? nfe * fb : n * w * 8;
    DevTemps tmp(ctx);
    int* dbad = nullptr;
    const void* kin = in;
    void* kout = out;
    if (op == 1) {
        int rcf = flag_begin(ctx, st);
        if (rcf) return rcf;
        dbad = ctx->dflag;
    }
    if (loc == SR_HOST) {
        CU(tmp.upload((void**)&kin, in, in_bytes, st));
        CU(tmp.alloc(&kout, out_bytes));
    } else if (!aligned16(in) || !aligned16(out)) {
        return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
    }
    CU(sr::serial_launch(ring, op, kin, kout, nfe, dbad, st));
    ctx->launches++;
    if (loc == SR_HOST) CU(cudaMemcpyAsync(out, kout, out_bytes, cudaMemcpyDeviceToHost, st));
    int bad = 0;
    if (op == 1) {
        int rcf = flag_end(ctx, st, &bad);
        if (rcf) return rcf;
    } else if (loc == SR_HOST) {
        CU(cudaStreamSynchronize(st));
    }
    if (bad) return fail(ctx, SR_ERR_INVALID, "InvalidData: a serialized integer is not below the modulus");
    return SR_OK;
}

}  // namespace

extern "C" {

const char* sr_version(void) { return "stark-rings-b200 0.1 (sm_100a)"; }

size_t sr_elem_limbs(int ring) { return elem_limbs(ring); }

int sr_init(int device, sr_ctx** out) {
    if (!out) return SR_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return SR_ERR_CUDA;  // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return SR_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SR_ERR_CUDA;
    if (prop.major != 10) {
        fprintf(stderr, "stark-rings-b200: device %d is sm_%d%d; this library carries sm_100a code only\n", device,
                prop.major, prop.minor);
        return SR_ERR_CUDA;
    }
    sr_ctx* ctx = new sr_ctx();
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&ctx->t0) == cudaSuccess && cudaEventCreate(&ctx->t1) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_stream, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < sr_ctx::NBUF; i++)
        ok = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        delete ctx;
        return SR_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SR_OK;
}

int sr_destroy(sr_ctx* ctx) {
    if (!ctx) return SR_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < sr_ctx::NBUF; i++) {
        for (int j = 0; j < 3; j++)
            if (ctx->stage[i][j]) cudaFree(ctx->stage[i][j]);
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->mv_scratch) cudaFree(ctx->mv_scratch);
    if (ctx->mv_rows) cudaFree(ctx->mv_rows);
    if (ctx->commit_counters) cudaFree(ctx->commit_counters);
    for (auto& slot : ctx->pool)
        if (slot.first) cudaFree(slot.first);
    if (ctx->ev_stream) cudaEventDestroy(ctx->ev_stream);
    if (ctx->dflag) cudaFree(ctx->dflag);
    if (ctx->hflag) cudaFreeHost(ctx->hflag);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
    return SR_OK;
}

const char* sr_last_error(sr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// The context owns single scratch buffers (mat-vec partials, row table, tickets): when the stream changes, the new
// stream is ordered after the work already enqueued on the old one, so that two calls can never overlap on them.
static int switch_stream(sr_ctx* ctx, cudaStream_t next) {
    if (next == ctx->stream) return SR_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(next, &cs) != cudaSuccess) cs = cudaStreamCaptureStatusNone;
    cudaStreamCaptureStatus co = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &co) != cudaSuccess) co = cudaStreamCaptureStatusNone;
    cudaGetLastError();
    if (cs == cudaStreamCaptureStatusNone && co == cudaStreamCaptureStatusNone) {  // (never tie a capture to outside work)
        CU(cudaSetDevice(ctx->device));
        CU(cudaEventRecord(ctx->ev_stream, ctx->stream));
        CU(cudaStreamWaitEvent(next, ctx->ev_stream, 0));
    }
    ctx->stream = next;
    return SR_OK;
}

int sr_set_stream(sr_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return switch_stream(ctx, (cudaStream_t)cuda_stream);  // NULL is CUDA's (legacy) default stream
}

int sr_set_pipelined(sr_ctx* ctx, int on) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->pipelined = on != 0;
    return SR_OK;
}

int sr_reset_stream(sr_ctx* ctx) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return switch_stream(ctx, ctx->own_stream);
}

int sr_sync(sr_ctx* ctx) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream != ctx->own_stream) CU(cudaStreamSynchronize(ctx->own_stream));
    return SR_OK;
}

uint64_t sr_kernel_launches(sr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sr_dev_alloc(sr_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
    if (e == cudaErrorMemoryAllocation) return fail(ctx, SR_ERR_NOMEM, "cudaMalloc: out of device memory");
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
    return SR_OK;
}
int sr_dev_free(sr_ctx* ctx, void* dptr) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(dptr));
    return SR_OK;
}
int sr_host_alloc(sr_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 16);
    if (e == cudaErrorMemoryAllocation) return fail(ctx, SR_ERR_NOMEM, "cudaMallocHost: out of memory");
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMallocHost");
    return SR_OK;
}
int sr_host_free(sr_ctx* ctx, void* hptr) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaFreeHost(hptr));
    return SR_OK;
}
int sr_h2d(sr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SR_OK;
}
int sr_d2h(sr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SR_OK;
}

int sr_timer_start(sr_ctx* ctx) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->t0, ctx->stream));
    return SR_OK;
}
int sr_timer_stop(sr_ctx* ctx, float* ms) {
    if (!ctx || !ms) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->t1, ctx->stream));
    CU(cudaEventSynchronize(ctx->t1));
    CU(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return SR_OK;
}

int sr_imad_peak(sr_ctx* ctx, double* tops3) {
    if (!ctx || !tops3) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(sr::imad_peak_measure(ctx->sms, ctx->stream, tops3));
    ctx->launches += 18;
    return SR_OK;
}

int sr_crt_batch(sr_ctx* ctx, int ring, uint64_t* buf, size_t n_limbs, int loc) {
    return batch(ctx, ring, sr::OP_CRT, buf, nullptr, buf, n_limbs, loc);
}
int sr_icrt_batch(sr_ctx* ctx, int ring, uint64_t* buf, size_t n_limbs, int loc) {
    return batch(ctx, ring, sr::OP_ICRT, buf, nullptr, buf, n_limbs, loc);
}
int sr_ntt_mul_batch(sr_ctx* ctx, int ring, uint64_t* a, const uint64_t* b, size_t n_limbs, int loc) {
    return batch(ctx, ring, sr::OP_NTT_MUL, a, b, a, n_limbs, loc);
}
int sr_ring_mul_batch(sr_ctx* ctx, int ring, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n_limbs,
                      int loc) {
    return batch(ctx, ring, sr::OP_RING_MUL, a, b, out, n_limbs, loc);
}

int sr_matvec(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols, const uint64_t* v,
              size_t v_limbs, uint64_t* out, int loc) {
    return matvec_impl(ctx, ring, rows, nrows, ncols, v, v_limbs, out, loc);
}
int sr_matvec_partial(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols,
                      const uint64_t* v, size_t v_limbs, uint64_t* partial_out, int loc) {
    return matvec_impl(ctx, ring, rows, nrows, ncols, v, v_limbs, partial_out, loc);
}
int sr_modsum_partials(sr_ctx* ctx, int ring, const uint64_t* gathered, size_t nranks, size_t nrows, uint64_t* out,
                       int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (nrows == 0) return SR_OK;
    if (!gathered || !out || nranks == 0) return fail(ctx, SR_ERR_INVALID, "null buffer / zero ranks");
    CU(cudaSetDevice(ctx->device));
    if (loc == SR_DEVICE) {
        CU(sr::modsum_launch(ring, gathered, nranks, nrows, nrows, out, ctx->stream, nullptr));
        ctx->launches++;
        return SR_OK;
    }
    if (loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    void *dg = nullptr, *dout = nullptr;
    const size_t gbytes = nranks * nrows * w * 8, obytes = nrows * w * 8;
    CU(pool_get(ctx, 0, gbytes, &dg));
    CU(pool_get(ctx, 1, obytes, &dout));
    CU(cudaMemcpyAsync(dg, gathered, gbytes, cudaMemcpyHostToDevice, ctx->own_stream));
    CU(sr::modsum_launch(ring, (const u64*)dg, nranks, nrows, nrows, (u64*)dout, ctx->own_stream, nullptr));
    CU(cudaMemcpyAsync(out, dout, obytes, cudaMemcpyDeviceToHost, ctx->own_stream));
    CU(cudaStreamSynchronize(ctx->own_stream));
    ctx->launches++;
    return SR_OK;
}

static int coeff_impl(sr_ctx* ctx, int ring, int op, const uint64_t* in, size_t in_limbs, size_t len, uint64_t* out,
                      int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    const size_t N = (ring == SR_STARK) ? 4 : 1, D = w / N;
    if (op == 1) len = D;
    if (len < D || len > 2 * D) return fail(ctx, SR_ERR_BAD_LENGTH, "coefficients per polynomial must be in [D, 2D]");
    if (in_limbs % (len * N) != 0) return fail(ctx, SR_ERR_BAD_LENGTH, "slice length is not a whole number of polynomials");
    const size_t n = in_limbs / (len * N);
    if (n == 0) return SR_OK;
    if (!in || !out) return fail(ctx, SR_ERR_INVALID, "null buffer");
    if (in == out) return fail(ctx, SR_ERR_INVALID, "reduce / rot are out of place: out must differ from in");
    CU(cudaSetDevice(ctx->device));
    if (loc == SR_DEVICE) {
        if (!aligned16(in) || !aligned16(out)) return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
        CU(sr::coeff_launch(ring, op, in, out, n, (int)len, ctx->stream));
        ctx->launches++;
        return SR_OK;
    }
    if (loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    void *din = nullptr, *dout = nullptr;
    const size_t ib = in_limbs * 8, ob = n * w * 8;
    CU(pool_get(ctx, 0, ib, &din));
    CU(pool_get(ctx, 1, ob, &dout));
    CU(cudaMemcpyAsync(din, in, ib, cudaMemcpyHostToDevice, ctx->own_stream));
    CU(sr::coeff_launch(ring, op, (const u64*)din, (u64*)dout, n, (int)len, ctx->own_stream));
    CU(cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, ctx->own_stream));
    CU(cudaStreamSynchronize(ctx->own_stream));
    ctx->launches++;
    return SR_OK;
}

static int addsub_impl(sr_ctx* ctx, int ring, int op, uint64_t* a, const uint64_t* b, size_t n_limbs, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (n_limbs % w != 0) return fail(ctx, SR_ERR_BAD_LENGTH, "slice length is not a whole number of elements");
    const size_t n = n_limbs / w;
    if (n == 0) return SR_OK;
    if (!a || (op != 2 && !b)) return fail(ctx, SR_ERR_INVALID, "null buffer");
    CU(cudaSetDevice(ctx->device));
    if (loc == SR_DEVICE) {
        if (!aligned16(a) || (op != 2 && !aligned16(b)))
            return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
        CU(sr::addsub_launch(ring, op, a, b, n, ctx->stream));
        ctx->launches++;
        return SR_OK;
    }
    if (loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    cudaStream_t st = ctx->own_stream;
    DevTemps tmp(ctx);
    u64 *da = nullptr, *db = nullptr;
    CU(tmp.upload((void**)&da, a, n_limbs * 8, st));
    if (op != 2) CU(tmp.upload((void**)&db, b, n_limbs * 8, st));
    CU(sr::addsub_launch(ring, op, da, db, n, st));
    ctx->launches++;
    CU(cudaMemcpyAsync(a, da, n_limbs * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SR_OK;
}
int sr_add_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, const uint64_t* b, size_t n_limbs, int loc) {
    return addsub_impl(ctx, ring, 0, a_inout, b, n_limbs, loc);
}
int sr_sub_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, const uint64_t* b, size_t n_limbs, int loc) {
    return addsub_impl(ctx, ring, 1, a_inout, b, n_limbs, loc);
}
int sr_neg_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, size_t n_limbs, int loc) {
    return addsub_impl(ctx, ring, 2, a_inout, nullptr, n_limbs, loc);
}
// Sum of a slice of ring elements (ntt_form.rs:640-654: fold from ZERO with Add): the modular sum kernel of the
// column-sharded commitment with one "row" and n "ranks".
int sr_sum_batch(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint64_t* out, int loc) {
    const size_t w = elem_limbs(ring);
    if (!ctx) return SR_ERR_INVALID;
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (n_limbs % w != 0) return fail(ctx, SR_ERR_BAD_LENGTH, "slice length is not a whole number of elements");
    if (!out) return fail(ctx, SR_ERR_INVALID, "null buffer");
    if (n_limbs == 0) {  // Sum of nothing = ZERO
        if (loc == SR_HOST) { memset(out, 0, w * 8); return SR_OK; }
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaSetDevice(ctx->device));
        CU(cudaMemsetAsync(out, 0, w * 8, ctx->stream));
        return SR_OK;
    }
    return sr_modsum_partials(ctx, ring, in, n_limbs / w, 1, out, loc);
}

int sr_reduce_batch(sr_ctx* ctx, int ring, const uint64_t* in, size_t in_limbs, size_t coeffs_per_poly, uint64_t* out,
                    int loc) {
    return coeff_impl(ctx, ring, 0, in, in_limbs, coeffs_per_poly, out, loc);
}
int sr_rot_batch(sr_ctx* ctx, int ring, const uint64_t* in, uint64_t* out, size_t n_limbs, int loc) {
    return coeff_impl(ctx, ring, 1, in, n_limbs, 0, out, loc);
}

static int decomp_impl(sr_ctx* ctx, int ring, int op, const uint64_t* in, size_t in_limbs, uint64_t b_lo, uint64_t b_hi,
                       size_t pad, uint64_t* out, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ring != SR_GOLDILOCKS && ring != SR_BABYBEAR)
        return fail(ctx, SR_ERR_INVALID, "gadget (de/re)composition is implemented for the Fp64 rings only");
    // decompose_balanced_in_place asserts: basis not 0 or 1, and even (mod.rs:63-69)
    if (b_hi != 0 || b_lo < 2 || (b_lo & 1) || b_lo >= (1ull << 62))
        return fail(ctx, SR_ERR_INVALID, "decomposition basis must be even, >= 2 and < 2^62");
    if (pad == 0 || pad > (1u << 20)) return fail(ctx, SR_ERR_INVALID, "padding size out of range");
    const size_t w = elem_limbs(ring);
    const size_t in_per = (op == 0) ? w : w * pad;
    if (in_limbs % in_per != 0) return fail(ctx, SR_ERR_BAD_LENGTH, "slice length is not a whole number of elements");
    const size_t n = in_limbs / in_per;
    if (n == 0) return SR_OK;
    if (!in || !out || in == out) return fail(ctx, SR_ERR_INVALID, "null or aliased buffer");
    const uint64_t p = (ring == SR_GOLDILOCKS) ? 0xFFFFFFFF00000001ull : 2013265921ull;
    const uint64_t b_std = b_lo % p;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (loc == SR_DEVICE) ? ctx->stream : ctx->own_stream;
    if (loc != SR_DEVICE && loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    int rcf = flag_begin(ctx, st);
    if (rcf) return rcf;
    DevTemps tmp(ctx);
    const size_t ib = in_limbs * 8, ob = (op == 0 ? n * pad : n) * w * 8;
    const u64* kin = in;
    u64* kout = out;
    if (loc == SR_HOST) {
        CU(tmp.upload((void**)&kin, in, ib, st));
        CU(tmp.alloc((void**)&kout, ob));
    }
    CU(sr::decomp_launch(ring, op, kin, kout, n, b_lo, b_std, (int)pad, ctx->dflag, st));
    ctx->launches++;
    if (loc == SR_HOST) CU(cudaMemcpyAsync(out, kout, ob, cudaMemcpyDeviceToHost, st));
    int flag = 0;
    rcf = flag_end(ctx, st, &flag);  // the overflow flag makes this call synchronous
    if (rcf) return rcf;
    if (flag) return fail(ctx, SR_ERR_BAD_LENGTH, "padding_size too small for the decomposition (the reference panics)");
    return SR_OK;
}

int sr_gadget_decompose(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint64_t b_lo, uint64_t b_hi,
                        size_t padding_size, uint64_t* out, int loc) {
    return decomp_impl(ctx, ring, 0, in, n_limbs, b_lo, b_hi, padding_size, out, loc);
}
int sr_gadget_recompose(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint64_t b_lo, uint64_t b_hi,
                        size_t padding_size, uint64_t* out, int loc) {
    return decomp_impl(ctx, ring, 1, in, n_limbs, b_lo, b_hi, padding_size, out, loc);
}

int sr_sparse_matvec(sr_ctx* ctx, int ring, size_t nrows, size_t ncols, const uint64_t* row_ptr,
                     const uint64_t* col_idx, const uint64_t* vals, const uint64_t* v, size_t v_limbs, uint64_t* out,
                     int loc) {
    return sparse_matvec_impl(ctx, ring, nrows, ncols, row_ptr, col_idx, vals, v, v_limbs, out, loc);
}
int sr_matmat(sr_ctx* ctx, int ring, const uint64_t* const* a_rows, size_t a_nrows, size_t a_ncols,
              const uint64_t* const* m_rows, size_t m_nrows, size_t m_ncols, uint64_t* const* out_rows, int loc) {
    return matmat_impl(ctx, ring, a_rows, a_nrows, a_ncols, m_rows, m_nrows, m_ncols, out_rows, loc);
}
int sr_ntt_scale_batch(sr_ctx* ctx, int ring, uint64_t* a_inout, size_t n_limbs, const uint64_t* r, int loc) {
    return scale_impl(ctx, ring, a_inout, n_limbs, r, loc);
}

/* ---- sparse x sparse product (SparseMatrix::checked_mul_mat, sparse_matrix.rs:219-275) ---------------------------
 * Symbolic phase on the host, exactly the reference's: the columns of M are gathered in row order (:224-229), and for
 * every (row i of A, column j of M) the two index lists are merge-joined in their STORED order (:235-262; sorted
 * lists pair equal indices, unsorted ones skip entries just as the reference does).  Every (i, j) with at least one
 * match is a candidate; whether it becomes an entry is decided by the values (a match whose product is the zero
 * element does not count, :249-256), i.e. by the numeric phase. */
int sr_sparse_matmat_symbolic(size_t a_nrows, const uint64_t* a_row_ptr, const uint64_t* a_col_idx, size_t m_nrows,
                              size_t m_ncols, const uint64_t* m_row_ptr, const uint64_t* m_col_idx, size_t* ncand,
                              size_t* npairs, uint64_t* cand_row, uint64_t* cand_col, uint64_t* pair_ptr,
                              uint64_t* pair_a, uint64_t* pair_m) {
    if (!a_row_ptr || !m_row_ptr || !ncand || !npairs) return SR_ERR_INVALID;
    const size_t a_nnz = a_row_ptr[a_nrows], m_nnz = m_row_ptr[m_nrows];
    if ((a_nnz && !a_col_idx) || (m_nnz && !m_col_idx)) return SR_ERR_INVALID;
    // m_cols[j] = entries of column j in row order, as (entry index, row index)
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> m_cols(m_ncols);
    for (size_t r = 0; r < m_nrows; r++) {
        if (m_row_ptr[r] > m_row_ptr[r + 1]) return SR_ERR_INVALID;
        for (uint64_t e = m_row_ptr[r]; e < m_row_ptr[r + 1]; e++) {
            if (m_col_idx[e] >= m_ncols) return SR_ERR_INVALID;  // the reference panics on m_cols[*col_idx]
            m_cols[m_col_idx[e]].push_back(std::make_pair(e, (uint64_t)r));
        }
    }
    const bool fill = cand_row && cand_col && pair_ptr && pair_a && pair_m;
    size_t nc = 0, np = 0;
    for (size_t i = 0; i < a_nrows; i++) {
        if (a_row_ptr[i] > a_row_ptr[i + 1]) return SR_ERR_INVALID;
        for (size_t j = 0; j < m_ncols; j++) {
            const auto& col = m_cols[j];
            uint64_t ra = a_row_ptr[i];
            size_t ci = 0;
            const size_t np0 = np;
            while (ra < a_row_ptr[i + 1] && ci < col.size()) {
                const uint64_t r_idx = a_col_idx[ra], c_idx = col[ci].second;
                if (r_idx < c_idx) ra++;
                else if (r_idx > c_idx) ci++;
                else {
                    if (fill) { pair_a[np] = ra; pair_m[np] = col[ci].first; }
                    np++;
                    ra++;
                    ci++;
                }
            }
            if (np > np0) {
                if (fill) { cand_row[nc] = i; cand_col[nc] = j; pair_ptr[nc] = np0; }
                nc++;
            }
        }
    }
    if (fill) pair_ptr[nc] = np;
    *ncand = nc;
    *npairs = np;
    return SR_OK;
}

/* Numeric phase: out_vals[c] = sum over the candidate's pairs of a_vals[pair_a] * m_vals[pair_m];
 * nonzero[c] = 1 iff some product is not the zero element (host array, always). */
int sr_sparse_matmat_values(sr_ctx* ctx, int ring, const uint64_t* a_vals, const uint64_t* m_vals, size_t ncand,
                            const uint64_t* pair_ptr, const uint64_t* pair_a, const uint64_t* pair_m,
                            uint64_t* out_vals, int* nonzero_host, int loc) {
    if (!ctx) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0) return fail(ctx, SR_ERR_INVALID, "unknown ring id");
    if (ncand == 0) return SR_OK;
    if (!a_vals || !m_vals || !pair_ptr || !pair_a || !pair_m || !out_vals || !nonzero_host)
        return fail(ctx, SR_ERR_INVALID, "null buffer");
    if (loc != SR_DEVICE && loc != SR_HOST) return fail(ctx, SR_ERR_INVALID, "unknown loc");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (loc == SR_DEVICE) ? ctx->stream : ctx->own_stream;
    const size_t npairs = (size_t)pair_ptr[ncand];  // the index arrays are host arrays in both modes
    DevTemps tmp(ctx);
    u64 *d_pp = nullptr, *d_pa = nullptr, *d_pm = nullptr;
    int* d_nz = nullptr;
    CU(tmp.upload((void**)&d_pp, pair_ptr, (ncand + 1) * 8, st));
    CU(tmp.upload((void**)&d_pa, pair_a, npairs * 8, st));
    CU(tmp.upload((void**)&d_pm, pair_m, npairs * 8, st));
    CU(tmp.alloc((void**)&d_nz, ncand * sizeof(int)));
    CU(cudaMemsetAsync(d_nz, 0, ncand * sizeof(int), st));
    const u64 *k_a = a_vals, *k_m = m_vals;
    u64* k_out = out_vals;
    if (loc == SR_HOST) {
        size_t na = 0, nm = 0;
        for (size_t e = 0; e < npairs; e++) {
            if (pair_a[e] + 1 > na) na = pair_a[e] + 1;
            if (pair_m[e] + 1 > nm) nm = pair_m[e] + 1;
        }
        CU(tmp.upload((void**)&k_a, a_vals, na * w * 8, st));
        CU(tmp.upload((void**)&k_m, m_vals, nm * w * 8, st));
        CU(tmp.alloc((void**)&k_out, ncand * w * 8));
    } else if (!aligned16(a_vals) || !aligned16(m_vals) || !aligned16(out_vals)) {
        return fail(ctx, SR_ERR_INVALID, "device buffers must be 16-byte aligned");
    }
    CU(sr::sparse_pairs_launch(ring, k_a, k_m, d_pp, d_pa, d_pm, ncand, k_out, d_nz, st));
    ctx->launches++;
    if (loc == SR_HOST) CU(cudaMemcpyAsync(out_vals, k_out, ncand * w * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(nonzero_host, d_nz, ncand * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SR_OK;
}

/* ---- mailbox: column-sharded commitment over NVLink peer memory -------------------------------------------- */
static size_t mailbox_bytes(size_t w, size_t nrows_max, int nranks) {
    return sr_mailbox::HEADER + (size_t)sr::MAILBOX_DEPTH * nranks * nrows_max * w * 8;
}

int sr_mailbox_create(sr_ctx* ctx, int ring, size_t nrows_max, int nranks, sr_mailbox** out, unsigned char* handle_out) {
    if (!ctx || !out) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0 || nrows_max == 0 || nranks < 1 || nranks > sr_mailbox::MAX_RANKS)
        return fail(ctx, SR_ERR_INVALID, "mailbox: bad ring / nrows_max / nranks");
    CU(cudaSetDevice(ctx->device));
    sr_mailbox* mb = new sr_mailbox();
    mb->ring = ring; mb->nrows_max = nrows_max; mb->nranks = nranks; mb->owner = true;
    mb->bytes = mailbox_bytes(w, nrows_max, nranks);
    cudaError_t e = cudaMalloc(&mb->base, mb->bytes);
    if (e == cudaSuccess) e = cudaMemset(mb->base, 0, mb->bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&mb->sent, 8);
    if (e == cudaSuccess) e = cudaMemset(mb->sent, 0, 8);
    if (e == cudaSuccess && handle_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, mb->base);
        if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
    }
    if (e != cudaSuccess) {
        if (mb->base) cudaFree(mb->base);
        if (mb->sent) cudaFree(mb->sent);
        delete mb;
        return cuda_fail(ctx, e, "sr_mailbox_create");
    }
    *out = mb;
    return SR_OK;
}

int sr_mailbox_open(sr_ctx* ctx, int ring, size_t nrows_max, int nranks, const unsigned char* handle, sr_mailbox** out) {
    if (!ctx || !out || !handle) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t w = elem_limbs(ring);
    if (w == 0 || nrows_max == 0 || nranks < 1 || nranks > sr_mailbox::MAX_RANKS)
        return fail(ctx, SR_ERR_INVALID, "mailbox: bad ring / nrows_max / nranks");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    u64* sent = nullptr;
    cudaError_t e = cudaMalloc((void**)&sent, 8);
    if (e == cudaSuccess) e = cudaMemset(sent, 0, 8);
    if (e != cudaSuccess) {
        cudaIpcCloseMemHandle(base);
        if (sent) cudaFree(sent);
        return cuda_fail(ctx, e, "sr_mailbox_open");
    }
    sr_mailbox* mb = new sr_mailbox();
    mb->sent = sent;
    mb->ring = ring; mb->nrows_max = nrows_max; mb->nranks = nranks; mb->owner = false;
    mb->base = base;
    mb->bytes = mailbox_bytes(w, nrows_max, nranks);
    *out = mb;
    return SR_OK;
}

int sr_mailbox_destroy(sr_ctx* ctx, sr_mailbox* mb) {
    if (!ctx || !mb) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    if (mb->sent) cudaFree(mb->sent);
    if (mb->base) {
        if (mb->owner) CU(cudaFree(mb->base));
        else CU(cudaIpcCloseMemHandle(mb->base));
    }
    delete mb;
    return SR_OK;
}

int sr_mailbox_error(sr_ctx* ctx, sr_mailbox* mb, int* timed_out) {
    if (!ctx || !mb || !timed_out) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(timed_out, mb->err(), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SR_OK;
}

int sr_mailbox_set_timeout(sr_ctx* ctx, sr_mailbox* mb, uint64_t nanoseconds) {
    if (!ctx || !mb || nanoseconds == 0) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    mb->timeout_ns = nanoseconds;
    return SR_OK;
}

// role 1 (writer) or 3 (root, fused): this rank's share of the product, delivered through the mailbox
static int commit_product(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols,
                          const uint64_t* v, size_t v_limbs, sr_mailbox* box, int rank, uint64_t epoch, int role,
                          uint64_t* out) {
    if (!ctx || !box) return SR_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ring != box->ring || nrows == 0 || nrows > box->nrows_max || rank < 0 || rank >= box->nranks)
            return fail(ctx, SR_ERR_INVALID, "commit: ring / nrows / rank do not fit the mailbox");
        if (role == 3 && (!box->owner || !out))
            return fail(ctx, SR_ERR_INVALID, "sr_commit_root: needs the mailbox this rank created and an output buffer");
    }
    sr::PeerSync ps = {};
    ps.role = role;
    ps.nranks = box->nranks;
    ps.rank = rank;
    ps.slots = box->slots();
    ps.slot_stride = box->nrows_max * elem_limbs(ring);
    ps.flags = box->flags();
    ps.consumed = box->consumed();
    ps.epoch = epoch;
    ps.epoch_ctr = box->sent;
    ps.err = box->err();
    ps.timeout_ns = box->timeout_ns;
    return matvec_impl(ctx, ring, rows, nrows, ncols, v, v_limbs, role == 3 ? out : box->slots(), SR_DEVICE, &ps);
}

int sr_commit_send(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols, const uint64_t* v,
                   size_t v_limbs, sr_mailbox* root_box, int rank, uint64_t epoch) {
    return commit_product(ctx, ring, rows, nrows, ncols, v, v_limbs, root_box, rank, epoch, 1, nullptr);
}

int sr_commit_root(sr_ctx* ctx, int ring, const uint64_t* const* rows, size_t nrows, size_t ncols, const uint64_t* v,
                   size_t v_limbs, sr_mailbox* own_box, int rank, uint64_t epoch, uint64_t* out) {
    return commit_product(ctx, ring, rows, nrows, ncols, v, v_limbs, own_box, rank, epoch, 3, out);
}

int sr_commit_reduce(sr_ctx* ctx, int ring, sr_mailbox* own_box, size_t nrows, uint64_t epoch, uint64_t* out) {
    if (!ctx || !own_box || !out) return SR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!own_box->owner || ring != own_box->ring || nrows == 0 || nrows > own_box->nrows_max)
        return fail(ctx, SR_ERR_INVALID, "sr_commit_reduce: needs the mailbox this rank created, matching ring / nrows");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_counters(ctx);
    if (rc) return rc;
    sr::PeerSync ps = {};
    ps.role = 2;
    ps.nranks = own_box->nranks;
    ps.slots = own_box->slots();
    ps.slot_stride = own_box->nrows_max * elem_limbs(ring);
    ps.flags = own_box->flags();
    ps.consumed = own_box->consumed();
    ps.epoch = epoch;
    ps.epoch_ctr = own_box->consumed();  // device-resident epoch of the root = last epoch summed + 1
    ps.counter = ctx->commit_counters + 1;
    ps.err = own_box->err();
    ps.timeout_ns = own_box->timeout_ns;
    CU(sr::modsum_launch(ring, own_box->slots(), (size_t)own_box->nranks, own_box->nrows_max, nrows, out, ctx->stream,
                         &ps));
    ctx->launches++;
    return SR_OK;
}

size_t sr_serialized_bytes(int ring, size_t n_elems) {
    const size_t w = elem_limbs(ring);
    return w ? n_elems * (w / (ring == SR_STARK ? 4 : 1)) * fe_bytes(ring) : 0;
}
int sr_serialize_batch(sr_ctx* ctx, int ring, const uint64_t* in, size_t n_limbs, uint8_t* out_bytes, int loc) {
    return serial_impl(ctx, ring, 0, in, n_limbs, out_bytes, loc);
}
int sr_deserialize_batch(sr_ctx* ctx, int ring, const uint8_t* in_bytes, size_t n_bytes, uint64_t* out, int loc) {
    return serial_impl(ctx, ring, 1, in_bytes, n_bytes, out, loc);
}

#define SR_DEFINE_RING(tag, RING)                                                                             \
    int sr_##tag##_crt_batch(sr_ctx* c, uint64_t* buf, size_t n, int loc) { return sr_crt_batch(c, RING, buf, n, loc); }   \
    int sr_##tag##_icrt_batch(sr_ctx* c, uint64_t* buf, size_t n, int loc) { return sr_icrt_batch(c, RING, buf, n, loc); } \
    int sr_##tag##_ntt_mul_batch(sr_ctx* c, uint64_t* a, const uint64_t* b, size_t n, int loc) {             \
        return sr_ntt_mul_batch(c, RING, a, b, n, loc);                                                       \
    }                                                                                                         \
    int sr_##tag##_ring_mul_batch(sr_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n,  \
                                  int loc) {                                                                  \
        return sr_ring_mul_batch(c, RING, a, b, out, n, loc);                                                 \
    }                                                                                                         \
    int sr_##tag##_matvec(sr_ctx* c, const uint64_t* const* rows, size_t nrows, size_t ncols, const uint64_t* v, \
                          size_t v_limbs, uint64_t* out, int loc) {                                           \
        return sr_matvec(c, RING, rows, nrows, ncols, v, v_limbs, out, loc);                                  \
    }

SR_DEFINE_RING(gl, SR_GOLDILOCKS)
SR_DEFINE_RING(bb, SR_BABYBEAR)
SR_DEFINE_RING(sp, SR_STARK)

}  // extern "C"
