// Ring matrix x vector product y_i = sum_j A[i][j] * v[j] over NTT-form elements
// (reference: linear_algebra/src/matrix.rs:168-178 with R = RqNTT; the inner product is
// ntt_form.rs:521-536 (slot-wise Mul) folded with ntt_form.rs:588-601 (Add) from ZERO).
//
// The product is independent per CRT slot, so the unit of work is one slot of one column.  ONE kernel per pass of
// up to four matrix rows does everything: every CTA accumulates its share of the columns, reduces per slot index
// through shared memory and leaves one partial element per row in scratch; the LAST CTA to arrive (a device-wide
// ticket) adds the partials of all CTAs in a fixed order -- so the result does not depend on which CTA that is --
// and writes the result rows.  In the column-sharded commitment the same tail stores the rank's partial
// rows straight into the root rank's mailbox over NVLink and publishes an epoch flag; on the root it then acquires
// the flags of all ranks and adds their partials mod p.  A commitment is therefore one launch per rank (per four
// rows), with no separate reduction kernel, no collective call and no host synchronisation.
#include <cuda_runtime.h>

#include <type_traits>

#define SR_GL_EPS_ON_ALU  // these kernels are bound by the multiply-add pipe: see gl_ring.cuh plus_eps_if

#include "bb_ring.cuh"
#include "gl_ring.cuh"
#include "sp_ring.cuh"
#include "sr_launch.cuh"
#include "sr_slots.cuh"
#include "sr_tma.cuh"

namespace sr {

// ---- peer-memory primitives ---------------------------------------------------------------------------------------
SR_D u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SR_D void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SR_D unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want; false (and *err = 1) when the budget runs out: a lost peer must not hang the GPU
SR_D bool spin_until(const u64* p, u64 want, int* err, unsigned long long timeout_ns) {
    if (ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(p) < want) {
        __nanosleep(100);
        if (global_ns() - t0 > timeout_ns) {
            atomicExch(err, 1);
            return false;
        }
    }
    return true;
}

// What the tail of a product kernel needs.
struct MvTail {
    u64* partial;       // scratch: [gridDim.x][nrows] partial elements
    u64* out;           // result rows (role 0: nrows elements; role 3: the final sum over all ranks)
    unsigned* counter;  // arrival ticket, zero between launches
    int last_pass;      // this launch covers the last rows of the product: publish
    PeerSync ps;
};

// Programmatic dependent launch (sm_90+): with the launch attribute set, the CTAs of the NEXT kernel on the stream may
// start once every CTA of this one has executed launch_dependents (or exited), i.e. while this kernel's last CTA is
// still in its tail; `wait` blocks until the previous kernel has completed and its writes are visible.  Without the
// attribute both are no-ops.  A pipelined launch uses one CTA slot less than the GPU has, so that the slot still
// held by the previous kernel's tail does not turn one CTA of this kernel into a straggler (the chunks are assigned
// statically; drawing them from a device-wide counter instead was measured 12% slower).
SR_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SR_D void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- the fused tail ---------------------------------------------------------------------------------------------------
// Called by every thread of the CTA once its partial rows [row0, row0 + rb) are in t.partial.
// red: shared memory for blockDim.x values; sflag: two shared ints.
template <class S>
SR_D void mv_tail(const MvTail& t, size_t nrows, size_t row0, int rb, typename S::Val* red, int* sflag) {
    typedef typename S::Val Val;
    const int T = (int)blockDim.x, tid = (int)threadIdx.x;
    __threadfence();  // this CTA's partials before its ticket
    __syncthreads();
    if (tid == 0) sflag[0] = (atomicAdd(t.counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!sflag[0]) return;
    __threadfence();
    const PeerSync& ps = t.ps;
    const bool mailbox = (ps.role == 1 || ps.role == 3);
    u64 epoch = 0;
    size_t slot0 = 0;
    u64* dst = t.out;
    if (mailbox) {
        epoch = ps.epoch ? ps.epoch : *reinterpret_cast<volatile u64*>(ps.epoch_ctr) + 1;
        slot0 = (size_t)(epoch % MAILBOX_DEPTH) * ps.nranks;
        dst = ps.slots + (slot0 + ps.rank) * ps.slot_stride;
    }
    // Slot reuse: the root must have summed the epoch that used this mailbox slot before.  One peer load, ISSUED here
    // and only looked at after the summation below, so that its NVLink round trip is off the critical path.
    u64 consumed_now = ~0ull;
    const bool check_slot = mailbox && epoch > (u64)MAILBOX_DEPTH && tid == T - 1;
    if (check_slot) consumed_now = ld_acquire_sys(ps.consumed);
    // sum the gridDim.x partials of every (row, slot) of this pass: SUB threads per unit, each over a strided
    // subset (four independent loads in flight at a time), then the SUB sums in order.  Fixed order for a given
    // grid: deterministic.
    const int U = rb * S::SLOTS;
    int SUB = T / U;
    if (SUB > 32) SUB = 32;
    const int G = (int)gridDim.x;
    if (tid < U * SUB) {
        const int u = tid / SUB, sub = tid - u * SUB;
        const int r = u / S::SLOTS, slot = u - r * S::SLOTS;
        const u64* p = t.partial + (row0 + r) * S::ELEM_U64 + slot * S::SLOT_U64;
        const size_t stride = nrows * S::ELEM_U64;
        Val s = S::zero();
        int k = sub;
        for (; k + 3 * SUB < G; k += 4 * SUB) {
            const Val v0 = S::load_cv(p + (size_t)k * stride), v1 = S::load_cv(p + (size_t)(k + SUB) * stride);
            const Val v2 = S::load_cv(p + (size_t)(k + 2 * SUB) * stride), v3 = S::load_cv(p + (size_t)(k + 3 * SUB) * stride);
            S::acc(s, v0); S::acc(s, v1); S::acc(s, v2); S::acc(s, v3);
        }
        for (; k < G; k += SUB) S::acc(s, S::load_cv(p + (size_t)k * stride));
        red[tid] = s;
    }
    if (tid == T - 1) {
        int ok = 1;
        if (check_slot && consumed_now < epoch - MAILBOX_DEPTH)
            ok = spin_until(ps.consumed, epoch - MAILBOX_DEPTH, ps.err, ps.timeout_ns) ? 1 : 0;
        sflag[1] = ok;
    }
    __syncthreads();
    const int ok = sflag[1];
    if (tid < U * SUB && (tid % SUB) == 0 && ok) {
        Val s = red[tid];
        for (int k = 1; k < SUB; k++) S::acc(s, red[tid + k]);
        const int u = tid / SUB, r = u / S::SLOTS, slot = u - r * S::SLOTS;
        S::store(dst + (row0 + r) * S::ELEM_U64 + slot * S::SLOT_U64, s);
    }
    if (!mailbox) {
        if (tid == 0) *t.counter = 0;  // ready for the next launch on this stream
        return;
    }
    __threadfence_system();  // the result stores (possibly to peer memory) before the flag
    __syncthreads();
    if (tid == 0) {
        *t.counter = 0;
        if (t.last_pass && ok) {
            *ps.epoch_ctr = epoch;
            __threadfence_system();
            st_release_sys(ps.flags + ps.rank, epoch);
        }
    }
    if (ps.role != 3 || !t.last_pass) return;
    // root, fused: wait for every rank's flag (a lane per rank), then out[row] = sum over ranks of their partial rows
    if (tid == 0) sflag[0] = ok;
    __syncthreads();
    for (int r = tid; r < ps.nranks; r += T)
        if (r != ps.rank && !spin_until(ps.flags + r, epoch, ps.err, ps.timeout_ns)) sflag[0] = 0;
    __syncthreads();
    const int all = sflag[0];
    const u64* slots = ps.slots + slot0 * ps.slot_stride;
    for (int u = tid; u < (int)nrows * S::SLOTS; u += T) {
        const size_t off = (size_t)(u / S::SLOTS) * S::ELEM_U64 + (size_t)(u % S::SLOTS) * S::SLOT_U64;
        if (all) {
            Val s = S::zero();
            for (int r = 0; r < ps.nranks; r++) S::acc(s, S::load_cv(slots + (size_t)r * ps.slot_stride + off));
            S::store(t.out + off, s);
        } else {
            S::store_poison(t.out + off);  // not a canonical residue: a timed-out commitment cannot pass for a result
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && all) st_release_sys(ps.consumed, epoch);
}

// ---- Goldilocks -------------------------------------------------------------------------------------------------------
// Lazy accumulation: sums of 64 x 64 -> 128-bit products are kept UNREDUCED in 160-bit carry-save accumulators
// (gl::Acc: every partial product is one IMAD.WIDE.U32 with carry, no modular reduction in the column loop).
// The kernel is bound by the wide multiply-add pipe (ncu, profiles/r01b_gl_ncu.md: fmaheavy 77% busy with nine
// products per slot and row), so the slot product in F_p[u]/(u^3 - r) is taken the Karatsuba way, with SIX
// products instead of nine:
//   P0 = a0 x0, P1 = a1 x1, P2 = a2 x2, P01 = (a0+a1)(x0+x1), P02 = (a0+a2)(x0+x2), P12 = (a1+a2)(x1+x2)
//   c0 = P0 + r (P12 - P1 - P2),  c1 = (P01 - P0 - P1) + r P2,  c2 = (P02 - P0 - P2) + P1
// and because the six sums over the columns are linear, the subtractions and the two multiplications by r = 2^40
// happen ONCE per thread after the column loop.  Per slot, row and column: 24 wide multiply-adds + three modular
// additions of the row's limbs (the vector's three sums are shared by all rows) instead of 36.
//
// Data movement: the CTA streams chunks of CS consecutive slots (CS * 24 bytes) of v and of RB matrix rows through an
// NS-deep ring of shared-memory stages filled by 1-D bulk copies (cp.async.bulk + mbarrier complete_tx; SASS
// UBLKCP) issued by a producer warp; RB consumer groups of CS threads (one row each) recycle the stages through
// per-stage "empty" mbarriers, so there is no CTA barrier in the column loop.  Per-thread 24-byte slot reads from a
// stage are conflict-free (stride 6 words, LDS.64).
// (Thread counts are kept at multiples of 128: ptxas sizes the register budget for the thread count rounded up to
// 128, so a dedicated producer warp on top of 4 x 128 consumers cost 24 registers per thread and spilled.  The
// copies are issued by thread 0 of the CTA instead, one stage behind the one being consumed.)
// A thread owns SPT slots of a chunk, CS apart: CS is a multiple of 8, so they have the same slot index and feed the
// same six accumulators; the per-chunk costs (barrier wait, stage release, loop control) are paid once per SPT slots.
// The first version of this kernel (one slot per trip, 64-bit chunk arithmetic) issued 250 instructions per slot and
// row of which 82 were arithmetic, and ran at the issue limit.
template <int RB, int NS, int CS, int SPT>
__global__ void __launch_bounds__(CS * RB, (CS * RB <= 128 ? 4 : (CS * RB <= 256 ? 2 : 1)))
gl_matvec_k6_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                    const u64* __restrict__ v, MvTail tail) {
    typedef GLSlot S;
    constexpr int CHUNK = CS * SPT;                 // slots per chunk
    constexpr uint32_t ROWB = CHUNK * 24;           // bytes per row per stage
    constexpr int ROW_U64 = CHUNK * 3, STAGE_U64 = (RB + 1) * ROW_U64;
    extern __shared__ __align__(128) unsigned char smem[];
    u64* stage = reinterpret_cast<u64*>(smem);      // [NS][RB + 1][CHUNK * 3]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * (RB + 1) * ROWB);
    uint64_t* empty = full + NS;
    int* sflag = reinterpret_cast<int*>(empty + NS);
    const u64** srow = reinterpret_cast<const u64**>(sflag + 4);  // the RB row pointers (read by the issuing thread)
    S::Val* red = reinterpret_cast<S::Val*>(smem);  // the stages are dead once the column loop is over

    const int slot = threadIdx.x % CS, grp = threadIdx.x / CS;  // row row0 + grp
    const size_t total = ncols * S::SLOTS;
    const unsigned nchunks = (unsigned)((total + CHUNK - 1) / CHUNK);
    const int last_valid = (int)(total - (size_t)(nchunks - 1) * CHUNK);  // slots in the globally last chunk
    const int my_chunks = (nchunks > blockIdx.x) ? (int)((nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    auto issue = [&](int it, int s) {  // thread 0 only: chunk `it` of this CTA into stage s
        const unsigned chunk = blockIdx.x + (unsigned)it * gridDim.x;
        const size_t slot0 = (size_t)chunk * CHUNK;
        const uint32_t bytes = (uint32_t)((chunk == nchunks - 1 ? last_valid : CHUNK) * 24);
        mbar_arrive_expect_tx(&full[s], bytes * (RB + 1));
        u64* dst = stage + (size_t)s * STAGE_U64;
        tma_load_1d(dst, v + slot0 * 3, bytes, &full[s]);
#pragma unroll
        for (int r = 0; r < RB; r++) tma_load_1d(dst + (size_t)(r + 1) * ROW_U64, srow[r] + slot0 * 3, bytes, &full[s]);
    };
    if (threadIdx.x == 0) {
        // (the row pointers are parked in shared memory: re-reading the table from global memory at every refill put
        // four dependent L2 round trips into the one warp every stage waits for)
        for (int r = 0; r < RB; r++) srow[r] = (row0 + r < nrows) ? rows[row0 + r] : rows[nrows - 1];
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], RB * CS / 32); }
        mbar_fence_init();
        for (int it = 0; it < NS && it < my_chunks; it++) issue(it, it);
    }
    __syncthreads();

    gl::Acc P0, P1, P2, P01, P02, P12;
    gl::acc_zero(P0); gl::acc_zero(P1); gl::acc_zero(P2); gl::acc_zero(P01); gl::acc_zero(P02); gl::acc_zero(P12);

    const u64* tbase = stage + slot * 3;  // this thread's first slot of the vector in stage 0
    // one chunk: stage s holds chunk `it` of this CTA.  RAGGED: the globally last chunk may be short (stale bytes of
    // an earlier chunk sit behind it in the stage): only then are the loads guarded.
    auto step = [&](int it, int s, uint32_t phase, auto ragged) {
        if (threadIdx.x == 0 && it > 0 && it - 1 + NS < my_chunks) {
            // refill the stage that was consumed one trip ago (no CTA barrier: the other warps never wait for this
            // one, and this one only waits for warps that lag a whole trip behind).  (Rotating the refill over the
            // warps, to spread its ~90 instructions, was measured 50% SLOWER: a refill issued by whichever warp
            // happens to lag arrives late for everybody.)
            const int ps = (s + NS - 1) % NS;
            mbar_wait(&empty[ps], s == 0 ? (phase ^ 1) : phase);
            issue(it - 1 + NS, ps);
        }
        __syncwarp();
        mbar_wait(&full[s], phase);
        const u64* bx = tbase + s * STAGE_U64;
        const u64* ba = bx + (grp + 1) * ROW_U64;
        u64 x[SPT][3], a[SPT][3];
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            const bool live = !decltype(ragged)::value || slot + q * CS < last_valid;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                x[q][k] = live ? bx[q * CS * 3 + k] : 0;
                a[q][k] = live ? ba[q * CS * 3 + k] : 0;
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            // limbs in memory are canonical, so the weak-form additions are exact residues (a + b, b canonical)
            const u64 x01 = gl::add_cc(x[q][0], x[q][1]), x02 = gl::add_cc(x[q][0], x[q][2]), x12 = gl::add_cc(x[q][1], x[q][2]);
            const u64 a01 = gl::add_cc(a[q][0], a[q][1]), a02 = gl::add_cc(a[q][0], a[q][2]), a12 = gl::add_cc(a[q][1], a[q][2]);
            gl::acc_mad(P0, a[q][0], x[q][0]);
            gl::acc_mad(P1, a[q][1], x[q][1]);
            gl::acc_mad(P2, a[q][2], x[q][2]);
            gl::acc_mad(P01, a01, x01);
            gl::acc_mad(P02, a02, x02);
            gl::acc_mad(P12, a12, x12);
        }
    };
    // this CTA's last chunk is ragged iff it is the globally last chunk and that one is short
    const bool own_ragged = my_chunks > 0 && last_valid != CHUNK &&
                            blockIdx.x + (unsigned)(my_chunks - 1) * gridDim.x == nchunks - 1;
    const int full_chunks = my_chunks - (own_ragged ? 1 : 0);
    uint32_t phase = 0;
    int it = 0;
    for (; it + NS <= full_chunks; it += NS, phase ^= 1) {
#pragma unroll
        for (int s = 0; s < NS; s++) step(it + s, s, phase, std::false_type());
    }
    for (int s = 0; it < full_chunks; it++, s++) step(it, s, phase, std::false_type());
    if (own_ragged) step(it, it % NS, phase, std::true_type());
    // the column loop is over: the next kernel on the stream may start its own (its tail will wait for ours)
    pdl_wait();
    pdl_launch_dependents();
    __syncthreads();  // every stage has been consumed: `red` may overwrite them
    {
        const u64 p0 = gl::acc_reduce(P0), p1 = gl::acc_reduce(P1), p2 = gl::acc_reduce(P2);
        const u64 p01 = gl::acc_reduce(P01), p02 = gl::acc_reduce(P02), p12 = gl::acc_reduce(P12);
        constexpr int E = gl::root_exp(1);  // u^3 = r = 2^40
        const u64 c0 = gl::add(gl::mul_pow2<E>(gl::sub(gl::sub(p12, p1), p2)), p0);
        const u64 c1 = gl::add(gl::sub(gl::sub(p01, p0), p1), gl::canon(gl::mul_pow2<E>(p2)));
        const u64 c2 = gl::add(gl::sub(gl::sub(p02, p0), p2), p1);
        S::Val mine;  // Montgomery layout: the product of two raw limbs carries an extra 2^-64 = 2^128
        mine.c[0] = gl::canon(gl::mul_pow2<128>(c0));
        mine.c[1] = gl::canon(gl::mul_pow2<128>(c1));
        mine.c[2] = gl::canon(gl::mul_pow2<128>(c2));
        red[threadIdx.x] = mine;
    }
    __syncthreads();
    if (threadIdx.x < RB * S::SLOTS) {  // thread (r, s8): CTA sum of slot index s8 of row r, fixed order
        const int r = threadIdx.x / S::SLOTS, s8 = threadIdx.x % S::SLOTS;
        S::Val sacc = red[r * CS + s8];
        for (int k = s8 + S::SLOTS; k < CS; k += S::SLOTS) S::acc(sacc, red[r * CS + k]);
        if (row0 + r < nrows)
            S::store(tail.partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + s8 * S::SLOT_U64, sacc);
    }
    const int rb = (nrows - row0 < (size_t)RB) ? (int)(nrows - row0) : RB;
    __syncthreads();
    mv_tail<S>(tail, nrows, row0, rb, red, sflag);
}

// <<<>>> with the programmatic-dependent-launch attribute when `pdl` is set
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t st,
                              bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int RB, int NS, int CS, int SPT>
static cudaError_t gl_k6_launch(const u64* const* d_rows, size_t nrows, size_t row0, size_t ncols, const u64* v,
                                const MvTail& tail, int max_grid, cudaStream_t st, int sms, bool pdl) {
    auto kern = gl_matvec_k6_kernel<RB, NS, CS, SPT>;
    const int threads = CS * RB;
    const size_t smem = (size_t)NS * (RB + 1) * CS * SPT * 24 + 2 * NS * sizeof(uint64_t) + 16 + 4 * sizeof(void*);
    static KernelCache cache;
    int bps = 0;
    cudaError_t e = cache.configure(kern, threads, smem, &bps);
    if (e != cudaSuccess) return e;
    const size_t nchunks = (ncols * GLSlot::SLOTS + CS * SPT - 1) / (CS * SPT);
    if (nchunks >= (1ull << 31)) return cudaErrorInvalidValue;
    size_t grid = (size_t)sms * bps;
    if (grid > (size_t)max_grid) grid = max_grid;
    if (pdl && grid > 1) grid -= 1;  // the slot the previous kernel's tail may still hold
    if (grid > nchunks) grid = nchunks;
    return launch_pdl(kern, (unsigned)grid, (unsigned)threads, smem, st, pdl, d_rows, nrows, row0, ncols, v, tail);
}

#ifndef SR_MV_T
#define SR_MV_T 256
#endif
#ifndef SR_MV_RB
#define SR_MV_RB 4
#endif
#ifndef SR_MV_BB_PREFETCH
#define SR_MV_BB_PREFETCH 1  // BabyBear: next trip's loads before this trip's products
#endif
#ifndef SR_MV_BB_GROUPS
#define SR_MV_BB_GROUPS 2  // BabyBear: row groups per CTA (matvec_partial_kernel)
#endif
constexpr int MV_T = SR_MV_T;  // threads per CTA (multiple of every SLOTS)

// BabyBear / Starknet prime: thread t owns slot (t mod SLOTS) of columns t / SLOTS, t / SLOTS + stride, ...;
// consecutive threads read consecutive slots, i.e. each warp reads one contiguous span of a row.  Every thread keeps
// one accumulator per matrix row (RB rows per pass): S::Accum, which for BabyBear is the slot product's nine 64-bit
// sums kept unreduced over all columns (one reduction pair per coefficient at the end), with the vector operand
// prepared once per slot and shared by the rows (S::Prep).
// G thread groups per CTA split the RB rows of the pass between them (BabyBear: two groups of two rows, 36 accumulator
// registers per thread instead of 72, which keeps two CTAs per SM resident); the groups walk the same slots, so the
// second group's reads of v hit L1.
template <class S, int RB, int G, bool PF = false>
__global__ void __launch_bounds__(MV_T)
matvec_partial_kernel(const u64* const* __restrict__ rows, size_t nrows, size_t row0, size_t ncols,
                      const u64* __restrict__ v, MvTail tail) {
    static_assert(RB % G == 0 && MV_T % G == 0 && (MV_T / G) % S::SLOTS == 0, "row groups");
    constexpr int RPT = RB / G, TG = MV_T / G;  // rows per thread, threads per group
    __shared__ typename S::Val red[MV_T];
    __shared__ int sflag[2];
    const int grp = threadIdx.x / TG, tl = threadIdx.x % TG;
    typename S::Accum acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) S::accum_zero(acc[r]);
    const u64* rp[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) rp[r] = (row0 + grp * RPT + r < nrows) ? rows[row0 + grp * RPT + r] : nullptr;

    const size_t total = ncols * S::SLOTS;  // slots per row
    const size_t stride = (size_t)gridDim.x * TG;
    if (PF) {
        // software pipeline: the loads of the next trip are issued before this trip's products (ncu of the unreduced
        // BabyBear kernel: long_scoreboard 4.0 warps per issue, the multiply-add pipe 62 % busy)
        size_t g = (size_t)blockIdx.x * TG + tl;
        typename S::Val xn = S::zero(), an[RPT];
#pragma unroll
        for (int r = 0; r < RPT; r++) an[r] = S::zero();
        if (g < total) {
            xn = S::load_cached(v + g * S::SLOT_U64);
#pragma unroll
            for (int r = 0; r < RPT; r++)
                if (rp[r]) an[r] = S::load(rp[r] + g * S::SLOT_U64);
        }
        while (g < total) {
            const typename S::Val xc = xn;
            typename S::Val a[RPT];
#pragma unroll
            for (int r = 0; r < RPT; r++) a[r] = an[r];
            const size_t g2 = g + stride;
            if (g2 < total) {
                xn = S::load_cached(v + g2 * S::SLOT_U64);
#pragma unroll
                for (int r = 0; r < RPT; r++)
                    if (rp[r]) an[r] = S::load(rp[r] + g2 * S::SLOT_U64);
            }
            const typename S::Prep x = S::prep(xc);
#pragma unroll
            for (int r = 0; r < RPT; r++)
                if (rp[r]) S::accum_mad_p(acc[r], a[r], x);
            g = g2;
        }
    } else {
        for (size_t g = (size_t)blockIdx.x * TG + tl; g < total; g += stride) {
            const typename S::Prep x = S::prep(S::load_cached(v + g * S::SLOT_U64));
            typename S::Val a[RPT];
#pragma unroll
            for (int r = 0; r < RPT; r++)
                if (rp[r]) a[r] = S::load(rp[r] + g * S::SLOT_U64);
#pragma unroll
            for (int r = 0; r < RPT; r++)
                if (rp[r]) S::accum_mad_p(acc[r], a[r], x);
        }
    }
    pdl_wait();  // the previous kernel on the stream (its tail reads the scratch this one is about to write) is done
    pdl_launch_dependents();
    // CTA reduction per slot index: thread tl of the row's group holds slot tl % SLOTS
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (row0 + r >= nrows) break;
        if (grp == r / RPT) red[tl] = S::accum_result(acc[r % RPT]);
        __syncthreads();
        if (threadIdx.x < S::SLOTS) {
            typename S::Val s = red[threadIdx.x];
            for (int k = threadIdx.x + S::SLOTS; k < TG; k += S::SLOTS) S::acc(s, red[k]);
            S::store(tail.partial + ((size_t)blockIdx.x * nrows + row0 + r) * S::ELEM_U64 + threadIdx.x * S::SLOT_U64, s);
        }
        __syncthreads();
    }
    const int rb = (nrows - row0 < (size_t)RB) ? (int)(nrows - row0) : RB;
    mv_tail<S>(tail, nrows, row0, rb, red, sflag);
}

// ---- stand-alone modular sum of partial rows ---------------------------------------------------------------------------
// out[row] = sum_k parts[(k * stride_rows + row)]  (k < nparts).  One CTA per (row, slot), fixed-order tree:
// deterministic.  This is sr_modsum_partials (the rank-0 sum behind an NCCL all-gather) and, with ps.role 2, the
// root's separate mailbox reduction (sr_commit_reduce): `parts` is replaced by the epoch's nranks slots; the kernel
// first acquires all rank flags and publishes `consumed` at the end.
template <class S>
__global__ void __launch_bounds__(128)
sum_partials_kernel(const u64* __restrict__ parts, size_t nparts, size_t stride_rows, size_t nrows,
                    u64* __restrict__ out, PeerSync ps) {
    __shared__ typename S::Val red[128];
    __shared__ int sok;
    u64 epoch = 0;
    int ok = 1;
    if (ps.role == 2) {  // uniform per launch
        epoch = ps.epoch ? ps.epoch : *reinterpret_cast<volatile u64*>(ps.epoch_ctr) + 1;
        parts = ps.slots + (size_t)(epoch % MAILBOX_DEPTH) * ps.nranks * ps.slot_stride;
        if (threadIdx.x == 0) sok = 1;
        __syncthreads();
        for (int r = threadIdx.x; r < ps.nranks; r += 128)  // a lane per rank
            if (!spin_until(ps.flags + r, epoch, ps.err, ps.timeout_ns)) sok = 0;
        __syncthreads();
        ok = sok;
    }
    const size_t idx = blockIdx.x;  // (row, slot)
    const size_t row = idx / S::SLOTS, slot = idx % S::SLOTS;
    typename S::Val s = S::zero();
    for (size_t k = threadIdx.x; k < nparts; k += 128)
        S::acc(s, S::load_cv(parts + (k * stride_rows + row) * S::ELEM_U64 + slot * S::SLOT_U64));
    red[threadIdx.x] = s;
    __syncthreads();
#pragma unroll
    for (int w = 64; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) {
            typename S::Val t = red[threadIdx.x];
            S::acc(t, red[threadIdx.x + w]);
            red[threadIdx.x] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (ok) S::store(out + row * S::ELEM_U64 + slot * S::SLOT_U64, red[0]);
        else S::store_poison(out + row * S::ELEM_U64 + slot * S::SLOT_U64);
    }
    if (ps.role == 2) {
        // every block has read the epoch counter before the last one arrives, so it may be advanced here
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned arrived = atomicAdd(ps.counter, 1u);
            if (arrived == gridDim.x - 1) {
                *ps.counter = 0;  // ready for the next launch on this stream
                if (ok) st_release_sys(ps.consumed, epoch);
            }
        }
    }
}

static int mv_grid(int sms) { return sms * 4; }

size_t matvec_scratch_bytes(int ring, size_t nrows, int sms) {
    const size_t w = ring == RING_GL ? 24 : ring == RING_BB ? 72 : 64;
    return (size_t)mv_grid(sms) * nrows * w * 8 + 16;
}

// Sum of nothing = ZERO (ncols == 0): one CTA runs the tail over zero partials, so that a mailbox flag is still
// published and the root still sums the other ranks.
template <class S>
__global__ void __launch_bounds__(MV_T)
matvec_empty_kernel(size_t nrows, size_t row0, MvTail tail) {
    __shared__ typename S::Val red[MV_T];
    __shared__ int sflag[2];
    const int rb = (nrows - row0 < (size_t)SR_MV_RB) ? (int)(nrows - row0) : SR_MV_RB;
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x < rb * S::SLOTS)
        S::store(tail.partial + (row0 + threadIdx.x / S::SLOTS) * S::ELEM_U64 + (threadIdx.x % S::SLOTS) * S::SLOT_U64,
                 S::zero());
    mv_tail<S>(tail, nrows, row0, rb, red, sflag);
}

// CTA shapes of the Goldilocks kernel per number of rows in the pass (B200, m = 2^20, graph-replayed commitments,
// profiles/r02_matvec_tuning.md): <rows, stages, slots per row group, slots per thread>.  Four rows: 64 threads per
// row and four slots per thread, two 256-thread CTAs per SM: 0.767 of the HBM roofline against 0.670 for one
// 512-thread CTA with two slots per thread (more independent work per thread, half the per-chunk overhead per slot,
// and two CTAs that are never in the same phase).
#define SR_GLK_SHAPE_1 1, 3, 64, 4
#define SR_GLK_SHAPE_2 2, 3, 128, 3
#define SR_GLK_SHAPE_3 3, 3, 64, 4
#ifndef SR_GLK4_NS  // (overridable for tuning builds: make EXTRA="-DSR_GLK4_NS=4 -DSR_GLK4_CS=128 -DSR_GLK4_SPT=2")
#define SR_GLK4_NS 3
#define SR_GLK4_CS 64
#define SR_GLK4_SPT 4
#endif
#define SR_GLK_SHAPE_4 4, SR_GLK4_NS, SR_GLK4_CS, SR_GLK4_SPT
// Short products (a rank's shard of a sharded commitment): one 512-thread CTA per SM, three slots per thread.  With
// two CTAs per SM the hand-off tail of a commitment was measured NOT to hide behind the next commitment's column loop
// (8.3 us exposed per commitment at 2^17 columns per rank against 3.2 us for this shape, 38.2 vs 34.8 us per step).
#define SR_GLK_SHAPE_4_SHORT 4, 3, 128, 3
constexpr size_t GLK_SHORT_COLS = (size_t)1 << 19;
template <class S>
static cudaError_t matvec_launch_t(int ring, const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v,
                                   u64* out, void* scratch, unsigned* counters, unsigned* seq, bool pdl,
                                   cudaStream_t st, int sms, int* launches, const PeerSync& ps) {
    *launches = 0;
    if (nrows == 0) return cudaSuccess;
    MvTail tail = {};
    tail.partial = reinterpret_cast<u64*>(scratch);
    tail.out = out;
    tail.counter = counters;
    tail.ps = ps;
    constexpr int RB = SR_MV_RB;
    for (size_t row0 = 0; row0 < nrows; row0 += RB) {
        tail.last_pass = (row0 + RB >= nrows) ? 1 : 0;
        (*seq)++;
        cudaError_t e = cudaSuccess;
        if (ncols == 0) {
            e = launch_pdl(matvec_empty_kernel<S>, 1u, (unsigned)MV_T, 0, st, pdl, nrows, row0, tail);
        } else if (ring == RING_GL) {
            const size_t left = nrows - row0;
            const int mg = mv_grid(sms);
            if (left >= 4 && ncols < GLK_SHORT_COLS)
                e = gl_k6_launch<SR_GLK_SHAPE_4_SHORT>(d_rows, nrows, row0, ncols, v, tail, mg, st, sms, pdl);
            else if (left >= 4) e = gl_k6_launch<SR_GLK_SHAPE_4>(d_rows, nrows, row0, ncols, v, tail, mg, st, sms, pdl);
            else if (left == 3) e = gl_k6_launch<SR_GLK_SHAPE_3>(d_rows, nrows, row0, ncols, v, tail, mg, st, sms, pdl);
            else if (left == 2) e = gl_k6_launch<SR_GLK_SHAPE_2>(d_rows, nrows, row0, ncols, v, tail, mg, st, sms, pdl);
            else e = gl_k6_launch<SR_GLK_SHAPE_1>(d_rows, nrows, row0, ncols, v, tail, mg, st, sms, pdl);
        } else if constexpr (!std::is_same<S, GLSlot>::value) {
            constexpr int G = std::is_same<S, BBSlot>::value ? SR_MV_BB_GROUPS : 1;
            constexpr bool PF = std::is_same<S, BBSlot>::value && SR_MV_BB_PREFETCH;  // (Starknet: 32 more registers than fit)
            const size_t total = ncols * S::SLOTS;
            int grid = mv_grid(sms);
            const size_t need = (total + MV_T / G - 1) / (MV_T / G);
            if ((size_t)grid > need) grid = (int)need;
            if (G > 1 && nrows - row0 <= (size_t)(RB / G)) {
                // the rows left fit one group: every thread takes all of them (no idle group)
                const size_t need1 = (total + MV_T - 1) / MV_T;
                int grid1 = mv_grid(sms);
                if ((size_t)grid1 > need1) grid1 = (int)need1;
                e = launch_pdl(matvec_partial_kernel<S, RB / G, 1, PF>, (unsigned)grid1, (unsigned)MV_T, 0, st, pdl, d_rows,
                               nrows, row0, ncols, v, tail);
            } else {
                e = launch_pdl(matvec_partial_kernel<S, RB, G, PF>, (unsigned)grid, (unsigned)MV_T, 0, st, pdl, d_rows,
                               nrows, row0, ncols, v, tail);
            }
        }
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    return cudaSuccess;
}

// ps (optional): the result goes to / through the root's mailbox; see PeerSync
cudaError_t matvec_launch(int ring, const u64* const* d_rows, size_t nrows, size_t ncols, const u64* v, u64* out,
                          void* scratch, unsigned* counters, unsigned* seq, bool pdl, cudaStream_t st, int sms,
                          int* launches, const PeerSync* ps) {
    const PeerSync none = {};
    const PeerSync& p = ps ? *ps : none;
    switch (ring) {
    case RING_GL: return matvec_launch_t<GLSlot>(ring, d_rows, nrows, ncols, v, out, scratch, counters, seq, pdl, st, sms, launches, p);
    case RING_BB: return matvec_launch_t<BBSlot>(ring, d_rows, nrows, ncols, v, out, scratch, counters, seq, pdl, st, sms, launches, p);
    case RING_SP: return matvec_launch_t<SPSlot>(ring, d_rows, nrows, ncols, v, out, scratch, counters, seq, pdl, st, sms, launches, p);
    }
    return cudaErrorInvalidValue;
}

// out[i] = sum_r gathered[(r * stride_rows + i)] over nranks partials; ps (optional): the root's mailbox reduction
template <class S>
static cudaError_t modsum_t(const u64* g, size_t nranks, size_t stride_rows, size_t nrows, u64* out,
                            const PeerSync& ps, cudaStream_t st) {
    const size_t n = nrows * S::SLOTS;
    sum_partials_kernel<S><<<(unsigned)n, 128, 0, st>>>(g, nranks, stride_rows, nrows, out, ps);
    return cudaGetLastError();
}
cudaError_t modsum_launch(int ring, const u64* gathered, size_t nranks, size_t stride_rows, size_t nrows, u64* out,
                          cudaStream_t st, const PeerSync* ps) {
    const PeerSync none = {};
    const PeerSync& p = ps ? *ps : none;
    switch (ring) {
    case RING_GL: return modsum_t<GLSlot>(gathered, nranks, stride_rows, nrows, out, p, st);
    case RING_BB: return modsum_t<BBSlot>(gathered, nranks, stride_rows, nrows, out, p, st);
    case RING_SP: return modsum_t<SPSlot>(gathered, nranks, stride_rows, nrows, out, p, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sr
