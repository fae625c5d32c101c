"""Column-sharded ring matrix x vector product (Ajtai-style commitment) over several GPUs.

One process per GPU (torchrun).  The m columns are split evenly; rank r holds A[:, cols_r] and v[cols_r],
computes nrows partial ring elements with sr_matvec_partial, the partials are all-gathered as raw u64
limbs (an NCCL sum would wrap mod 2^64, not mod p) and rank 0 adds them mod p with sr_modsum_partials.
The reference has no counterpart (it is single-process); the single-GPU semantics are those of
Matrix::checked_mul_vec (linear_algebra/src/matrix.rs:168-178).

Two exchange paths: `sharded_commit` (NCCL all-gather of the partials, then sr_modsum_partials on rank 0) and
`PeerCommit` (the partials are stored straight into the root's HBM over NVLink by the tail of the kernel that
produces them; the tail of the root's own kernel acquires the per-rank flags and sums: one launch per rank and no
collective call on the data path).
"""
from __future__ import annotations

import ctypes


def shard_columns(ncols: int, world: int, rank: int):
    """[lo, hi) of the columns rank owns: contiguous, sizes differ by at most one."""
    return ncols * rank // world, ncols * (rank + 1) // world


def gather_partials(partial, world: int, group=None):
    """all_gather of one rank's nrows*limbs partial limbs -> tensor laid out [rank][row][limb]."""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    if world == 1:
        out.copy_(partial)
    else:
        dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out


def sharded_commit(matrix_shard, v_shard, world: int, rank: int, group=None, partial_fn=None, modsum_fn=None):
    """y = A v with A, v column-sharded.  Returns the nrows result elements (flat limbs) on rank 0 and
    None elsewhere.  partial_fn / modsum_fn default to the CUDA library; tests inject CPU checkers."""
    if partial_fn is None:
        partial_fn = lambda A, v: A.partial_mul_vec(v).data
    part = partial_fn(matrix_shard, v_shard)
    gathered = gather_partials(part, world, group)
    if rank != 0:
        return None
    nrows = matrix_shard.nrows
    if modsum_fn is None:
        import torch
        from . import _lib as L
        from .rings import default_context
        cfg = matrix_shard.config
        c = matrix_shard.ctx or default_context(gathered.device.index)
        c.use_torch_stream()
        out = torch.empty(nrows * cfg.limbs, dtype=gathered.dtype, device=gathered.device)
        c.check(L.lib.sr_modsum_partials(c.h, cfg.ring_id, ctypes.c_void_p(gathered.data_ptr()), world, nrows,
                                         ctypes.c_void_p(out.data_ptr()), L.SR_DEVICE), "sr_modsum_partials")
        return out
    return modsum_fn(gathered, world, nrows)


class CommitTimeout(RuntimeError):
    """A wait inside a commitment kernel ran out of its budget (a lost or very late peer): the result is invalid."""


class PeerCommit:
    """Column-sharded commitment over NVLink peer memory (sr_mailbox_* / sr_commit_* of the C ABI).

    The root rank owns a mailbox in its HBM; every rank maps it through CUDA IPC (the 64-byte handle travels once,
    at construction, through `exchange`, by default torch.distributed.broadcast_object_list, followed by a barrier so
    that no rank starts before all have opened the mailbox).  Per commitment every rank launches ONE kernel per
    four matrix rows: its tail stores the rank's partial elements straight into the root's mailbox and publishes an
    epoch flag; on the root the same tail then acquires the flags of all ranks and adds the partials mod p
    (`fused=True`, sr_commit_root).  No NCCL call, no host synchronisation and no separate reduction kernel on the
    data path.  With `fused=False` the root sends like everybody else and runs sr_commit_reduce as a second kernel.

    Calls are asynchronous.  A rank that is later than the mailbox's wait budget (`timeout_s`, default 4 s) makes
    the kernels give up instead of hanging the GPU: the result is then all-ones limbs and `check()` /
    `synchronize()` raise CommitTimeout.  Call one of them before trusting a result.

    `ranks_here` emulation (tests): several ranks inside ONE process on one GPU send one after the other with
    `send(..., as_rank=r)` into the process's own mailbox, the root's share last (`root_commit`) or followed by
    `reduce`.
    """

    def __init__(self, config, nrows_max, world, rank, ctx, root=0, exchange=None, device_epochs=False, fused=True,
                 timeout_s=None):
        """device_epochs: the kernels count the commitments themselves (epoch argument 0), so that every step issues
        identical launches and can be captured in a CUDA graph."""
        import ctypes as C
        from . import _lib as L
        self.config, self.world, self.rank, self.root, self.ctx = config, world, rank, root, ctx
        self.nrows_max, self.epoch, self.device_epochs, self.fused = nrows_max, 0, device_epochs, fused
        self.box = C.c_void_p()
        handle = None
        if rank == root:
            buf = C.create_string_buffer(L.SR_IPC_HANDLE_BYTES)
            ctx.check(L.lib.sr_mailbox_create(ctx.h, config.ring_id, nrows_max, world, C.byref(self.box), buf),
                      "sr_mailbox_create")
            handle = buf.raw
        barrier = None
        if exchange is None and world > 1:
            import torch.distributed as dist

            def exchange(h):
                obj = [h]
                dist.broadcast_object_list(obj, src=root)
                return obj[0]
            barrier = dist.barrier
        if world > 1 and exchange is not None:
            handle = exchange(handle)
            if rank != root:
                ctx.check(L.lib.sr_mailbox_open(ctx.h, config.ring_id, nrows_max, world, handle, C.byref(self.box)),
                          "sr_mailbox_open")
        if timeout_s is not None:
            ctx.check(L.lib.sr_mailbox_set_timeout(ctx.h, self.box, int(timeout_s * 1e9)), "sr_mailbox_set_timeout")
        if barrier is not None:
            barrier()  # nobody commits before every rank has the mailbox mapped

    def _product(self, matrix_shard, v_shard, as_rank, out, ctx=None):
        import ctypes as C
        from . import _lib as L
        from .rings import _ptr_loc
        cfg = self.config
        ctx = ctx or self.ctx
        pv, nv, loc, dev = _ptr_loc(v_shard.data)
        if loc != L.SR_DEVICE:
            raise ValueError("PeerCommit works on device-resident shards")
        ptrs = (C.c_void_p * max(matrix_shard.nrows, 1))()
        for i, r in enumerate(matrix_shard.vals):
            ptrs[i] = _ptr_loc(r.data)[0]
        ctx.use_torch_stream()
        rank = self.rank if as_rank is None else as_rank
        epoch = 0 if self.device_epochs else self.epoch
        if out is None:
            rc = L.lib.sr_commit_send(ctx.h, cfg.ring_id, ptrs, matrix_shard.nrows, matrix_shard.ncols, pv, nv,
                                      self.box, rank, epoch)
            ctx.check(rc, "sr_commit_send")
        else:
            rc = L.lib.sr_commit_root(ctx.h, cfg.ring_id, ptrs, matrix_shard.nrows, matrix_shard.ncols, pv, nv,
                                      self.box, rank, epoch, C.c_void_p(out.data_ptr()))
            ctx.check(rc, "sr_commit_root")

    def send(self, matrix_shard, v_shard, as_rank=None, ctx=None):
        """This rank's share of the product, written into the root's mailbox (asynchronous).  `ctx`: the context
        whose scratch / row table the product uses (default: the one given at construction); a caller that commits
        several resident matrices in turn keeps one context per matrix, so that no row table is re-uploaded."""
        self._product(matrix_shard, v_shard, as_rank, None, ctx)

    def root_commit(self, matrix_shard, v_shard, out, as_rank=None, ctx=None):
        """Root only: the root's share AND the modular sum over all ranks in one kernel (asynchronous)."""
        self._product(matrix_shard, v_shard, as_rank, out, ctx)
        return out

    def reduce(self, nrows, out):
        """Root only, two-kernel mode: out <- sum of the partials of the current epoch mod p (asynchronous)."""
        import ctypes as C
        from . import _lib as L
        self.ctx.use_torch_stream()
        self.ctx.check(L.lib.sr_commit_reduce(self.ctx.h, self.config.ring_id, self.box, nrows,
                                              0 if self.device_epochs else self.epoch,
                                              C.c_void_p(out.data_ptr())), "sr_commit_reduce")
        return out

    def commit(self, matrix_shard, v_shard, out=None, ctx=None):
        """One commitment: every rank calls it with its column shard; returns the result on the root, None elsewhere."""
        import torch
        self.epoch += 1
        if self.rank != self.root:
            self.send(matrix_shard, v_shard, ctx=ctx)
            return None
        if out is None:
            out = torch.empty(matrix_shard.nrows * self.config.limbs, dtype=v_shard.data.dtype,
                              device=v_shard.data.device)
        if self.fused:
            return self.root_commit(matrix_shard, v_shard, out, ctx=ctx)
        self.send(matrix_shard, v_shard, ctx=ctx)
        return self.reduce(matrix_shard.nrows, out)

    def timed_out(self) -> bool:
        """Synchronises the context's stream and reads the mailbox's error flag."""
        import ctypes as C
        from . import _lib as L
        flag = C.c_int(0)
        self.ctx.check(L.lib.sr_mailbox_error(self.ctx.h, self.box, C.byref(flag)), "sr_mailbox_error")
        return bool(flag.value)

    def check(self):
        """Raises CommitTimeout if any wait of any commitment so far gave up (synchronises the stream)."""
        if self.timed_out():
            raise CommitTimeout("a commitment kernel gave up waiting for a peer: results since then are invalid")

    synchronize = check

    def close(self):
        from . import _lib as L
        if self.box:
            L.lib.sr_mailbox_destroy(self.ctx.h, self.box)
            self.box = None
