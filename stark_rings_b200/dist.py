"""Column-sharded ring matrix x vector product (Ajtai-style commitment) over several GPUs.

One process per GPU (torchrun).  The m columns are split evenly; rank r holds A[:, cols_r] and v[cols_r],
computes nrows partial ring elements with sr_matvec_partial, the partials are all-gathered as raw u64
limbs (an NCCL sum would wrap mod 2^64, not mod p) and rank 0 adds them mod p with sr_modsum_partials.
The reference has no counterpart (it is single-process); the single-GPU semantics are those of
Matrix::checked_mul_vec (linear_algebra/src/matrix.rs:168-178).
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
