"""world_size-2 gloo test (CPU) of the sharded-commit host logic: column partition, gather layout
[rank][row][limb], rank-0 modular sum.  The per-rank partial products and the modular sum are
injected from the oracle here (the product has no CPU path); the GPU run uses the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle as C
from tests.util import WORDS, rand_raw

NAME, KAPPA, M = "goldilocks", 3, 37


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Shard:
    def __init__(self, rows, nrows):
        self.rows, self.nrows, self.config, self.ctx = rows, nrows, None, None


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stark_rings_b200.dist import shard_columns, sharded_commit
    w = WORDS[NAME]
    rows = [rand_raw(NAME, M, 40 + i) for i in range(KAPPA)]
    v = rand_raw(NAME, M, 50)
    lo, hi = shard_columns(M, world, rank)
    shard = _Shard([r[lo * w:hi * w].copy() for r in rows], KAPPA)
    vs = v[lo * w:hi * w].copy()

    def partial_fn(A, vv):
        return torch.from_numpy(C.matvec(NAME, A.rows, vv).view(np.int64))

    def modsum_fn(gathered, nranks, nrows):
        g = gathered.numpy().view(np.uint64)
        # an all-ones "vector" turns the oracle mat-vec into a modular sum over ranks: use plain field adds
        out = np.zeros(nrows * w, dtype=np.uint64)
        p = 18446744069414584321
        for r in range(nranks):
            blk = g[r * nrows * w:(r + 1) * nrows * w]
            out = np.array([(int(x) + int(y)) % p for x, y in zip(out, blk)], dtype=np.uint64)
        return out

    y = sharded_commit(shard, vs, world, rank, partial_fn=partial_fn, modsum_fn=modsum_fn)
    if rank == 0:
        want = C.matvec(NAME, rows, v)
        q.put(bool(np.array_equal(y, want)))
    else:
        assert y is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_columns_partition():
    from stark_rings_b200.dist import shard_columns
    for m in (0, 1, 7, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_columns(m, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_sharded_commit_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
