"""Column-sharded commitment over the peer-memory mailbox (sr_mailbox_* / sr_commit_*), against the oracle's full
mat-vec.  Single process: G ranks emulated on one GPU (same kernels, flags, epochs and slot reuse; the mailbox is
local).  Two processes on two GPUs (CUDA IPC over NVLink) when the box has them."""
import os
import socket

import numpy as np
import pytest

from oracle import c_oracle as C
from tests.util import WORDS, rand_raw

pytestmark = pytest.mark.gpu

ALL = ["goldilocks", "babybear", "stark_prime"]


def dev(a, device="cuda"):
    import torch
    return torch.from_numpy(a.view(np.int64)).to(device)


def host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("kappa,m,G", [(4, 1030, 4), (1, 64, 2), (7, 333, 8)])
def test_peer_commit_emulated_ranks(name, kappa, m, G, fused):
    """fused=False: every rank sends, then the separate reduction kernel; fused=True: ranks 1..G-1 send, then the
    root's product kernel sums everything in its own tail (sr_commit_root)."""
    import torch
    import stark_rings_b200 as S
    from stark_rings_b200.dist import PeerCommit, shard_columns
    cfg, w = S.CONFIGS[name], WORDS[name]
    ctx = S.default_context(0)
    pc = PeerCommit(cfg, kappa, G, 0, ctx, exchange=lambda h: h)
    try:
        for epoch in range(1, 11):  # more than twice the mailbox depth: slots and flags are reused
            rows = [rand_raw(name, m, 40 + i + 100 * epoch) for i in range(kappa)]
            v = rand_raw(name, m, 50 + epoch)
            want = C.matvec(name, rows, v, threads=8)
            pc.epoch += 1
            out = torch.empty(kappa * w, dtype=torch.int64, device="cuda")
            for r in list(range(1, G)) + [0]:  # the root's share last: in one process nothing may wait for later work
                lo, hi = shard_columns(m, G, r)
                A = S.Matrix([S.RqNTT(cfg, dev(x[lo * w:hi * w].copy())) for x in rows])
                vs = S.RqNTT(cfg, dev(v[lo * w:hi * w].copy()))
                if fused and r == 0:
                    pc.root_commit(A, vs, out, as_rank=0)
                else:
                    pc.send(A, vs, as_rank=r)
            if not fused:
                pc.reduce(kappa, out)
            assert np.array_equal(host(out), want), epoch
        pc.check()
    finally:
        pc.close()


@pytest.mark.parametrize("fused", [False, True])
def test_peer_commit_lost_peer_times_out_instead_of_hanging(fused):
    """A rank that never sends: the root gives up after the mailbox's wait budget (0.2 s here), flags the error and
    leaves all-ones limbs (not a canonical residue) instead of a plausible commitment.  ONE kernel waits, alone on
    the GPU."""
    import torch
    import stark_rings_b200 as S
    from stark_rings_b200.dist import CommitTimeout, PeerCommit
    name = "goldilocks"
    cfg, w = S.CONFIGS[name], WORDS[name]
    ctx = S.default_context(0)
    pc = PeerCommit(cfg, 2, 2, 0, ctx, exchange=lambda h: h, timeout_s=0.2)
    try:
        rows = [rand_raw(name, 8, i) for i in range(2)]
        A = S.Matrix([S.RqNTT(cfg, dev(x)) for x in rows])
        vs = S.RqNTT(cfg, dev(rand_raw(name, 8, 9)))
        out = torch.zeros(2 * w, dtype=torch.int64, device="cuda")
        pc.epoch += 1
        if fused:
            pc.root_commit(A, vs, out, as_rank=0)  # rank 1 stays silent
        else:
            pc.send(A, vs, as_rank=0)
            pc.reduce(2, out)
        assert pc.timed_out()
        with pytest.raises(CommitTimeout):
            pc.check()
        assert (host(out) == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    finally:
        pc.close()


def test_peer_commit_device_epochs_in_a_cuda_graph():
    """epoch = 0: the kernels count the commitments themselves, so the captured step can be replayed."""
    import torch
    import stark_rings_b200 as S
    from stark_rings_b200.dist import PeerCommit
    name, kappa, m = "goldilocks", 4, 4100
    cfg, w = S.CONFIGS[name], WORDS[name]
    ctx = S.default_context(0)
    pc = PeerCommit(cfg, kappa, 1, 0, ctx, device_epochs=True)
    try:
        rows = [dev(rand_raw(name, m, i)) for i in range(kappa)]
        v = dev(rand_raw(name, m, 77))
        A = S.Matrix([S.RqNTT(cfg, r, ctx) for r in rows], ctx)
        out = torch.empty(kappa * w, dtype=torch.int64, device="cuda")
        step = lambda: pc.commit(A, S.RqNTT(cfg, v, ctx), out=out)
        step()  # warm-up: row table upload, counters
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        for it in range(9):  # past twice the mailbox depth
            new_rows = [rand_raw(name, m, 1000 * it + i) for i in range(kappa)]
            new_v = rand_raw(name, m, 1000 * it + 99)
            for r, nr in zip(rows, new_rows):
                r.copy_(dev(nr))
            v.copy_(dev(new_v))
            g.replay()
            torch.cuda.synchronize()
            assert np.array_equal(host(out), C.matvec(name, new_rows, new_v, threads=8)), it
        ctx.use_torch_stream()
        pc.check()
    finally:
        pc.close()


def _worker(rank, world, port, name, kappa, m, q, device_epochs=False, fused=True):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import stark_rings_b200 as S
    from stark_rings_b200.dist import PeerCommit, shard_columns
    cfg, w = S.CONFIGS[name], WORDS[name]
    ctx = S.Context(rank)
    pc = PeerCommit(cfg, kappa, world, rank, ctx, device_epochs=device_epochs, fused=fused)
    ok = True
    for epoch in range(1, 10):
        rows = [rand_raw(name, m, 40 + i + 100 * epoch) for i in range(kappa)]
        v = rand_raw(name, m, 50 + epoch)
        lo, hi = shard_columns(m, world, rank)
        d = "cuda:%d" % rank
        A = S.Matrix([S.RqNTT(cfg, dev(x[lo * w:hi * w].copy(), d), ctx) for x in rows], ctx)
        out = pc.commit(A, S.RqNTT(cfg, dev(v[lo * w:hi * w].copy(), d), ctx))
        torch.cuda.synchronize()
        if rank == 0:
            ok = ok and np.array_equal(host(out), C.matvec(name, rows, v, threads=4))
    ok = ok and not pc.timed_out()
    dist.barrier()
    pc.close()
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("name,device_epochs,fused", [("goldilocks", False, True), ("babybear", False, True),
                                                      ("goldilocks", True, True), ("goldilocks", False, False),
                                                      ("stark_prime", True, True)])
def test_peer_commit_two_processes(name, device_epochs, fused):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, name, 4, 40000, q, device_epochs, fused)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
