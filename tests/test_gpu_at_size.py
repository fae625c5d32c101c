"""Whole-output oracle comparisons at the BASELINE sizes.

The Goldilocks commit kernel streams 256- / 384-slot chunks through a 3-stage TMA ring, 148 - 888 CTAs depending on
the row count: a CTA only gets a second chunk and only recycles a stage (the `empty` mbarriers, the phase-parity
arithmetic) for m of a few tens of thousands of columns; BASELINE config 4 (m = 2^20) gives every CTA 55 - 110
chunks.  The BabyBear / Starknet kernels wrap their grid stride at m > 18 944 / 9 472.  Every case below compares the WHOLE output with the C oracle's mat-vec
(reference semantics: linear_algebra/src/matrix.rs:168-178), bit for bit.  The batch kernels are compared on whole
2^20-element buffers for all three rings."""
import numpy as np
import pytest

from oracle import c_oracle as C
from tests.util import WORDS, rand_raw

pytestmark = pytest.mark.gpu

ALL = ["goldilocks", "babybear", "stark_prime"]


@pytest.fixture(scope="module")
def S():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import stark_rings_b200 as S
    S.default_context(0)
    return S


def dev(a):
    import torch
    return torch.from_numpy(a.view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def _inputs(name, kappa, m, seed):
    rows = [rand_raw(name, m, seed + 31 * i) for i in range(kappa)]
    v = rand_raw(name, m, seed + 7)
    return rows, v


def _check_matvec(S, name, kappa, m, seed):
    cfg = S.CONFIGS[name]
    rows, v = _inputs(name, kappa, m, seed)
    want = C.matvec(name, rows, v, threads=min(kappa, 16))
    A = S.Matrix([S.RqNTT(cfg, dev(r)) for r in rows])
    y = A.try_mul_vec(S.RqNTT(cfg, dev(v)))
    assert np.array_equal(host(y.data), want)
    return rows, v, want, A


# 2 chunks per CTA / stage recycling / ragged tail past a power of two
@pytest.mark.parametrize("m", [4737, 18945, (1 << 17) + 3])
@pytest.mark.parametrize("kappa", [1, 2, 3, 4, 5, 8, 16])
def test_goldilocks_commit_multichunk(S, kappa, m):
    _check_matvec(S, "goldilocks", kappa, m, 5000 + kappa + m)


@pytest.mark.parametrize("kappa", [1, 2, 3, 4, 5, 8, 16])
def test_goldilocks_commit_baseline_config4(S, kappa):
    """BASELINE config 4: kappa x 2^20 Goldilocks matrix x vector, whole output against the oracle."""
    _check_matvec(S, "goldilocks", kappa, 1 << 20, 9000 + kappa)


@pytest.mark.parametrize("name,m", [("babybear", 1 << 15), ("babybear", (1 << 17) + 5), ("stark_prime", 1 << 15),
                                    ("stark_prime", (1 << 17) + 5)])
@pytest.mark.parametrize("kappa", [1, 2, 3, 4, 5, 7])  # one / two row groups per CTA, a group with a missing row, 4 + 1, 4 + 3
def test_babybear_starknet_commit_grid_wrap(S, name, m, kappa):
    _check_matvec(S, name, kappa, m, 7000 + kappa + m)


@pytest.mark.parametrize("name,kappa,m,G", [("goldilocks", 4, 1 << 20, 8), ("goldilocks", 5, (1 << 18) + 9, 2),
                                            ("babybear", 4, 1 << 16, 4), ("stark_prime", 3, 1 << 15, 2)])
def test_peer_commit_emulated_ranks_at_size(S, name, kappa, m, G):
    """The column-sharded commit (mailbox writer / root kernels) with every shard large enough to recycle the stage
    ring; G ranks emulated in one process on one GPU (one kernel after the other: nothing ever waits)."""
    import torch
    from stark_rings_b200.dist import PeerCommit, shard_columns
    cfg, w = S.CONFIGS[name], WORDS[name]
    ctx = S.default_context(0)
    rows, v = _inputs(name, kappa, m, 12000 + m)
    want = C.matvec(name, rows, v, threads=min(kappa, 16))
    pc = PeerCommit(cfg, kappa, G, 0, ctx, exchange=lambda h: h)
    try:
        shards = []
        for r in range(G):
            lo, hi = shard_columns(m, G, r)
            shards.append((S.Matrix([S.RqNTT(cfg, dev(x[lo * w:hi * w].copy())) for x in rows]),
                           S.RqNTT(cfg, dev(v[lo * w:hi * w].copy()))))
        for epoch in range(3):
            pc.epoch += 1
            out = torch.empty(kappa * w, dtype=torch.int64, device="cuda")
            for r in range(1, G):
                pc.send(*shards[r], as_rank=r)
            if epoch == 1:  # two-kernel mode: the root sends as well, then the separate reduction
                pc.send(*shards[0], as_rank=0)
                pc.reduce(kappa, out)
            else:           # fused: the root's product kernel sums all ranks in its tail
                pc.root_commit(shards[0][0], shards[0][1], out, as_rank=0)
            assert np.array_equal(host(out), want), epoch
        assert not pc.timed_out()
    finally:
        pc.close()


@pytest.mark.parametrize("name", ALL)
def test_whole_buffer_batch_ops_2p20(S, name):
    """crt / icrt / ntt_mul / ring_mul on 2^20 elements: every output limb against the C oracle."""
    cfg = S.CONFIGS[name]
    n = 1 << 20
    a, b = rand_raw(name, n, 31337), rand_raw(name, n, 31338)
    T = 16
    da, db = dev(a), dev(b)
    out = cfg.ring_mul_batch(da, db)
    assert np.array_equal(host(out), C.ring_mul(name, a, b, threads=T))
    del out
    t = da.clone()
    cfg.crt_batch(t)
    assert np.array_equal(host(t), C.crt(name, a.copy(), threads=T))
    t = da.clone()
    cfg.icrt_batch(t)
    assert np.array_equal(host(t), C.icrt(name, a.copy(), threads=T))
    t = da.clone()
    cfg.ntt_mul_batch(t, db)
    assert np.array_equal(host(t), C.ntt_mul(name, a.copy(), b.copy(), threads=T))


@pytest.mark.parametrize("name", ALL)
def test_pipelined_products_back_to_back(S, name):
    """sr_set_pipelined: successive products of one context overlap (the column loop of product i + 1 runs while
    product i is in its tail).  Twenty products enqueued
    without any synchronisation, alternating two matrices of different shapes, every output against the oracle."""
    import torch
    cfg = S.CONFIGS[name]
    ctx = S.Context(0)
    ctx.set_pipelined(True)
    try:
        shapes = [(4, 70001), (5, 9000)] if name == "goldilocks" else [(4, 9001), (3, 2000)]
        mats = []
        for i, (kappa, m) in enumerate(shapes):
            rows, v = _inputs(name, kappa, m, 777 + i)
            want = C.matvec(name, rows, v, threads=kappa)
            A = S.Matrix([S.RqNTT(cfg, dev(r), ctx) for r in rows], ctx)
            mats.append((A, S.RqNTT(cfg, dev(v), ctx), want))
        torch.cuda.synchronize()
        outs = []
        for it in range(20):
            A, v, want = mats[it % 2]
            outs.append((A.try_mul_vec(v), want))
        torch.cuda.synchronize()
        for y, want in outs:
            assert np.array_equal(host(y.data), want)
    finally:
        ctx.close()
