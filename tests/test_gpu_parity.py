"""GPU parity: the CUDA path, called through the C ABI (ctypes -> libstarkrings_cuda.so), against
the oracle on identical seeded inputs.  Bit-exact (integer arithmetic): np.array_equal everywhere."""
import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import ref_py as O
from tests.util import WORDS, rand_raw

pytestmark = pytest.mark.gpu

ALL = ["goldilocks", "babybear", "stark_prime"]


@pytest.fixture(scope="module")
def S():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import stark_rings_b200 as S
    S.default_context(0)  # raises if the extension cannot drive the GPU
    return S


def dev(a):
    import torch
    return torch.from_numpy(a.view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n", [1, 2, 63, 129, 1000, 4099])
def test_crt_icrt_device(S, name, n):
    cfg = S.CONFIGS[name]
    a = rand_raw(name, n, 100 + n)
    want = C.crt(name, a.copy(), threads=8)
    d = dev(a)
    cfg.crt_batch(d)
    assert np.array_equal(host(d), want)
    cfg.icrt_batch(d)
    assert np.array_equal(host(d), a)
    # icrt of arbitrary data, too
    want_i = C.icrt(name, a.copy(), threads=8)
    d = dev(a)
    cfg.icrt_batch(d)
    assert np.array_equal(host(d), want_i)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n", [1, 129, 2050])
def test_ntt_mul_and_ring_mul_device(S, name, n):
    cfg = S.CONFIGS[name]
    a, b = rand_raw(name, n, 200 + n), rand_raw(name, n, 300 + n)
    w = WORDS[name]
    if n >= 3:
        b[: 2 * w] = a[w: 3 * w]  # 0 * (p-1), (p-1) * random
    want_nm = C.ntt_mul(name, a.copy(), b.copy(), threads=8)
    da, db = dev(a), dev(b)
    cfg.ntt_mul_batch(da, db)
    assert np.array_equal(host(da), want_nm)
    assert np.array_equal(host(db), b)
    want_rm = C.ring_mul(name, a, b, threads=8)
    da = dev(a)
    out = cfg.ring_mul_batch(da, db)
    assert np.array_equal(host(out), want_rm)
    assert np.array_equal(host(da), a)
    # aliasing: out = a
    cfg.ring_mul_batch(da, db, out=da)
    assert np.array_equal(host(da), want_rm)


@pytest.mark.parametrize("name", ALL)
def test_host_buffers_pipeline(S, name):
    """SR_HOST path: chunked H2D -> kernel -> D2H; n chosen to span several 64 MiB chunks."""
    cfg = S.CONFIGS[name]
    n = (3 * (64 << 20)) // (WORDS[name] * 8) + 77
    a, b = rand_raw(name, n, 7), rand_raw(name, n, 8)
    a0 = a.copy()
    out = cfg.ring_mul_batch(a, b)
    assert np.array_equal(a, a0)
    # sample-check against the oracle (full check would take the CPU minutes)
    w = WORDS[name]
    idx = np.r_[0:64, n // 2:n // 2 + 64, n - 64:n]
    sel = lambda x: np.concatenate([x[i * w:(i + 1) * w] for i in idx])
    assert np.array_equal(sel(out), C.ring_mul(name, sel(a), sel(b), threads=8))
    # whole-buffer property: crt then icrt is the identity
    cfg.crt_batch(a)
    assert not np.array_equal(a, a0)
    cfg.icrt_batch(a)
    assert np.array_equal(a, a0)


@pytest.mark.parametrize("name", ALL)
def test_golden_and_kats_on_device(S, name, golden):
    cfg, M = S.CONFIGS[name], O.MODELS[name]
    g = golden(name)
    for kat in g["crt_kats"]:
        coeffs = [int(x) for x in kat["coeffs"]]
        raw = np.array(O.to_raw(M, coeffs), dtype=np.uint64)
        p = S.RqPoly(cfg, dev(raw))
        ntt = p.crt()
        got = O.from_raw(M, host(ntt.data).tolist())
        key = "evaluations" if name == "stark_prime" else "slot_remainders"
        assert M.dehomogenize(got) == [int(x) for x in kat[key]]
        back = ntt.icrt()
        assert O.from_raw(M, host(back.data).tolist()) == coeffs
    gd = golden(name + "_derived")
    for case in gd["cases"]:
        a = np.array([int(x) for x in case["a_raw"]], dtype=np.uint64)
        b = np.array(O.to_raw(M, [int(x) for x in case["b"]]), dtype=np.uint64)
        prod = S.RqPoly(cfg, dev(a)) * S.RqPoly(cfg, dev(b))
        assert host(prod.data).tolist() == [int(x) for x in case["ring_mul_raw"]]
        assert host(cfg.crt(dev(a))).tolist() == [int(x) for x in case["crt_a_raw"]]
    # crt(ONE) == ONE (models/*/mod.rs test_crt_one)
    one = np.array(O.to_raw(M, [1] + [0] * (M.D - 1)), dtype=np.uint64)
    ntt_one = []
    for _ in range(M.D // M.d):
        ntt_one += [1] + [0] * (M.d - 1)
    assert host(cfg.crt(dev(one))).tolist() == O.to_raw(M, ntt_one)


@pytest.mark.parametrize("name", ALL)
def test_length_errors_and_empty(S, name):
    cfg = S.CONFIGS[name]
    import torch
    bad = torch.zeros(cfg.limbs + 1, dtype=torch.int64, device="cuda")
    with pytest.raises(S.LengthPanic):
        cfg.crt_batch(bad)
    with pytest.raises(S.LengthPanic):
        cfg.crt_in_place(torch.zeros(2 * cfg.limbs, dtype=torch.int64, device="cuda"))
    with pytest.raises(S.LengthPanic):
        cfg.icrt_batch(np.zeros(cfg.limbs - 1, dtype=np.uint64))
    empty = torch.zeros(0, dtype=torch.int64, device="cuda")
    cfg.crt_batch(empty)
    cfg.ring_mul_batch(empty, empty)
    assert cfg.crt_batch(np.zeros(0, dtype=np.uint64)).size == 0


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("kappa,m", [(1, 1), (3, 7), (4, 1000), (5, 2049), (9, 300)])
def test_matvec_device(S, name, kappa, m):
    cfg = S.CONFIGS[name]
    rows = [rand_raw(name, m, 1000 + 17 * i + m) for i in range(kappa)]
    v = rand_raw(name, m, 999 + m)
    want = C.matvec(name, rows, v, threads=8)
    A = S.Matrix([S.RqNTT(cfg, dev(r)) for r in rows])
    y = A.try_mul_vec(S.RqNTT(cfg, dev(v)))
    assert np.array_equal(host(y.data), want)
    # host buffers
    Ah = S.Matrix([S.RqNTT(cfg, r) for r in rows])
    yh = Ah @ S.RqNTT(cfg, v)
    assert np.array_equal(yh.data, want)


@pytest.mark.parametrize("name", ALL)
def test_matvec_length_mismatch(S, name):
    """matrix.rs:232-243: a too-short vector -> checked None / try Err(DifferentLengths)."""
    cfg = S.CONFIGS[name]
    rows = [rand_raw(name, 3, i) for i in range(2)]
    A = S.Matrix([S.RqNTT(cfg, dev(r)) for r in rows])
    v = S.RqNTT(cfg, dev(rand_raw(name, 2, 5)))
    assert A.checked_mul_vec(v) is None
    with pytest.raises(S.DifferentLengths) as ei:
        A.try_mul_vec(v)
    assert ei.value.lengths == (3, 2)


@pytest.mark.parametrize("name", ALL)
def test_sharded_commit_single_process(S, name):
    """Column sharding emulated on one GPU: partials of G column shards, modular sum == full product."""
    import ctypes
    import torch
    from stark_rings_b200 import _lib as L
    cfg = S.CONFIGS[name]
    kappa, m, G = 4, 1030, 4
    rows = [rand_raw(name, m, 40 + i) for i in range(kappa)]
    v = rand_raw(name, m, 50)
    want = C.matvec(name, rows, v, threads=8)
    w = WORDS[name]
    parts = []
    for r in range(G):
        lo, hi = m * r // G, m * (r + 1) // G
        A = S.Matrix([S.RqNTT(cfg, dev(x[lo * w:hi * w].copy())) for x in rows])
        parts.append(A.partial_mul_vec(S.RqNTT(cfg, dev(v[lo * w:hi * w].copy()))).data)
    gathered = torch.cat(parts)
    out = torch.empty(kappa * w, dtype=torch.int64, device="cuda")
    c = S.default_context(0)
    c.check(L.lib.sr_modsum_partials(c.h, cfg.ring_id, ctypes.c_void_p(gathered.data_ptr()), G, kappa,
                                     ctypes.c_void_p(out.data_ptr()), L.SR_DEVICE), "modsum")
    assert np.array_equal(host(out), want)


@pytest.mark.parametrize("name,n", [("goldilocks", 1 << 20), ("babybear", 1 << 19), ("stark_prime", 1 << 18)])
def test_large_batch_properties(S, name, n):
    """Size-independent properties at large n: round trip, commutativity, multiplication by ONE,
    distributivity over a sampled oracle check."""
    import torch
    cfg, M = S.CONFIGS[name], O.MODELS[name]
    a, b = rand_raw(name, n, 1), rand_raw(name, n, 2)
    da, db = dev(a), dev(b)
    ab = cfg.ring_mul_batch(da, db)
    ba = cfg.ring_mul_batch(db, da)
    assert torch.equal(ab, ba)
    t = da.clone()
    cfg.crt_batch(t)
    cfg.icrt_batch(t)
    assert torch.equal(t, da)
    one = np.tile(np.array(O.to_raw(M, [1] + [0] * (M.D - 1)), dtype=np.uint64), n)
    assert torch.equal(cfg.ring_mul_batch(da, dev(one)), da)
    w = WORDS[name]
    idx = np.r_[0:32, n - 32:n]
    sel = lambda x: np.concatenate([x[i * w:(i + 1) * w] for i in idx])
    assert np.array_equal(sel(host(ab)), C.ring_mul(name, sel(a), sel(b), threads=8))


@pytest.mark.parametrize("name,log2n", [("goldilocks", 16), ("babybear", 24), ("stark_prime", 20)])
def test_baseline_config_sizes_properties(S, name, log2n):
    """BASELINE.json configs 1-3 at their full sizes (Goldilocks 2^16, BabyBear 2^24, Starknet 2^20), inputs
    generated on the device: CRT/ICRT round trip, NTT-form product == fused product, commutativity,
    multiplication by ONE, and an oracle check on a sample of elements."""
    import torch
    from bench import gen_raw_device
    cfg, M = S.CONFIGS[name], O.MODELS[name]
    tag = cfg.tag
    n = 1 << log2n
    devc = torch.device("cuda", 0)
    a = gen_raw_device(torch, tag, n, 101, devc)
    b = gen_raw_device(torch, tag, n, 202, devc)
    ab = cfg.ring_mul_batch(a, b)
    assert torch.equal(ab, cfg.ring_mul_batch(b, a))
    # crt -> slot-wise mul -> icrt through the three separate kernels == fused kernel
    ta, tb = a.clone(), b.clone()
    cfg.crt_batch(ta)
    cfg.crt_batch(tb)
    cfg.ntt_mul_batch(ta, tb)
    cfg.icrt_batch(ta)
    assert torch.equal(ta, ab)
    del ta
    cfg.icrt_batch(tb)
    assert torch.equal(tb, b)
    del tb
    one = torch.from_numpy(np.array(O.to_raw(M, [1] + [0] * (M.D - 1)), dtype=np.uint64).view(np.int64)).to(devc)
    assert torch.equal(cfg.ring_mul_batch(a, one.repeat(n)), a)
    w = WORDS[name]
    idx = torch.tensor([0, 1, n // 3, n - 2, n - 1], device=devc)
    cols = (idx[:, None] * w + torch.arange(w, device=devc)[None, :]).reshape(-1)
    sa, sb, sab = (host(x[cols]) for x in (a, b, ab))
    assert np.array_equal(sab, C.ring_mul(name, sa, sb, threads=4))


@pytest.mark.parametrize("name", ALL)
def test_reduce_and_rot_device(S, name):
    """batched reduce_in_place and Cyclotomic::rot, device and host buffers."""
    cfg, M = S.CONFIGS[name], O.MODELS[name]
    n = 1000
    for length in (M.D, M.D + 7, 2 * M.D - 1, 2 * M.D):
        rng = np.random.default_rng(length)
        per = length * M.limbs
        src = rand_raw(name, 2 * n, 70 + length)[: n * per].copy()
        want = C.reduce(name, src, length)
        assert np.array_equal(host(cfg.reduce_batch(dev(src), length)), want)
        assert np.array_equal(cfg.reduce_batch(src, length), want)
    a = rand_raw(name, n, 81)
    want = C.rot(name, a)
    p = S.RqPoly(cfg, dev(a))
    assert np.array_equal(host(p.rot().data), want)
    assert np.array_equal(cfg.rot_batch(a), want)
    # rot == multiplication by the monomial X through the fused ring-mul kernel
    x = np.tile(np.array(O.to_raw(M, [0, 1] + [0] * (M.D - 2)), dtype=np.uint64), n)
    assert np.array_equal(host(cfg.ring_mul_batch(dev(a), dev(x))), want)
    with pytest.raises(S.LengthPanic):
        cfg.reduce_batch(dev(a), M.D - 1)


@pytest.mark.parametrize("name", ["goldilocks", "babybear"])
@pytest.mark.parametrize("b,pad", [(2, 66), (4, 34), (16, 18), (1 << 16, 5), (10, 21)])
def test_gadget_decompose_device(S, name, b, pad):
    """balanced gadget decomposition / recomposition vs the oracle, device and host buffers, and the
    decompose -> CRT -> commit chain it feeds."""
    cfg = S.CONFIGS[name]
    if name == "babybear" and b == 2:
        pad = 34
    n = 300
    a = rand_raw(name, n, 900 + b)
    want = C.gadget_decompose(name, a, b, pad)
    got = cfg.gadget_decompose(dev(a), b, pad)
    assert np.array_equal(host(got), want)
    assert np.array_equal(cfg.gadget_decompose(a, b, pad), want)
    back = cfg.gadget_recompose(got, b, pad)
    assert np.array_equal(host(back), a)
    assert np.array_equal(cfg.gadget_recompose(want, b, pad), a)
    with pytest.raises(S.LengthPanic):
        cfg.gadget_decompose(dev(a), b, 1)
    with pytest.raises(S.StarkRingsError):
        cfg.gadget_decompose(dev(a), 3, pad)  # odd basis: the reference asserts


def test_gadget_decompose_reference_kat_device(S):
    """balanced_decomposition/mod.rs:470-514 on the GPU."""
    cfg, M = S.CONFIGS["goldilocks"], O.GOLDILOCKS
    elems = np.array([w for e in ([15] * 24, [M.p - 15] * 24) for w in O.to_raw(M, e)], dtype=np.uint64)
    want = np.array([w for e in ([[1] * 24] * 4 + [[M.p - 1] * 24] * 4) for w in O.to_raw(M, e)], dtype=np.uint64)
    assert np.array_equal(host(cfg.gadget_decompose(dev(elems), 2, 4)), want)
    assert np.array_equal(host(cfg.gadget_recompose(dev(want), 2, 4)), elems)
    with pytest.raises(S.StarkRingsError):
        S.CONFIGS["stark_prime"].gadget_decompose(dev(rand_raw("stark_prime", 2, 1)), 2, 300)


@pytest.mark.parametrize("name", ["goldilocks", "babybear"])
def test_decompose_crt_commit_chain_on_device(S, name):
    """The pipeline the hot path serves (decompose -> CRT -> commit) without leaving the device, against the
    oracle composing the same three steps; plus linearity: A * recompose-weights recovers A * crt(w)."""
    cfg, M = S.CONFIGS[name], O.MODELS[name]
    w = WORDS[name]
    n_wit, b, pad, kappa = 40, 1 << 8, 9 if name == "goldilocks" else 5, 3
    wit = rand_raw(name, n_wit, 321)
    m = n_wit * pad
    rows = [rand_raw(name, m, 400 + i) for i in range(kappa)]
    # oracle
    dec = C.gadget_decompose(name, wit, b, pad)
    want = C.matvec(name, rows, C.crt(name, dec.copy(), threads=4), threads=4)
    # device
    d = cfg.gadget_decompose(dev(wit), b, pad)
    v = S.RqPoly(cfg, d).crt()
    A = S.Matrix([S.RqNTT(cfg, dev(r)) for r in rows])
    y = A.try_mul_vec(v)
    assert np.array_equal(host(y.data), want)


def test_concurrent_host_threads_one_context_each(S):
    """SURVEY 8b threading row: with `parallel`, rayon workers call the ring operations concurrently
    (linear_algebra/src/matrix.rs:174).  The library's contract is one context per host thread (plus an internal mutex
    per context); four threads, each with its own context and stream, must all get the oracle's bits."""
    import threading
    import torch
    name = "goldilocks"
    cfg = S.CONFIGS[name]
    results, errors = {}, []

    def worker(k):
        try:
            ctx = S.Context(0)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for it in range(5):
                    a, b = rand_raw(name, 3000 + k, 10 * k + it), rand_raw(name, 3000 + k, 10 * k + it + 5)
                    out = cfg.ring_mul_batch(dev(a), dev(b), ctx=ctx)
                    y = cfg.crt_batch(dev(a), ctx=ctx)
                    stream.synchronize()
                    ok = np.array_equal(host(out), C.ring_mul(name, a, b, threads=1)) and \
                        np.array_equal(host(y), C.crt(name, a.copy(), threads=1))
                    results[(k, it)] = ok
            ctx.close()
        except Exception as ex:  # pragma: no cover
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 20 and all(results.values())
    # and one SHARED context used from several threads: calls are serialised by its mutex, results still exact
    shared = S.Context(0)
    res2 = {}

    def worker2(k):
        a = rand_raw(name, 500, 77 + k)
        h = a.copy()
        cfg.crt_batch(h, ctx=shared)  # host-buffer path: synchronous
        res2[k] = np.array_equal(h, C.crt(name, a.copy(), threads=1))

    threads = [threading.Thread(target=worker2, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    shared.close()
    assert len(res2) == 4 and all(res2.values())


def _boundary_raw(name, n, seed):
    """Elements whose raw limbs are drawn from the values where the kernels' lazy arithmetic changes branch: 0, 1, p - 1,
    p - 2, and the word / carry boundaries of each prime (2^32 - 1, 2^32, the high word all ones, 2^63, ...)."""
    import random
    M = O.MODELS[name]
    p = M.p
    if name == "goldilocks":
        eps = (1 << 32) - 1
        vals = [0, 1, 2, eps - 1, eps, eps + 1, 1 << 32, (1 << 63), (1 << 63) - 1, p - 1, p - 2, p - eps, p - eps - 1,
                0xFFFFFFFF00000000, 0xFFFFFFFEFFFFFFFF, 0xFFFFFFFE00000001, 0x00000001FFFFFFFF, 0x8000000080000000]
    elif name == "babybear":
        vals = [0, 1, 2, p - 1, p - 2, 1 << 27, (1 << 27) - 1, 15 << 27, (1 << 30), (1 << 30) - 1, 0x77FFFFFF, 0x40000001]
    else:
        vals = [0, 1, 2, p - 1, p - 2, (1 << 251), (1 << 251) - 1, (1 << 192) * 17, (1 << 192) * 17 - 1, (1 << 224) - 1,
                (1 << 128) - 1, (1 << 250) + (1 << 64) - 1, p - (1 << 192), p >> 1]
    vals = [v % p for v in vals]
    rng = random.Random(seed)
    nl = 4 if name == "stark_prime" else 1  # u64 limbs per coefficient
    out = []
    for e in range(n):
        for c in range(M.D):
            v = vals[(e + c) % len(vals)] if e < len(vals) else rng.choice(vals)
            for k in range(nl):
                out.append((v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF)
    return np.array(out, dtype=np.uint64)


@pytest.mark.parametrize("name", ALL)
def test_boundary_values_all_ops(S, name):
    """crt / icrt / ntt_mul / ring_mul / mat-vec on raw limbs at the word and carry boundaries of each prime."""
    cfg = S.CONFIGS[name]
    n = 700
    a, b = _boundary_raw(name, n, 1), _boundary_raw(name, n, 3)
    b[: WORDS[name] * 64] = a[WORDS[name] * 3: WORDS[name] * 67]  # the deterministic patterns against each other, shifted
    for op, want in (("crt", C.crt(name, a.copy(), threads=8)), ("icrt", C.icrt(name, a.copy(), threads=8))):
        d = dev(a)
        getattr(cfg, op + "_batch")(d)
        assert np.array_equal(host(d), want), op
    da, db = dev(a), dev(b)
    cfg.ntt_mul_batch(da, db)
    assert np.array_equal(host(da), C.ntt_mul(name, a.copy(), b.copy(), threads=8))
    out = cfg.ring_mul_batch(dev(a), db)
    assert np.array_equal(host(out), C.ring_mul(name, a, b, threads=8))
    ctx = S.default_context(0)
    rows = [_boundary_raw(name, n, 10 + i) for i in range(3)]
    want = C.matvec(name, [r.copy() for r in rows], a.copy(), threads=3)
    A = S.Matrix([S.RqNTT(cfg, dev(r), ctx) for r in rows], ctx)
    got = A.try_mul_vec(S.RqNTT(cfg, dev(a), ctx))
    assert np.array_equal(host(got.data), want)
