"""The kernels' per-element arithmetic (stark_rings_b200/csrc/*_ring.cuh), compiled for the host,
against the oracle.  Lets the device math be validated without a GPU; the GPU parity tests
(tests/test_gpu_*.py) then validate the same code through the C ABI on the device."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import ref_py as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    d = os.path.join(HERE, "hostcheck")
    subprocess.run(["make", "-s", "-C", d], check=True)
    return ctypes.CDLL(os.path.join(d, "libhostcheck.so"))


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def rand_raw(name, n, seed):
    M = O.MODELS[name]
    rng = random.Random(seed)
    elems = [[rng.randrange(M.p) for _ in range(M.D)] for _ in range(n)]
    elems[0] = [0] * M.D
    elems[1] = [M.p - 1] * M.D
    return np.array([w for e in elems for w in O.to_raw(M, e)], dtype=np.uint64)


@pytest.mark.parametrize("name,tag", [("babybear", "bb"), ("goldilocks", "gl"), ("stark_prime", "sp")])
def test_elementwise_against_oracle(hc, name, tag):
    if not hasattr(hc, "hc_%s_crt" % tag):
        pytest.skip("not built yet")
    w = C.words(name)
    n = 24
    a, b = rand_raw(name, n, 21), rand_raw(name, n, 22)
    b[: 2 * w] = a[w: 3 * w]  # (0, p-1) x (p-1, random)
    want_crt = C.crt(name, a.copy())
    want_icrt = C.icrt(name, a.copy())
    want_nm = C.ntt_mul(name, a.copy(), b.copy())
    want_rm = C.ring_mul(name, a, b)
    for i in range(n):
        ea, eb = a[i * w:(i + 1) * w].copy(), b[i * w:(i + 1) * w].copy()
        x = ea.copy(); getattr(hc, "hc_%s_crt" % tag)(_p(x))
        assert np.array_equal(x, want_crt[i * w:(i + 1) * w]), ("crt", i)
        x = ea.copy(); getattr(hc, "hc_%s_icrt" % tag)(_p(x))
        assert np.array_equal(x, want_icrt[i * w:(i + 1) * w]), ("icrt", i)
        x = ea.copy(); getattr(hc, "hc_%s_ntt_mul" % tag)(_p(x), _p(eb))
        assert np.array_equal(x, want_nm[i * w:(i + 1) * w]), ("ntt_mul", i)
        out = np.zeros(w, dtype=np.uint64)
        getattr(hc, "hc_%s_ring_mul" % tag)(_p(ea), _p(eb), _p(out))
        assert np.array_equal(out, want_rm[i * w:(i + 1) * w]), ("ring_mul", i)
        if tag == "gl":  # degree-6 formulation used by the fused kernel (gl_fused6.cuh)
            out2 = np.zeros(w, dtype=np.uint64)
            hc.hc_gl_ring_mul_fused6(_p(ea), _p(eb), _p(out2))
            assert np.array_equal(out2, want_rm[i * w:(i + 1) * w]), ("ring_mul_fused6", i)
            x = ea.copy(); hc.hc_gl_ntt_mul_rolled(_p(x), _p(eb))
            assert np.array_equal(x, want_nm[i * w:(i + 1) * w]), ("ntt_mul_rolled", i)
            x = ea.copy(); hc.hc_gl_icrt_row(_p(x))  # the row formulation the ICRT kernel runs
            assert np.array_equal(x, want_icrt[i * w:(i + 1) * w]), ("icrt_row", i)
        if tag == "sp":  # four-threads-per-element formulation used by the kernels (sp_quad.cuh)
            x = ea.copy(); hc.hc_sp_crt_quad(_p(x))
            assert np.array_equal(x, want_crt[i * w:(i + 1) * w]), ("crt_quad", i)
            x = ea.copy(); hc.hc_sp_icrt_quad(_p(x))
            assert np.array_equal(x, want_icrt[i * w:(i + 1) * w]), ("icrt_quad", i)
            x = ea.copy(); hc.hc_sp_crt_quad_lazy(_p(x))  # the unreduced schedules the CRT / ICRT kernels run
            assert np.array_equal(x, want_crt[i * w:(i + 1) * w]), ("crt_quad_lazy", i)
            x = ea.copy(); hc.hc_sp_icrt_quad_lazy(_p(x))
            assert np.array_equal(x, want_icrt[i * w:(i + 1) * w]), ("icrt_quad_lazy", i)
            out4 = np.zeros(w, dtype=np.uint64)
            hc.hc_sp_ring_mul_quad(_p(ea), _p(eb), _p(out4))
            assert np.array_equal(out4, want_rm[i * w:(i + 1) * w]), ("ring_mul_quad", i)
            out5 = np.zeros(w, dtype=np.uint64)  # the unreduced schedule the fused kernel runs (bounds trap on the host)
            hc.hc_sp_ring_mul_quad_lazy(_p(ea), _p(eb), _p(out5))
            assert np.array_equal(out5, want_rm[i * w:(i + 1) * w]), ("ring_mul_quad_lazy", i)
        if tag == "bb":  # two-threads-per-element formulation used by the fused kernel
            out2 = np.zeros(w, dtype=np.uint64)
            hc.hc_bb_ring_mul_half(_p(ea), _p(eb), _p(out2))
            assert np.array_equal(out2, want_rm[i * w:(i + 1) * w]), ("ring_mul_half", i)


def test_sp_lazy_bounds_stress(hc):
    """The unreduced Starknet schedule on many random and adversarial inputs (all p-1, alternating 0 / p-1, single
    non-zero coefficients): the host build traps on any violated bound, and the result must equal the oracle's."""
    name, w = "stark_prime", 64
    M = O.MODELS[name]
    pm1 = [((M.p - 1) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    pats = []
    for mask in (0xFFFF, 0xAAAA, 0x5555, 0x00FF, 0xFF00, 0x0001, 0x8000, 0x0101):
        e = np.zeros(w, dtype=np.uint64)
        for c in range(16):
            if (mask >> c) & 1:
                e[4 * c:4 * c + 4] = pm1
        pats.append(e)
    a = np.concatenate(pats + [rand_raw(name, 1500, 901)])
    b = np.concatenate(list(reversed(pats)) + [rand_raw(name, 1500, 902)])
    want = C.ring_mul(name, a, b, threads=4)
    want_crt, want_icrt = C.crt(name, a.copy()), C.icrt(name, a.copy())
    out = np.zeros(w, dtype=np.uint64)
    for i in range(a.size // w):
        ea, eb = a[i * w:(i + 1) * w].copy(), b[i * w:(i + 1) * w].copy()
        hc.hc_sp_ring_mul_quad_lazy(_p(ea), _p(eb), _p(out))
        assert np.array_equal(out, want[i * w:(i + 1) * w]), i
        x = ea.copy(); hc.hc_sp_crt_quad_lazy(_p(x))
        assert np.array_equal(x, want_crt[i * w:(i + 1) * w]), ("crt", i)
        x = ea.copy(); hc.hc_sp_icrt_quad_lazy(_p(x))
        assert np.array_equal(x, want_icrt[i * w:(i + 1) * w]), ("icrt", i)


def test_bb_lazy_dot(hc):
    """The unreduced BabyBear slot-product sums of the mat-vec kernels (bb::SlotAcc) against the oracle's mat-vec, on random
    columns and on all-(p-1) columns (the largest products); the host build traps when an accumulator leaves its bound."""
    name, w = "babybear", 72
    M = O.MODELS[name]
    n = 300
    a, x = rand_raw(name, n, 31), rand_raw(name, n, 32)
    worst = np.array(O.to_raw(M, [M.p - 1] * M.D) * 40, dtype=np.uint64)
    for aa, xx in ((a, x), (worst, worst), (np.concatenate([worst, a]), np.concatenate([worst, x]))):
        m = aa.size // w
        want = C.matvec(name, [aa.copy()], xx.copy(), threads=1)
        out = np.zeros(w, dtype=np.uint64)
        hc.hc_bb_dot(_p(aa), _p(xx), ctypes.c_size_t(m), _p(out))
        assert np.array_equal(out, want), m


def test_sp_lazy_dot(hc):
    """The unreduced Starknet-prime product sums of the mat-vec kernels (sp::DotAcc) against the oracle's mat-vec on random
    and all-(p-1) columns; the host build traps when the running sum would leave 2^256 or a result is not canonical."""
    name, w = "stark_prime", 64
    M = O.MODELS[name]
    a, x = rand_raw(name, 100, 41), rand_raw(name, 100, 42)
    worst = np.array(O.to_raw(M, [M.p - 1] * M.D) * 47, dtype=np.uint64)
    for aa, xx in ((a, x), (worst, worst), (np.concatenate([worst, a]), np.concatenate([worst, x]))):
        m = aa.size // w
        want = C.matvec(name, [aa.copy()], xx.copy(), threads=1)
        out = np.zeros(w, dtype=np.uint64)
        hc.hc_sp_dot(_p(aa), _p(xx), ctypes.c_size_t(m), _p(out))
        assert np.array_equal(out, want), m


def test_gl_accumulator_reductions(hc):
    """gl::acc_reduce / acc_reduce_m128 on ARBITRARY accumulator states (random and extreme limbs, not only the ones a
    product sum reaches): value = sum e_k T^k + sum o_k T^k with T = 2^32, reduced mod p and times 2^128 mod p."""
    p = O.MODELS["goldilocks"].p
    rng = random.Random(77)
    ext = [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF]
    cases = []
    for _ in range(4000):
        e = [rng.choice(ext) if rng.random() < 0.4 else rng.getrandbits(32) for _ in range(4)]
        o = [rng.choice(ext) if rng.random() < 0.4 else rng.getrandbits(32) for _ in range(3)]
        e4 = rng.choice([0, 1, 2, 5, 1 << 10, (1 << 20) - 1])  # carries out of 2^128: tiny by construction
        cases.append(e + [e4] + o)
    cases += [[0] * 8, [0xFFFFFFFF] * 4 + [0] + [0xFFFFFFFF] * 3, [1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 1]]
    out = np.zeros(3, dtype=np.uint64)
    for c in cases:
        e0, e1, e2, e3, e4, o1, o2, o3 = c
        v = e0 + ((e1 + o1) << 32) + ((e2 + o2) << 64) + ((e3 + o3) << 96) + (e4 << 128)
        limbs = np.array(c, dtype=np.uint32)
        hc.hc_gl_acc_reduce(limbs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), _p(out))
        assert int(out[0]) == v % p, c
        assert int(out[1]) == (v << 128) % p, c
        assert int(out[2]) % p == v % p, c
