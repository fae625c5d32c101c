"""oracle/sr_oracle.c (C restatement, raw Montgomery limbs) == oracle/ref_py.py (big ints)."""
import random

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import ref_py as O

ALL = ["goldilocks", "babybear", "stark_prime"]


def raw(M, vals):
    return np.array(O.to_raw(M, vals), dtype=np.uint64)


def rand_elems(M, n, rng):
    return [[rng.randrange(M.p) for _ in range(M.D)] for _ in range(n)]


def flat(M, elems):
    return np.concatenate([raw(M, e) for e in elems])


@pytest.mark.parametrize("name", ALL)
def test_golden_derived_vectors(name, golden):
    g, M = golden(name + "_derived"), O.MODELS[name]
    for case in g["cases"]:
        a = np.array([int(x) for x in case["a_raw"]], dtype=np.uint64)
        b = raw(M, [int(x) for x in case["b"]])
        want_crt = np.array([int(x) for x in case["crt_a_raw"]], dtype=np.uint64)
        want_mul = np.array([int(x) for x in case["ring_mul_raw"]], dtype=np.uint64)
        assert np.array_equal(C.crt(name, a.copy()), want_crt)
        assert np.array_equal(C.icrt(name, want_crt.copy()), a)
        assert np.array_equal(C.ring_mul(name, a, b), want_mul)
        nm = C.ntt_mul(name, want_crt.copy(), C.crt(name, b.copy()))
        assert np.array_equal(nm, raw(M, [int(x) for x in case["ntt_mul"]]))
    mv = g["matvec"]
    rows = [np.concatenate([np.array([int(x) for x in e], dtype=np.uint64) for e in row]) for row in mv["rows_raw"]]
    v = np.concatenate([np.array([int(x) for x in e], dtype=np.uint64) for e in mv["v_raw"]])
    y = np.concatenate([np.array([int(x) for x in e], dtype=np.uint64) for e in mv["y_raw"]])
    assert np.array_equal(C.matvec(name, rows, v), y)
    assert C.matvec(name, rows, v[: v.size - C.words(name)]) is None


@pytest.mark.parametrize("name", ALL)
def test_reference_kats_through_c(name, golden):
    g, M = golden(name), O.MODELS[name]
    for kat in g["crt_kats"]:
        coeffs = [int(x) for x in kat["coeffs"]]
        got = C.crt(name, raw(M, coeffs))
        want = M.crt(coeffs)
        assert O.from_raw(M, got.tolist()) == want
        key = "evaluations" if name == "stark_prime" else "slot_remainders"
        assert M.dehomogenize(want) == [int(x) for x in kat[key]]
        assert O.from_raw(M, C.icrt(name, got.copy()).tolist()) == coeffs


@pytest.mark.parametrize("name", ALL)
def test_random_batches_and_threads(name):
    M = O.MODELS[name]
    rng = random.Random(11)
    n = 37
    A, B = rand_elems(M, n, rng), rand_elems(M, n, rng)
    # edge values
    A[0] = [0] * M.D
    A[1] = [M.p - 1] * M.D
    B[1] = [M.p - 1] * M.D
    A[2] = [1] + [0] * (M.D - 1)
    a, b = flat(M, A), flat(M, B)
    want_crt = flat(M, [M.crt(x) for x in A])
    got1 = C.crt(name, a.copy(), threads=1)
    got4 = C.crt(name, a.copy(), threads=4)
    assert np.array_equal(got1, want_crt) and np.array_equal(got4, want_crt)
    assert np.array_equal(C.icrt(name, got4.copy(), threads=3), a)
    want_mul = flat(M, [O.ring_mul(M, x, y) for x, y in zip(A[:8], B[:8])])
    w = C.words(name)
    assert np.array_equal(C.ring_mul(name, a[: 8 * w].copy(), b[: 8 * w].copy(), threads=2), want_mul)
    # empty batch
    assert C.crt(name, np.zeros(0, dtype=np.uint64)).size == 0


@pytest.mark.parametrize("name", ALL)
def test_reduce_and_rot(name):
    """SURVEY 8f-2 helpers: reduce_in_place (models/*/mod.rs) and Cyclotomic::rot; rot == multiplication by X."""
    M = O.MODELS[name]
    rng = random.Random(12)
    w = C.words(name)
    nl = M.limbs
    for length in (M.D, M.D + 5, 2 * M.D - 1, 2 * M.D):
        polys = [[rng.randrange(M.p) for _ in range(length)] for _ in range(5)]
        flat_in = np.array([x for pl in polys for x in O.to_raw(M, pl)], dtype=np.uint64)
        got = C.reduce(name, flat_in, length)
        want = flat(M, [M.reduce(pl) for pl in polys])
        assert np.array_equal(got, want), length
    elems = rand_elems(M, 6, rng)
    got = C.rot(name, flat(M, elems))
    assert np.array_equal(got, flat(M, [O.rot(M, e) for e in elems]))
    x = [0, 1] + [0] * (M.D - 2)
    assert O.rot(M, elems[0]) == O.poly_mul(M, elems[0], x)   # models/*/mod.rs test_cyclotomic


def test_gadget_decompose_reference_kat():
    """balanced_decomposition/mod.rs:470-514 (test_gadget_decompose / test_gadget_recompose): 15 and -15 in
    basis 2, padding 4 -> digits (1,1,1,1) / (-1,-1,-1,-1) in every coefficient."""
    M = O.GOLDILOCKS
    elems = [[15] * 24, [M.p - 15] * 24]
    want = [[1] * 24] * 4 + [[M.p - 1] * 24] * 4
    assert O.gadget_decompose(M, elems, 2, 4) == want
    assert O.gadget_recompose(M, want, 2, 4) == elems
    got = C.gadget_decompose("goldilocks", flat(M, elems), 2, 4)
    assert np.array_equal(got, flat(M, want))
    assert np.array_equal(C.gadget_recompose("goldilocks", got, 2, 4), flat(M, elems))


@pytest.mark.parametrize("name", ["goldilocks", "babybear"])
@pytest.mark.parametrize("b", [2, 4, 8, 16, 32, 1 << 16, 10, 1 << 40])
def test_gadget_decompose_properties(name, b):
    """mod.rs:405-468: every digit within [-b/2, b/2]; recompose(decompose(v)) == v; C == Python."""
    import math
    M = O.MODELS[name]
    rng = random.Random(b)
    pad = int(math.log(M.p, b)) + 2
    elems = rand_elems(M, 5, rng)
    elems[0] = [0] * M.D
    elems[1] = [M.p - 1] * M.D
    elems[2] = [(M.p - 1) // 2] * M.D
    elems[3] = [(M.p - 1) // 2 + 1] * M.D
    d = O.gadget_decompose(M, elems, b, pad)
    for de in d:
        for x in de:
            s = x - M.p if x > (M.p - 1) // 2 else x
            assert abs(s) <= b // 2
    assert O.gadget_recompose(M, d, b, pad) == elems
    got = C.gadget_decompose(name, flat(M, elems), b, pad)
    assert np.array_equal(got, flat(M, d))
    assert np.array_equal(C.gadget_recompose(name, got, b, pad), flat(M, elems))
    if b < M.p:  # one digit cannot hold (p-1)/2
        with pytest.raises(IndexError):
            O.gadget_decompose(M, elems, b, 1)
        with pytest.raises(IndexError):
            C.gadget_decompose(name, flat(M, elems), b, 1)


@pytest.mark.parametrize("name", ["goldilocks", "babybear", "stark_prime"])
def test_addsub_sum_match_the_python_oracle(name):
    """Element-wise Add / Sub / Neg and Sum (ntt_form.rs:588-626, 640-654): the C oracle against big-int arithmetic."""
    from oracle import ref_py as O
    from tests.util import rand_raw
    M = O.MODELS[name]
    n = 7
    a, b = rand_raw(name, n, 1), rand_raw(name, n, 2)
    va, vb = O.from_raw(M, a.tolist()), O.from_raw(M, b.tolist())
    assert O.from_raw(M, C.addsub(name, "add", a, b).tolist()) == [(x + y) % M.p for x, y in zip(va, vb)]
    assert O.from_raw(M, C.addsub(name, "sub", a, b).tolist()) == [(x - y) % M.p for x, y in zip(va, vb)]
    assert O.from_raw(M, C.addsub(name, "neg", a).tolist()) == [(-x) % M.p for x in va]
    got = O.from_raw(M, C.ring_sum(name, a).tolist())
    assert got == [sum(va[e * M.D + i] for e in range(n)) % M.p for i in range(M.D)]
    assert not C.ring_sum(name, a[:0]).any()
