"""Pins oracle/ref_py.py against the reference's own vectors (tests/golden/, made by
tools/gen_golden.py from the reference's inline #[test] literals and slot-map source)."""
import random

import pytest

from oracle import ref_py as O

PHI3 = ["goldilocks", "babybear"]
ALL = PHI3 + ["stark_prime"]


def ints(xs):
    return [int(x) for x in xs]


@pytest.mark.parametrize("name", PHI3)
def test_constants_match_reference_tables(name, golden):
    g, M = golden(name), O.MODELS[name]
    assert M.p == int(g["p"]) and M.D == g["D"]
    assert M.W == ints(g["roots"])          # also: table == powers of W[1] (ntt.rs test_roots_of_unity)
    assert M.r == int(g["nonresidue"]) == M.W[1]
    assert M.KAPPA == int(g["KAPPA"])
    assert M.EIGHT_INV == int(g["EIGHT_INV"]) and M.FOUR_INV == int(g["FOUR_INV"])


def test_stark_constants(golden):
    g, M = golden("stark_prime"), O.STARK
    assert M.p == int(g["p"])
    assert M.W == ints(g["roots"])
    assert M.SIXTEEN_INV == int(g["SIXTEEN_INV"])
    assert M.SIXTEEN_INV_W24 == int(g["SIXTEEN_INV_TIMES_ROOT_OF_UNITY_32_24"])


@pytest.mark.parametrize("name", PHI3)
def test_crt_icrt_kats(name, golden):
    """goldilocks/ntt.rs:564-787 (4 KATs), babybear/ntt.rs:866-1019 (72-value ICRT KAT)."""
    g, M = golden(name), O.MODELS[name]
    assert g["crt_kats"]
    for kat in g["crt_kats"]:
        coeffs, rem = ints(kat["coeffs"]), ints(kat["slot_remainders"])
        assert M.crt_stages(coeffs) == rem                      # crt_in_place ; dehomogenize
        assert M.dehomogenize(M.crt(coeffs)) == rem
        assert M.icrt(M.homogenize(rem)) == coeffs              # homogenize ; icrt_in_place
        assert M.icrt_stages(rem) == coeffs


def test_stark_kats(golden):
    """stark_prime/ntt.rs:377-545."""
    g, M = golden("stark_prime"), O.STARK
    assert len(g["crt_kats"]) == 2
    for kat in g["crt_kats"]:
        assert M.crt(ints(kat["coeffs"])) == ints(kat["evaluations"])
        assert M.icrt(ints(kat["evaluations"])) == ints(kat["coeffs"])


def test_stark_crt_is_evaluation_in_documented_order():
    """stark_prime/ntt.rs:57-64."""
    M = O.STARK
    rng = random.Random(1)
    c = [rng.randrange(M.p) for _ in range(16)]
    ev = M.crt(c)
    for i, o in enumerate(M.eval_order):
        x = M.W[o]
        assert ev[i] == sum(ci * pow(x, k, M.p) for k, ci in enumerate(c)) % M.p


@pytest.mark.parametrize("name", PHI3)
def test_slot_isomorphisms_match_reference_source(name, golden):
    """homogenize_* / dehomogenize_* interpreted from the reference source on random and unit
    vectors (goldilocks/ntt.rs:326-437, babybear/ntt.rs:324-588)."""
    g, M = golden(name), O.MODELS[name]
    assert len(g["homogenize"]) >= 8 + M.D
    for case in g["homogenize"]:
        assert M.homogenize(ints(case["in"])) == ints(case["out"])
    for case in g["dehomogenize"]:
        assert M.dehomogenize(ints(case["in"])) == ints(case["out"])


@pytest.mark.parametrize("name", PHI3)
def test_crt_slots_are_remainders(name):
    """babybear/ntt.rs:764-857: slot s of crt_stages(f) == f mod X^d - r^k."""
    M = O.MODELS[name]
    rng = random.Random(2)
    f = [rng.randrange(M.p) for _ in range(M.D)]
    st = M.crt_stages(f)
    for s, k in enumerate(O.SLOT_K):
        rem = [0] * M.d
        for i, c in enumerate(f):
            rem[i % M.d] = (rem[i % M.d] + c * pow(M.W[k], i // M.d, M.p)) % M.p
        assert st[s * M.d:(s + 1) * M.d] == rem


@pytest.mark.parametrize("name", ALL)
def test_crt_one_and_roundtrip(name):
    """models/*/mod.rs test_crt_one / test_icrt_one; crt.rs:85-122 round trips."""
    M = O.MODELS[name]
    one = [1] + [0] * (M.D - 1)
    ntt_one = []
    for _ in range(M.D // M.d):
        ntt_one += [1] + [0] * (M.d - 1)
    assert M.crt(one) == ntt_one and M.icrt(ntt_one) == one
    rng = random.Random(3)
    for _ in range(20):
        a = [rng.randrange(M.p) for _ in range(M.D)]
        assert M.icrt(M.crt(a)) == a and M.crt(M.icrt(a)) == a


@pytest.mark.parametrize("name", ALL)
def test_mul_crt(name):
    """models/*/mod.rs test_mul_crt: icrt(crt(a)*crt(b)) == schoolbook a*b mod Phi."""
    M = O.MODELS[name]
    rng = random.Random(4)
    for _ in range(3):
        a = [rng.randrange(M.p) for _ in range(M.D)]
        b = [rng.randrange(M.p) for _ in range(M.D)]
        assert O.ring_mul(M, a, b) == O.poly_mul(M, a, b)


def test_babybear_fq9_layout():
    """babybear/ntt.rs:716-742: Fq9 product == polynomial product mod X^9 - r under the
    (1 3)(2 6)(5 7) permutation; and the tower (Fq3[Y]/(Y^3-u)) product agrees."""
    M = O.BABYBEAR
    p, r = M.p, M.r
    rng = random.Random(5)
    x = [rng.randrange(p) for _ in range(9)]
    y = [rng.randrange(p) for _ in range(9)]

    def f3mul(a, b):
        return [(a[0] * b[0] + r * (a[1] * b[2] + a[2] * b[1])) % p,
                (a[0] * b[1] + a[1] * b[0] + r * a[2] * b[2]) % p,
                (a[0] * b[2] + a[1] * b[1] + a[2] * b[0]) % p]

    def f3add(a, b):
        return [(s + t) % p for s, t in zip(a, b)]

    def mulu(a):  # fq9.rs:19-26
        return [r * a[2] % p, a[0], a[1]]

    X = [x[0:3], x[3:6], x[6:9]]
    Y = [y[0:3], y[3:6], y[6:9]]
    c0 = f3add(f3mul(X[0], Y[0]), mulu(f3add(f3mul(X[1], Y[2]), f3mul(X[2], Y[1]))))
    c1 = f3add(f3add(f3mul(X[0], Y[1]), f3mul(X[1], Y[0])), mulu(f3mul(X[2], Y[2])))
    c2 = f3add(f3add(f3mul(X[0], Y[2]), f3mul(X[1], Y[1])), f3mul(X[2], Y[0]))
    assert M.slot_mul(x, y) == c0 + c1 + c2


def test_matvec_semantics():
    """matrix.rs:168-183 and :232-243 (length mismatch -> None)."""
    M = O.GOLDILOCKS
    rng = random.Random(6)
    rows = [[[rng.randrange(M.p) for _ in range(24)] for _ in range(3)] for _ in range(2)]
    v = [[rng.randrange(M.p) for _ in range(24)] for _ in range(3)]
    y = O.matvec(M, rows, v)
    assert len(y) == 2
    assert O.matvec(M, rows, v[:2]) is None
    # linearity in v
    v2 = [[rng.randrange(M.p) for _ in range(24)] for _ in range(3)]
    vs = [O.ntt_add(M, s, t) for s, t in zip(v, v2)]
    y2 = O.matvec(M, rows, v2)
    assert O.matvec(M, rows, vs) == [O.ntt_add(M, s, t) for s, t in zip(y, y2)]


@pytest.mark.parametrize("name", ALL)
def test_raw_limbs_roundtrip(name):
    M = O.MODELS[name]
    rng = random.Random(7)
    a = [rng.randrange(M.p) for _ in range(M.D)]
    raw = O.to_raw(M, a)
    assert len(raw) == M.D * M.limbs and all(0 <= w < 2**64 for w in raw)
    assert O.from_raw(M, raw) == a
