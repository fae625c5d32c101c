"""SURVEY 8f-3 on the CPU: the oracles' sparse mat-vec / dense mat-mat / scalar scaling.

Pinned against the reference's own known-answer tests, which use u32 entries
(linear_algebra/src/sparse_matrix.rs:305-377, matrix.rs:232-262): an integer c embeds into the ring as the constant
polynomial, whose NTT form is crt([c, 0, ..., 0]), and sums / products of constants are the constants of the integer
sums / products, so the same matrices and expected results apply verbatim over RqNTT.  Then C oracle == Python oracle
on random ring-valued inputs."""
import random

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import ref_py as O

ALL = ["goldilocks", "babybear", "stark_prime"]


def raw(M, vals):
    return np.array(O.to_raw(M, vals), dtype=np.uint64)


def flat(M, elems):
    return np.concatenate([raw(M, e) for e in elems]) if elems else np.empty(0, dtype=np.uint64)


def const(M, c):
    return M.crt([c % M.p] + [0] * (M.D - 1))


def as_int(M, e):
    """NTT-form element that is a constant -> that integer."""
    co = M.icrt(e)
    assert all(x == 0 for x in co[1:])
    return co[0]


# sparse_matrix.rs:308-316 sample_sparse / sample_dense
SAMPLE_SPARSE = [[(2, 1)], [], [(1, 0), (4, 1), (3, 2)]]
SAMPLE_DENSE = [[0, 2, 0], [0, 0, 0], [1, 4, 3]]


def csr(M, coeffs):
    row_ptr = np.zeros(len(coeffs) + 1, dtype=np.uint64)
    cols, vals = [], []
    for i, row in enumerate(coeffs):
        for e, j in row:
            vals.append(e)
            cols.append(j)
        row_ptr[i + 1] = len(cols)
    return row_ptr, np.array(cols, dtype=np.uint64), flat(M, vals)


@pytest.mark.parametrize("name", ALL)
def test_reference_kats_sparse_and_dense(name):
    M = O.MODELS[name]
    sp = [[(const(M, c), j) for c, j in row] for row in SAMPLE_SPARSE]
    v = [const(M, c) for c in (1, 2, 3)]
    # test_sparse_matrix_mul_vec (sparse_matrix.rs:340-351): [4, 0, 18]; a 2-vector is an error
    got = O.sparse_matvec(M, 3, sp, v)
    assert [as_int(M, e) for e in got] == [4, 0, 18]
    assert O.sparse_matvec(M, 3, sp, v[:2]) is None
    rp, ci, vals = csr(M, sp)
    gc = C.sparse_matvec(name, 3, 3, rp, ci, vals, flat(M, v))
    assert np.array_equal(gc, flat(M, got))
    assert C.sparse_matvec(name, 3, 3, rp, ci, vals, flat(M, v[:2])) is None
    # test_matrix_mul_vec (matrix.rs:232-243): the dense image gives the same
    dense = [[const(M, c) for c in row] for row in SAMPLE_DENSE]
    assert [as_int(M, e) for e in O.matvec(M, dense, v)] == [4, 0, 18]
    # test_matrix_mul_mat / test_sparse_matrix_mul_mat (matrix.rs:254-262): [[6, 8], [0, 0], [28, 36]]
    m2 = [[const(M, c) for c in row] for row in ([1, 2], [3, 4], [5, 6])]
    prod = O.matmat(M, dense, m2)
    assert [[as_int(M, e) for e in row] for row in prod] == [[6, 8], [0, 0], [28, 36]]
    m3 = [[const(M, c) for c in row] for row in ([1, 2], [3, 4])]
    assert O.matmat(M, dense, m3) is None
    pc = C.matmat(name, [flat(M, r) for r in dense], [flat(M, r) for r in m2])
    assert all(np.array_equal(x, flat(M, r)) for x, r in zip(pc, prod))
    assert C.matmat(name, [flat(M, r) for r in dense], [flat(M, r) for r in m3]) is None
    # test_matrix_mul_element / test_sparse_matrix_mul_element (matrix.rs:245-252): *= 3
    three = const(M, 3)
    scaled = [O.scale(M, row, three) for row in dense]
    assert [[as_int(M, e) for e in row] for row in scaled] == [[0, 6, 0], [0, 0, 0], [3, 12, 9]]
    got = C.scale(name, flat(M, dense[2]), raw(M, three))
    assert np.array_equal(got, flat(M, scaled[2]))
    # identity (sparse_matrix.rs:319-327): I * v = v
    ident = [[(const(M, 1), i)] for i in range(3)]
    assert O.sparse_matvec(M, 3, ident, v) == v


@pytest.mark.parametrize("name", ALL)
def test_c_equals_python_on_random_ring_values(name):
    M = O.MODELS[name]
    rng = random.Random(31)
    rnd = lambda: [rng.randrange(M.p) for _ in range(M.D)]
    nrows, ncols = 6, 5
    coeffs = []
    for i in range(nrows):
        cols = sorted(rng.sample(range(ncols), rng.randrange(0, ncols + 1)))
        coeffs.append([(rnd(), j) for j in cols])
    coeffs[1] = []                                             # empty row -> ZERO
    coeffs[2] = [(rnd(), 4), (rnd(), 4), (rnd(), 0)]           # repeated / unsorted columns are legal
    v = [rnd() for _ in range(ncols)]
    v[0] = [M.p - 1] * M.D
    want = O.sparse_matvec(M, ncols, coeffs, v)
    assert want[1] == [0] * M.D
    rp, ci, vals = csr(M, coeffs)
    assert np.array_equal(C.sparse_matvec(name, nrows, ncols, rp, ci, vals, flat(M, v)), flat(M, want))
    # sparse == its dense image through the dense mat-vec (to_dense, sparse_matrix.rs:108-117, keeps the last
    # entry of a repeated column, so only for rows without repeats)
    dense = [[[0] * M.D for _ in range(ncols)] for _ in range(nrows)]
    for i, row in enumerate(coeffs):
        if i != 2:
            for e, j in row:
                dense[i][j] = e
    dv = O.matvec(M, dense, v)
    assert all(dv[i] == want[i] for i in range(nrows) if i != 2)
    with pytest.raises(IndexError):
        C.sparse_matvec(name, 1, 2, np.array([0, 1], dtype=np.uint64), np.array([2], dtype=np.uint64),
                        raw(M, rnd()), flat(M, v[:2]))
    # mat-mat and scaling
    a = [[rnd() for _ in range(3)] for _ in range(2)]
    m = [[rnd() for _ in range(4)] for _ in range(3)]
    want = O.matmat(M, a, m)
    got = C.matmat(name, [flat(M, r) for r in a], [flat(M, r) for r in m])
    assert all(np.array_equal(g, flat(M, r)) for g, r in zip(got, want))
    # (A m) v == A (m v)
    v4 = [rnd() for _ in range(4)]
    assert O.matvec(M, want, v4) == O.matvec(M, a, O.matvec(M, m, v4))
    r = rnd()
    assert np.array_equal(C.scale(name, flat(M, a[0]), raw(M, r)), flat(M, O.scale(M, a[0], r)))


# ---- canonical serialization (ark-serialize 0.4 restated; the reference holds no serialized vector) ------
@pytest.mark.parametrize("name", ALL)
def test_serialization_oracles(name):
    M = O.MODELS[name]
    rng = random.Random(41)
    nb = {"goldilocks": 8, "babybear": 4, "stark_prime": 32}[name]
    assert O.fe_bytes(M) == nb == C.fe_bytes(name)
    elems = [[rng.randrange(M.p) for _ in range(M.D)] for _ in range(5)]
    elems[0] = [0] * M.D
    elems[1] = [M.p - 1] * M.D
    elems[2] = [1] + [0] * (M.D - 1)   # Ring::ONE: the byte 01 followed by zeros
    data = O.serialize(M, elems)
    assert len(data) == 5 * M.D * nb
    assert data[2 * M.D * nb: 3 * M.D * nb] == b"\x01" + b"\x00" * (M.D * nb - 1)
    assert O.deserialize(M, data) == elems
    got = C.serialize(name, flat(M, elems))
    assert got.tobytes() == data
    assert np.array_equal(C.deserialize(name, got), flat(M, elems))
    # an integer equal to the modulus is InvalidData
    bad = bytearray(data)
    bad[:nb] = M.p.to_bytes(nb, "little") if M.p.bit_length() <= 8 * nb else b"\xff" * nb
    with pytest.raises(ValueError):
        O.deserialize(M, bytes(bad))
    with pytest.raises(ValueError):
        C.deserialize(name, np.frombuffer(bytes(bad), dtype=np.uint8).copy())


# ---- randomised shapes (hypothesis): C oracle == Python oracle for the SURVEY 8f-3 / 8f-4 rows -----------------------
try:
    from hypothesis import given, settings, strategies as st
    HAVE_HYPOTHESIS = True
except Exception:  # pragma: no cover
    HAVE_HYPOTHESIS = False


if HAVE_HYPOTHESIS:
    @settings(max_examples=12, deadline=None)
    @given(name=st.sampled_from(ALL), nrows=st.integers(1, 6), ncols=st.integers(1, 6), seed=st.integers(0, 2**31 - 1))
    def test_sparse_matvec_random_shapes(name, nrows, ncols, seed):
        M = O.MODELS[name]
        rng = random.Random(seed)
        rnd = lambda: [rng.randrange(M.p) for _ in range(M.D)]
        coeffs = [[(rnd(), rng.randrange(ncols)) for _ in range(rng.randrange(0, 4))] for _ in range(nrows)]
        v = [rnd() for _ in range(ncols)]
        want = O.sparse_matvec(M, ncols, coeffs, v)
        rp, ci, vals = csr(M, coeffs)
        got = C.sparse_matvec(name, nrows, ncols, rp, ci, vals, flat(M, v))
        assert np.array_equal(got, flat(M, want))
        # linearity in the vector: A (v + w) == A v + A w
        w = [rnd() for _ in range(ncols)]
        vw = [O.ntt_add(M, a, b) for a, b in zip(v, w)]
        lhs = O.sparse_matvec(M, ncols, coeffs, vw)
        rhs = [O.ntt_add(M, a, b) for a, b in zip(want, O.sparse_matvec(M, ncols, coeffs, w))]
        assert lhs == rhs

    @settings(max_examples=12, deadline=None)
    @given(name=st.sampled_from(ALL), n=st.integers(0, 4), seed=st.integers(0, 2**31 - 1))
    def test_serialization_round_trip_random(name, n, seed):
        M = O.MODELS[name]
        rng = random.Random(seed)
        elems = [[rng.randrange(M.p) for _ in range(M.D)] for _ in range(n)]
        data = O.serialize(M, elems)
        assert len(data) == n * M.D * O.fe_bytes(M)
        assert O.deserialize(M, data) == elems
        if n:
            got = C.serialize(name, flat(M, elems))
            assert got.tobytes() == data
            assert np.array_equal(C.deserialize(name, got), flat(M, elems))
