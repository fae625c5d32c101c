"""CPU oracle (big-integer Python) for the stark-rings hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (stark_rings_b200/) may import
this module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker.

This is a *restatement* of the reference's algorithms, written from the mathematics
 and pinned against the reference's own known-answer vectors
(tests/golden/*.json, extracted from the reference's inline #[test] literals by
tools/gen_golden.py).  The reference is Rust + un-vendored ark-ff 0.4.2 and cannot be
compiled in this environment (no cargo/rustc); field semantics are those of a prime
field, and the raw-limb layout is ark-ff's MontBackend: N little-endian u64 limbs
holding x * 2^(64 N) mod p.

Reference files followed (paths relative to crates/ring/src/cyclotomic_ring/):
  models/goldilocks/ntt.rs:15-47   root table, KAPPA, EIGHT_INV, FOUR_INV
  models/goldilocks/ntt.rs:135-228 CRT schedule        :240-319 ICRT schedule
  models/goldilocks/ntt.rs:326-437 slot isomorphisms (homogenize / dehomogenize)
  models/babybear/ntt.rs:16-41,137-141  tables         :143-236 CRT   :238-317 ICRT
  models/babybear/ntt.rs:324-588   slot isomorphisms + (1 3)(2 6)(5 7) permutation
  models/stark_prime/ntt.rs:16-55  tables   :121-235 CRT   :245-346 ICRT
  ntt_form.rs:159-189,521-550      slot-wise Mul / MulUnchecked
  ntt_form.rs:588-601,640-654      Add / Sum
  coeff_form.rs:54-67              schoolbook poly_mul + reduce_in_place
  models/*/mod.rs reduce_in_place  (goldilocks:75-98, babybear:87-110, stark_prime:40-47)
  crt.rs:6-77                      CRT/ICRT traits, elementwise_* batch forms
  ../../linear_algebra/src/matrix.rs:168-183  checked_mul_vec / try_mul_vec
"""
from __future__ import annotations

SLOT_K = [1, 13, 7, 19, 5, 17, 11, 23]  # goldilocks/ntt.rs:49-58, babybear/ntt.rs:43-52


def _bfly(c, lo, span, w, p):
    """(a, b) <- (a + w b, a - w b) on c[lo+i], c[lo+span+i], i < span."""
    for i in range(span):
        a, b = c[lo + i], c[lo + span + i]
        t = w * b % p
        c[lo + i] = (a + t) % p
        c[lo + span + i] = (a - t) % p


def _ibfly(c, lo, span, w, p):
    """(a, b) <- (a + b, w (a - b))."""
    for i in range(span):
        a, b = c[lo + i], c[lo + span + i]
        c[lo + i] = (a + b) % p
        c[lo + span + i] = w * (a - b) % p


class Phi3Model:
    """Rings F_p[X]/(X^D - X^(D/2) + 1) with 8 slots of width d = D/8 (Goldilocks, BabyBear)."""

    def __init__(self, name, p, D, r, kappa, eight_inv, four_inv, homog, transpose3):
        self.name, self.p, self.D, self.d = name, p, D, D // 8
        self.limbs = 1
        self.R = 1 << 64
        self.W = [pow(r, i, p) for i in range(24)]
        self.r = r
        self.KAPPA, self.EIGHT_INV, self.FOUR_INV = kappa, eight_inv, four_inv
        self.homog = homog  # k -> (e_k, b_k): X |-> r^b_k * Y^e_k
        self.transpose3 = transpose3  # BabyBear: Fq9 stored as 3 Fq3's
        assert pow(r, 24, p) == 1 and pow(r, 12, p) == p - 1
        assert kappa * (2 * self.W[4] - 1) % p == 1
        assert eight_inv * 8 % p == 1 and four_inv * 4 % p == 1
        for k, (e, b) in homog.items():
            assert pow(r, self.d * b + e, p) == self.W[k], (name, k)
        # memory index of the coefficient of Y^j inside a slot
        if transpose3:
            self.mem_of_pow = [3 * (j % 3) + j // 3 for j in range(self.d)]
        else:
            self.mem_of_pow = list(range(self.d))

    # -- slot isomorphisms -------------------------------------------------------------
    def _slot_map(self, k):
        """Returns list of (dst_pow, scale) for source power j (monomial map X -> r^b Y^e)."""
        d, p = self.d, self.p
        if k == 1:
            return [(j, 1) for j in range(d)]
        e, b = self.homog[k]
        return [((j * e) % d, pow(self.r, b * j + (j * e) // d, p)) for j in range(d)]

    def homogenize(self, c):
        """goldilocks/ntt.rs:326-334, babybear/ntt.rs:324-333."""
        d, p = self.d, self.p
        out = list(c)
        for s, k in enumerate(SLOT_K):
            src = c[s * d:(s + 1) * d]
            dst = [0] * d
            for j, (dj, sc) in enumerate(self._slot_map(k)):
                dst[self.mem_of_pow[dj]] = src[j] * sc % p
            out[s * d:(s + 1) * d] = dst
        return out

    def dehomogenize(self, c):
        """goldilocks/ntt.rs:338-346, babybear/ntt.rs:337-346."""
        d, p = self.d, self.p
        out = list(c)
        for s, k in enumerate(SLOT_K):
            src = c[s * d:(s + 1) * d]
            dst = [0] * d
            for j, (dj, sc) in enumerate(self._slot_map(k)):
                dst[j] = src[self.mem_of_pow[dj]] * pow(sc, p - 2, p) % p
            out[s * d:(s + 1) * d] = dst
        return out

    # -- CRT / ICRT --------------------------------------------------------------------
    def crt_stages(self, c):
        """Butterfly stages only (no homogenize): goldilocks/ntt.rs:146-225, babybear/ntt.rs:154-233."""
        D, p, W = self.D, self.p, self.W
        c = [x % p for x in c]
        assert len(c) == D
        h = D // 2
        for i in range(h):
            a, b = c[i], c[h + i]
            z = W[4] * b % p
            c[i] = (a + z) % p
            c[h + i] = (a + b - z) % p
        q = D // 4
        _bfly(c, 0, q, W[2], p)
        _bfly(c, h, q, W[10], p)
        e = D // 8
        _bfly(c, 0, e, W[1], p)
        _bfly(c, q, e, W[7], p)
        _bfly(c, h, e, W[5], p)
        _bfly(c, 3 * q, e, W[11], p)
        return c

    def icrt_stages(self, c):
        """Inverse butterfly stages only: goldilocks/ntt.rs:250-318, babybear/ntt.rs:249-316."""
        D, p, W = self.D, self.p, self.W
        c = [x % p for x in c]
        h, q, e = D // 2, D // 4, D // 8
        _ibfly(c, 0, e, W[23], p)
        _ibfly(c, q, e, W[17], p)
        _ibfly(c, h, e, W[19], p)
        _ibfly(c, 3 * q, e, W[13], p)
        _ibfly(c, 0, q, W[22], p)
        _ibfly(c, h, q, W[14], p)
        for i in range(h):
            a, b = c[i], c[h + i]
            kd = self.KAPPA * (a - b) % p
            c[i] = self.EIGHT_INV * (a + b - kd) % p
            c[h + i] = self.FOUR_INV * kd % p
        return c

    def crt(self, c):
        return self.homogenize(self.crt_stages(c))

    def icrt(self, c):
        assert len(c) == self.D
        return self.icrt_stages(self.dehomogenize(c))

    # -- slot arithmetic ----------------------------------------------------------------
    def slot_mul(self, x, y):
        """Product in F_p[Y]/(Y^d - r) on memory-ordered coefficients (ark-ff CubicExtField
        semantics; Fq9 = Fq3[Y]/(Y^3-u), babybear/fq9.rs:19-58, goldilocks/mod.rs:34-52)."""
        d, p = self.d, self.p
        m = self.mem_of_pow
        xs = [x[m[j]] for j in range(d)]
        ys = [y[m[j]] for j in range(d)]
        acc = [0] * (2 * d - 1)
        for i in range(d):
            for j in range(d):
                acc[i + j] += xs[i] * ys[j]
        out = [0] * d
        for j in range(d):
            v = acc[j] + (self.r * acc[j + d] if j + d < 2 * d - 1 else 0)
            out[m[j]] = v % p
        return out

    def ntt_mul(self, a, b):
        """ntt_form.rs:159-175 (zero short-circuit is a semantic no-op)."""
        d = self.d
        out = []
        for s in range(8):
            out += self.slot_mul(a[s * d:(s + 1) * d], b[s * d:(s + 1) * d])
        return out

    def reduce(self, c):
        """X^D = X^(D/2) - 1 (goldilocks/mod.rs:75-98, babybear/mod.rs:87-110)."""
        D, p = self.D, self.p
        c = list(c) + [0] * (2 * D - len(c))
        h = D // 2
        out = [0] * D
        for i in range(h):
            out[i] = (c[i] - c[D + i] - c[D + h + i]) % p
        for i in range(h, D):
            out[i] = (c[i] + c[h + i]) % p
        return out


class StarkModel:
    """F_p[X]/(X^16+1), p = 2^251 + 17 2^192 + 1; fully splitting (stark_prime/ntt.rs)."""

    def __init__(self):
        self.name = "stark_prime"
        self.p = p = 3618502788666131213697322783095070105623107215331596699973092056135872020481
        self.D, self.d, self.limbs = 16, 1, 4
        self.R = 1 << 256
        w1 = 3409443867035641044245057348756544640549407421541289951053907001322227935403
        self.W = [pow(w1, i, p) for i in range(32)]
        assert pow(w1, 16, p) == p - 1
        self.SIXTEEN_INV = pow(16, p - 2, p)
        self.SIXTEEN_INV_W24 = self.SIXTEEN_INV * self.W[24] % p
        self.eval_order = [1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31]

    def crt(self, c):
        p, W = self.p, self.W
        c = [x % p for x in c]
        assert len(c) == 16
        _bfly(c, 0, 8, W[8], p)
        for lo, k in ((0, 4), (8, 12)):
            _bfly(c, lo, 4, W[k], p)
        for lo, k in ((0, 2), (4, 10), (8, 6), (12, 14)):
            _bfly(c, lo, 2, W[k], p)
        for lo, k in zip(range(0, 16, 2), (1, 9, 5, 13, 3, 11, 7, 15)):
            _bfly(c, lo, 1, W[k], p)
        return c

    def icrt(self, c):
        p, W = self.p, self.W
        c = [x % p for x in c]
        assert len(c) == 16
        for lo, k in zip(range(0, 16, 2), (31, 23, 27, 19, 29, 21, 25, 17)):
            _ibfly(c, lo, 1, W[k], p)
        for lo, k in ((0, 30), (4, 22), (8, 26), (12, 18)):
            _ibfly(c, lo, 2, W[k], p)
        for lo, k in ((0, 28), (8, 20)):
            _ibfly(c, lo, 4, W[k], p)
        for i in range(8):
            a, b = c[i], c[8 + i]
            c[i] = self.SIXTEEN_INV * (a + b) % p
            c[8 + i] = self.SIXTEEN_INV_W24 * (a - b) % p
        return c

    def crt_stages(self, c):
        return self.crt(c)

    def icrt_stages(self, c):
        return self.icrt(c)

    def homogenize(self, c):
        return list(c)

    def dehomogenize(self, c):
        return list(c)

    def ntt_mul(self, a, b):
        return [x * y % self.p for x, y in zip(a, b)]

    def reduce(self, c):
        """X^16 = -1 (stark_prime/mod.rs:40-47)."""
        c = list(c) + [0] * (32 - len(c))
        return [(c[i] - c[16 + i]) % self.p for i in range(16)]


GOLDILOCKS = Phi3Model(
    "goldilocks", 18446744069414584321, 24, 1099511627776,
    12297829382473034411, 16140901060737761281, 13835058052060938241,
    {13: (1, 12), 7: (1, 2), 19: (1, 6), 5: (2, 1), 17: (2, 5), 11: (2, 3), 23: (2, 7)},
    transpose3=False)

BABYBEAR = Phi3Model(
    "babybear", 2013265921, 72, 503591070,
    1807872479, 1761607681, 1509949441,
    {13: (4, 1), 7: (7, 0), 19: (1, 2), 5: (5, 0), 17: (8, 1), 11: (2, 1), 23: (5, 2)},
    transpose3=True)

STARK = StarkModel()

MODELS = {"goldilocks": GOLDILOCKS, "babybear": BABYBEAR, "stark_prime": STARK,
          "gl": GOLDILOCKS, "bb": BABYBEAR, "sp": STARK}


# -- generic helpers -------------------------------------------------------------------
def poly_mul(M, a, b):
    """coeff_form.rs:54-67: schoolbook then reduce mod Phi."""
    D, p = M.D, M.p
    acc = [0] * (2 * D)
    for i in range(D):
        for j in range(D):
            acc[i + j] += a[i] * b[j]
    return M.reduce([x % p for x in acc])


def ring_mul(M, a, b):
    """icrt(crt(a) * crt(b)) -- the fused unit of the metric."""
    return M.icrt(M.ntt_mul(M.crt(a), M.crt(b)))


def rot(M, c):
    """Cyclotomic::rot, multiplication by X (goldilocks/mod.rs:138-149, babybear/mod.rs:150-161,
    stark_prime/mod.rs:87-95)."""
    D, p = M.D, M.p
    last = c[D - 1]
    out = [(-last) % p] + [x % p for x in c[:D - 1]]
    if M.name != "stark_prime":
        out[D // 2] = (out[D // 2] + last) % p
    return out


def decompose_balanced(M, x, b, pad):
    """decompose_balanced_in_place (balanced_decomposition/mod.rs:62-103) of one field element (standard form)
    into `pad` digits in [-b/2, b/2], returned as field elements; raises IndexError like the reference's
    out-of-bounds panic when `pad` is too small."""
    assert b >= 2 and b % 2 == 0
    p = M.p
    cur = x - p if x > (p - 1) // 2 else x           # fq_convertible.rs:22-35
    out = [0] * pad
    i = 0
    while True:
        q = abs(cur) // b * (1 if cur >= 0 else -1)  # Rust: truncating division / remainder
        rem = cur - q * b
        if abs(rem) <= b // 2:
            digit, cur = rem, q
        else:
            digit = rem + b if rem < 0 else rem - b
            cur = q + (-1 if rem < 0 else 1)         # rounded_div(rem, b)
        out[i] = digit % p                           # IndexError == the reference's panic
        i += 1
        if cur == 0:
            break
    return out


def gadget_decompose(M, elems, b, pad):
    """GadgetDecompose for &[RqPoly] (mod.rs:163-175; per element coeff_form.rs:588-606)."""
    out = []
    for e in elems:
        digits = [decompose_balanced(M, c, b, pad) for c in e]
        for t in range(pad):
            out.append([digits[i][t] for i in range(M.D)])
    return out


def gadget_recompose(M, digit_elems, b, pad):
    """GadgetRecompose for &[R] (mod.rs:177-190): Horner in the ring."""
    out = []
    for j in range(0, len(digit_elems), pad):
        acc = [0] * M.D
        for d in reversed(digit_elems[j:j + pad]):
            acc = [(a * b + x) % M.p for a, x in zip(acc, d)]
        out.append(acc)
    return out


def ntt_add(M, a, b):
    return [(x + y) % M.p for x, y in zip(a, b)]


def matvec(M, rows, v):
    """matrix.rs:168-178 with R = RqNTT: y_i = sum_j A[i][j] * v[j]; None on length mismatch."""
    out = []
    for row in rows:
        if len(row) != len(v):
            return None
        acc = [0] * M.D
        for a, x in zip(row, v):
            acc = ntt_add(M, acc, M.ntt_mul(a, x))
        out.append(acc)
    return out


def sparse_matvec(M, ncols, coeffs, v):
    """sparse_matrix.rs:201-212 with R = RqNTT: coeffs[i] = [(element, column), ...];
    out[i] = sum of element * v[column], Sum folded from ZERO; None when ncols != len(v)."""
    if ncols != len(v):
        return None
    out = []
    for row in coeffs:
        acc = [0] * M.D
        for r, j in row:
            acc = ntt_add(M, acc, M.ntt_mul(r, v[j]))  # v[j] out of range: IndexError, the reference panics
        out.append(acc)
    return out


def matmat(M, a, m):
    """matrix.rs:148-166 with R = RqNTT: out[i][j] = sum_k a[i][k] * m[k][j]; None when a.ncols != m.nrows
    (ncols of a matrix = length of its first row, 0 when it has no rows: matrix.rs:100-108)."""
    a_ncols = len(a[0]) if a else 0
    m_ncols = len(m[0]) if m else 0
    if a_ncols != len(m):
        return None
    out = []
    for row in a:
        orow = []
        for j in range(m_ncols):
            acc = [0] * M.D
            for k in range(a_ncols):
                acc = ntt_add(M, acc, M.ntt_mul(row[k], m[k][j]))
            orow.append(acc)
        out.append(orow)
    return out


def scale(M, elems, r):
    """MulAssign<&R> on every entry (matrix.rs:207-211, sparse_matrix.rs:298-302)."""
    return [M.ntt_mul(x, r) for x in elems]


def fe_bytes(M):
    """ark-serialize 0.4 (restated; not vendored in the reference tree): ceil(MODULUS_BIT_SIZE / 8) bytes."""
    return (M.p.bit_length() + 7) // 8


def serialize(M, elems):
    """CanonicalSerialize of ring elements given as lists of standard-form ints in memory order
    (coeff_form.rs:154-167 / ntt_form.rs:24): little-endian bytes of each field element, no length prefix."""
    nb = fe_bytes(M)
    return b"".join(int(x).to_bytes(nb, "little") for e in elems for x in e)


def deserialize(M, data):
    """CanonicalDeserialize (coeff_form.rs:178-189): ValueError where ark-serialize returns InvalidData."""
    nb = fe_bytes(M)
    if len(data) % (nb * M.D):
        raise ValueError("not a whole number of ring elements")
    vals = [int.from_bytes(data[i:i + nb], "little") for i in range(0, len(data), nb)]
    if any(v >= M.p for v in vals):
        raise ValueError("InvalidData: integer not below the modulus")
    return [vals[i:i + M.D] for i in range(0, len(vals), M.D)]


def to_raw(M, vals):
    """standard-form ints -> flat list of little-endian u64 limbs of x*R mod p (ark-ff MontBackend)."""
    out = []
    for x in vals:
        m = x % M.p * M.R % M.p
        for _ in range(M.limbs):
            out.append(m & 0xFFFFFFFFFFFFFFFF)
            m >>= 64
    return out


def from_raw(M, limbs):
    rinv = pow(M.R, M.p - 2, M.p)
    out = []
    for i in range(0, len(limbs), M.limbs):
        m = 0
        for j in reversed(range(M.limbs)):
            m = (m << 64) | int(limbs[i + j])
        assert m < M.p, "non-canonical limb value"
        out.append(m * rinv % M.p)
    return out
