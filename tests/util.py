"""Seeded synthetic inputs shared by tests, smoke() and bench.py: canonical residues as raw limbs."""
import numpy as np

P = {
    "goldilocks": 18446744069414584321,
    "babybear": 2013265921,
    "stark_prime": 3618502788666131213697322783095070105623107215331596699973092056135872020481,
}
WORDS = {"goldilocks": 24, "babybear": 72, "stark_prime": 64}


def rand_raw(name, n, seed, edge=True):
    """n ring elements of `name` as a flat uint64 array of raw (Montgomery) limbs, every field
    element canonical (< p).  Raw limbs of uniform residues are uniform residues, so sampling the
    limbs directly is the same distribution as sampling values and converting."""
    rng = np.random.default_rng(seed)
    w = WORDS[name]
    if name == "stark_prime":
        a = rng.integers(0, 1 << 63, size=n * w, dtype=np.uint64) * np.uint64(2) + \
            rng.integers(0, 2, size=n * w, dtype=np.uint64)
        a[3::4] &= np.uint64((1 << 59) - 1)  # < 2^251 < p
        if edge and n >= 2:
            pm1 = P[name] - 1
            limbs = [(pm1 >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
            a[:w] = 0
            a[w:2 * w] = np.tile(np.array(limbs, dtype=np.uint64), 16)
    else:
        a = rng.integers(0, P[name], size=n * w, dtype=np.uint64)
        if edge and n >= 2:
            a[:w] = 0
            a[w:2 * w] = np.uint64(P[name] - 1)
    return a
