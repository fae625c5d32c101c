"""ctypes loader for the C restatement (oracle/sr_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RINGS = {"goldilocks": 0, "babybear": 1, "stark_prime": 2, "gl": 0, "bb": 1, "sp": 2}
_lib = None


def _bind(path):
    L = ctypes.CDLL(path)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    L.sro_elem_words.restype = ctypes.c_size_t
    L.sro_elem_words.argtypes = [ctypes.c_int]
    for name in ("sro_crt", "sro_icrt"):
        getattr(L, name).argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_int]
        getattr(L, name).restype = None
    L.sro_ntt_mul.argtypes = [ctypes.c_int, u64p, u64p, ctypes.c_size_t, ctypes.c_int]
    L.sro_ntt_mul.restype = None
    L.sro_ring_mul.argtypes = [ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t, ctypes.c_int]
    L.sro_ring_mul.restype = None
    L.sro_matvec.argtypes = [ctypes.c_int, ctypes.POINTER(u64p), ctypes.c_size_t, ctypes.c_size_t, u64p,
                             ctypes.c_size_t, u64p, ctypes.c_int]
    L.sro_matvec.restype = ctypes.c_int
    L.sro_reduce.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_size_t, u64p]
    L.sro_reduce.restype = None
    L.sro_rot.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p]
    L.sro_rot.restype = None
    for name in ("sro_gadget_decompose", "sro_gadget_recompose"):
        getattr(L, name).argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_size_t, u64p]
        getattr(L, name).restype = ctypes.c_int
    L.sro_sparse_matvec.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, u64p, u64p, u64p, u64p,
                                    ctypes.c_size_t, u64p]
    L.sro_sparse_matvec.restype = ctypes.c_int
    L.sro_matmat.argtypes = [ctypes.c_int, ctypes.POINTER(u64p), ctypes.c_size_t, ctypes.c_size_t,
                             ctypes.POINTER(u64p), ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(u64p)]
    L.sro_matmat.restype = ctypes.c_int
    L.sro_addsub.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, ctypes.c_size_t]
    L.sro_addsub.restype = None
    L.sro_sum.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p]
    L.sro_sum.restype = None
    L.sro_scale.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p]
    L.sro_scale.restype = None
    u8p = ctypes.POINTER(ctypes.c_uint8)
    L.sro_fe_bytes.argtypes = [ctypes.c_int]
    L.sro_fe_bytes.restype = ctypes.c_size_t
    L.sro_serialize.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u8p]
    L.sro_serialize.restype = None
    L.sro_deserialize.argtypes = [ctypes.c_int, u8p, ctypes.c_size_t, u64p]
    L.sro_deserialize.restype = ctypes.c_int
    L.sro_crt_stages.argtypes = [ctypes.c_int, u64p]
    L.sro_crt_stages.restype = None
    return L


def lib():
    """Load oracle/libsr_oracle.so (portable x86-64-v3 build), building it if missing."""
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libsr_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", HERE, "libsr_oracle.so"], check=True)
        _lib = _bind(path)
    return _lib


def lib_native():
    """Rebuild with -march=native on THIS machine (the timed CPU baseline on the GPU box's own
    host cores); falls back to the portable build if no compiler is available."""
    npath = os.path.join(HERE, "libsr_oracle_native.so")
    try:
        subprocess.run(["gcc", "-O3", "-march=native", "-fPIC", "-std=gnu11", "-shared", "-o", npath,
                        os.path.join(HERE, "sr_oracle.c"), "-lpthread"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return _bind(npath), "native"
    except Exception:
        return lib(), "x86-64-v3"


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def words(ring):
    return {0: 24, 1: 72, 2: 64}[RINGS[ring]]


def crt(ring, buf, threads=1, L=None):
    """In place on a uint64 array of n*words limbs; returns it."""
    (L or lib()).sro_crt(RINGS[ring], _p(buf), buf.size // words(ring), threads)
    return buf


def icrt(ring, buf, threads=1, L=None):
    (L or lib()).sro_icrt(RINGS[ring], _p(buf), buf.size // words(ring), threads)
    return buf


def ntt_mul(ring, a, b, threads=1, L=None):
    (L or lib()).sro_ntt_mul(RINGS[ring], _p(a), _p(b), a.size // words(ring), threads)
    return a


def ring_mul(ring, a, b, threads=1, L=None):
    out = np.empty_like(a)
    (L or lib()).sro_ring_mul(RINGS[ring], _p(a), _p(b), _p(out), a.size // words(ring), threads)
    return out


def addsub(ring, op, a, b=None, L=None):
    """op 'add' / 'sub' / 'neg' element-wise over the batch; returns a new array."""
    out = a.copy()
    code = {"add": 0, "sub": 1, "neg": 2}[op]
    (L or lib()).sro_addsub(RINGS[ring], code, _p(out), _p(b if b is not None else out), out.size // words(ring))
    return out


def ring_sum(ring, a, L=None):
    out = np.zeros(words(ring), dtype=np.uint64)
    (L or lib()).sro_sum(RINGS[ring], _p(a), a.size // words(ring), _p(out))
    return out


def matvec(ring, rows, v, threads=1, L=None):
    """rows: list of uint64 arrays (m*words each); v: uint64 array.  Returns kappa*words array or None."""
    w = words(ring)
    kappa = len(rows)
    m = rows[0].size // w if kappa else 0
    u64p = ctypes.POINTER(ctypes.c_uint64)
    arr = (u64p * max(kappa, 1))(*[_p(r) for r in rows])
    out = np.zeros(kappa * w, dtype=np.uint64)
    rc = (L or lib()).sro_matvec(RINGS[ring], arr, kappa, m, _p(v), v.size // w, _p(out), threads)
    return None if rc else out


def sparse_matvec(ring, nrows, ncols, row_ptr, col_idx, vals, v, L=None):
    """CSR arrays (uint64) -> nrows*words array; None on length mismatch; IndexError on a bad column index."""
    w = words(ring)
    out = np.zeros(nrows * w, dtype=np.uint64)
    pad = np.zeros(1, dtype=np.uint64)
    rc = (L or lib()).sro_sparse_matvec(RINGS[ring], nrows, ncols, _p(row_ptr), _p(col_idx if col_idx.size else pad),
                                        _p(vals if vals.size else pad), _p(v if v.size else pad), v.size // w, _p(out))
    if rc == 2:
        raise IndexError("column index out of range")
    return None if rc else out


def matmat(ring, a_rows, m_rows, L=None):
    """a_rows / m_rows: lists of uint64 row arrays -> list of output rows, or None on a shape mismatch."""
    w = words(ring)
    a_ncols = a_rows[0].size // w if a_rows else 0
    m_ncols = m_rows[0].size // w if m_rows else 0
    u64p = ctypes.POINTER(ctypes.c_uint64)
    outs = [np.zeros(m_ncols * w, dtype=np.uint64) for _ in a_rows]
    pad = np.zeros(1, dtype=np.uint64)
    tab = lambda rows: (u64p * max(len(rows), 1))(*[_p(r if r.size else pad) for r in rows])
    rc = (L or lib()).sro_matmat(RINGS[ring], tab(a_rows), len(a_rows), a_ncols, tab(m_rows), len(m_rows), m_ncols,
                                 tab(outs))
    return None if rc else outs


def scale(ring, a, r, L=None):
    (L or lib()).sro_scale(RINGS[ring], _p(a), a.size // words(ring), _p(r))
    return a


def fe_bytes(ring, L=None):
    return (L or lib()).sro_fe_bytes(RINGS[ring])


def serialize(ring, a, L=None):
    """n elements (raw limbs) -> uint8 array of their canonical serialization (no length prefix)."""
    L = L or lib()
    n = a.size // words(ring)
    nfe = n * (16 if RINGS[ring] == 2 else words(ring))
    out = np.zeros(nfe * L.sro_fe_bytes(RINGS[ring]), dtype=np.uint8)
    L.sro_serialize(RINGS[ring], _p(a), n, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


def deserialize(ring, data, L=None):
    """uint8 array -> raw limbs; ValueError on an integer >= p (InvalidData)."""
    L = L or lib()
    D = 16 if RINGS[ring] == 2 else words(ring)
    n = data.size // (D * L.sro_fe_bytes(RINGS[ring]))
    out = np.zeros(n * words(ring), dtype=np.uint64)
    if L.sro_deserialize(RINGS[ring], data.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), n, _p(out)):
        raise ValueError("InvalidData")
    return out


def reduce(ring, polys, coeffs_per_poly, L=None):
    """polys: flat uint64 array of n polynomials with coeffs_per_poly field elements each -> n ring elements."""
    w = words(ring)
    nlimb = 4 if RINGS[ring] == 2 else 1
    n = polys.size // (coeffs_per_poly * nlimb)
    out = np.empty(n * w, dtype=np.uint64)
    (L or lib()).sro_reduce(RINGS[ring], _p(polys), n, coeffs_per_poly, _p(out))
    return out


def rot(ring, a, L=None):
    out = np.empty_like(a)
    (L or lib()).sro_rot(RINGS[ring], _p(a), a.size // words(ring), _p(out))
    return out


def gadget_decompose(ring, a, b, pad, L=None):
    """n elements -> n * pad digit elements; raises IndexError when pad is too small (the reference panics)."""
    n = a.size // words(ring)
    out = np.empty(n * pad * words(ring), dtype=np.uint64)
    rc = (L or lib()).sro_gadget_decompose(RINGS[ring], _p(a), n, b, pad, _p(out))
    if rc == 1:
        raise IndexError("padding_size too small")
    if rc:
        raise ValueError("unsupported ring or basis")
    return out


def gadget_recompose(ring, digits, b, pad, L=None):
    n = digits.size // (words(ring) * pad)
    out = np.empty(n * words(ring), dtype=np.uint64)
    rc = (L or lib()).sro_gadget_recompose(RINGS[ring], _p(digits), n, b, pad, _p(out))
    if rc:
        raise ValueError("unsupported ring")
    return out
