"""Mirror of stark-rings-linalg's dense Matrix for R = RqNTT (linear_algebra/src/matrix.rs).

  Matrix { nrows, ncols, vals: Vec<Vec<R>> }   matrix.rs:17-21  -> Matrix(rows)
  checked_mul_vec                              matrix.rs:168-178 -> Matrix.checked_mul_vec (None on mismatch)
  try_mul_vec                                  matrix.rs:180-183 -> Matrix.try_mul_vec (raises DifferentLengths)
  Mul<&[R]> for &Matrix<R>                     matrix.rs:199-205 -> Matrix.__matmul__ / __mul__
  checked_mul_mat / try_mul_mat                matrix.rs:148-166,185-188 -> Matrix.checked_mul_mat / try_mul_mat
  MulAssign<&R> for Matrix<R>                  matrix.rs:207-211 -> Matrix.__imul__
and of SparseMatrix for R = RqNTT (linear_algebra/src/sparse_matrix.rs), SURVEY.md 8f-3:
  SparseMatrix { nrows, ncols, coeffs }        sparse_matrix.rs:17-21   -> SparseMatrix (CSR image of coeffs)
  identity / from_dense / to_dense             sparse_matrix.rs:86-97,108-137
  checked_mul_vec / try_mul_vec / Mul<&[R]>    sparse_matrix.rs:201-217,278-286
  MulAssign<&R>                                sparse_matrix.rs:298-302
The ring mat-vec is the hot path; the rest are the callers' neighbouring linear maps.  Padding, hconcat, rand,
serialization and the sparse x sparse product (whose output structure depends on which products are zero,
sparse_matrix.rs:219-275) stay with the caller.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib as L
from .errors import DifferentLengths
from .rings import RqNTT, _ptr_loc, default_context

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class Matrix:
    def __init__(self, rows, ctx=None):
        """rows: list of RqNTT batches (each row is its own allocation, as Vec<Vec<R>>)."""
        self.vals = list(rows)
        self.nrows = len(self.vals)
        self.ncols = len(self.vals[0]) if self.vals else 0
        self.config = self.vals[0].config if self.vals else None
        for r in self.vals:
            if not isinstance(r, RqNTT) or r.config is not self.config or len(r) != self.ncols:
                raise ValueError("rows must be RqNTT batches of one ring and equal length")
        self.ctx = ctx

    def _call(self, fn, v: RqNTT, partial=False):
        cfg = self.config
        if v.config is not cfg:
            raise TypeError("vector is over a different ring")
        pv, nv, loc, dev = _ptr_loc(v.data)
        ptrs = (ctypes.c_void_p * max(self.nrows, 1))()
        for i, r in enumerate(self.vals):
            pr, _, locr, _ = _ptr_loc(r.data)
            if locr != loc:
                raise ValueError("matrix rows and vector must live in the same place")
            ptrs[i] = pr
        if loc == L.SR_DEVICE:
            out = torch.empty(self.nrows * cfg.limbs, dtype=v.data.dtype, device=v.data.device)
        else:
            out = np.empty(self.nrows * cfg.limbs, dtype=np.uint64)
        po = _ptr_loc(out)[0]
        c = self.ctx or v.ctx or default_context(0 if dev is None else dev)
        if dev is not None:
            c.use_torch_stream()
        rc = fn(c.h, cfg.ring_id, ptrs, self.nrows, self.ncols, pv, nv, po, loc)
        if rc == L.SR_ERR_BAD_LENGTH:
            return None
        c.check(rc, "sr_matvec")
        return RqNTT(cfg, out, c)

    def checked_mul_vec(self, v: RqNTT):
        """matrix.rs:168-178: None when ncols != v.len()."""
        if self.nrows == 0:
            return RqNTT(v.config, np.empty(0, dtype=np.uint64)) if self.ncols == len(v) else None
        return self._call(L.lib.sr_matvec, v)

    def try_mul_vec(self, v: RqNTT) -> RqNTT:
        """matrix.rs:180-183: Err(AlgebraError::DifferentLengths(ncols, v.len()))."""
        out = self.checked_mul_vec(v)
        if out is None:
            raise DifferentLengths(self.ncols, len(v))
        return out

    def __matmul__(self, v: RqNTT) -> RqNTT:
        return self.try_mul_vec(v)

    __mul__ = __matmul__

    def checked_mul_mat(self, m: "Matrix"):
        """matrix.rs:148-166: out[i][j] = sum_k self[i][k] * m[k][j]; None when self.ncols != m.nrows."""
        if self.ncols != m.nrows:
            return None
        cfg = self.config or m.config
        if self.nrows == 0 or cfg is None:
            return Matrix([], self.ctx)
        if m.config is not None and m.config is not cfg:
            raise TypeError("matrices are over different rings")
        loc, dev, like = None, None, None
        pa = (ctypes.c_void_p * max(self.nrows, 1))()
        pm = (ctypes.c_void_p * max(m.nrows, 1))()
        for tab, rows in ((pa, self.vals), (pm, m.vals)):
            for i, r in enumerate(rows):
                p, _, locr, devr = _ptr_loc(r.data)
                if loc is None:
                    loc, dev, like = locr, devr, r.data
                if locr != loc:
                    raise ValueError("all rows must live in the same place")
                tab[i] = p
        outs, po = [], (ctypes.c_void_p * self.nrows)()
        for i in range(self.nrows):
            if loc == L.SR_DEVICE:
                o = torch.empty(m.ncols * cfg.limbs, dtype=like.dtype, device=like.device)
            else:
                o = np.empty(m.ncols * cfg.limbs, dtype=np.uint64)
            outs.append(o)
            po[i] = _ptr_loc(o)[0]
        c = self.ctx or default_context(0 if dev is None else dev)
        if dev is not None:
            c.use_torch_stream()
        rc = L.lib.sr_matmat(c.h, cfg.ring_id, pa, self.nrows, self.ncols, pm, m.nrows, m.ncols, po, loc)
        if rc == L.SR_ERR_BAD_LENGTH:
            return None
        c.check(rc, "sr_matmat")
        return Matrix([RqNTT(cfg, o, c) for o in outs], c)

    def try_mul_mat(self, m: "Matrix") -> "Matrix":
        """matrix.rs:185-188: Err(AlgebraError::DifferentLengths(self.ncols, m.nrows))."""
        out = self.checked_mul_mat(m)
        if out is None:
            raise DifferentLengths(self.ncols, m.nrows)
        return out

    def __imul__(self, r: RqNTT) -> "Matrix":
        """matrix.rs:207-211: every entry *= r (r: one NTT-form element)."""
        for row in self.vals:
            _scale(row, r, self.ctx)
        return self

    def partial_mul_vec(self, v: RqNTT) -> RqNTT:
        """One rank's share of a column-sharded commitment (sr_matvec_partial)."""
        out = self._call(L.lib.sr_matvec_partial, v)
        if out is None:
            raise DifferentLengths(self.ncols, len(v))
        return out


def _scale(batch: RqNTT, r: RqNTT, ctx=None):
    cfg = batch.config
    if r.config is not cfg or len(r) != 1:
        raise TypeError("the scalar must be one element of the same ring")
    pa, na, loc, dev = _ptr_loc(batch.data)
    pr, _, locr, _ = _ptr_loc(r.data)
    if locr != loc:
        raise ValueError("batch and scalar must live in the same place")
    c = ctx or batch.ctx or default_context(0 if dev is None else dev)
    if dev is not None:
        c.use_torch_stream()
    c.check(L.lib.sr_ntt_scale_batch(c.h, cfg.ring_id, pa, na, pr, loc), "sr_ntt_scale_batch")


class SparseMatrix:
    """SparseMatrix<RqNTT> (sparse_matrix.rs:17-21) held as the CSR image of `coeffs: Vec<Vec<(R, usize)>>`:
    row_ptr (nrows + 1 entry offsets), col_idx (nnz column indices) and vals (an RqNTT batch of nnz elements in row
    order).  row_ptr / col_idx are numpy.uint64 arrays or CUDA tensors living where vals lives."""

    def __init__(self, nrows, ncols, row_ptr, col_idx, vals: RqNTT, ctx=None):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.row_ptr, self.col_idx, self.vals = row_ptr, col_idx, vals
        self.config = vals.config
        self.ctx = ctx

    @classmethod
    def from_coeffs(cls, config, nrows, ncols, coeffs, device=None, ctx=None):
        """coeffs: per row a list of (element limbs as a numpy.uint64 array of config.limbs, column index)."""
        row_ptr = np.zeros(nrows + 1, dtype=np.uint64)
        cols, limbs = [], []
        for i, row in enumerate(coeffs):
            for val, j in row:
                limbs.append(np.asarray(val, dtype=np.uint64).reshape(config.limbs))
                cols.append(j)
            row_ptr[i + 1] = len(cols)
        row_ptr[len(coeffs) + 1:] = len(cols)  # pad_rows: trailing empty rows
        col_idx = np.asarray(cols, dtype=np.uint64)
        flat = np.concatenate(limbs) if limbs else np.empty(0, dtype=np.uint64)
        if device is not None:
            to = lambda a: torch.from_numpy(a.view(np.int64)).to(device)
            return cls(nrows, ncols, to(row_ptr), to(col_idx), RqNTT(config, to(flat), ctx), ctx)
        return cls(nrows, ncols, row_ptr, col_idx, RqNTT(config, flat, ctx), ctx)

    @classmethod
    def identity(cls, config, n, one_limbs, device=None, ctx=None):
        """sparse_matrix.rs:86-97; one_limbs = the limbs of R::one() in NTT form."""
        return cls.from_coeffs(config, n, n, [[(one_limbs, i)] for i in range(n)], device, ctx)

    def to_coeffs(self):
        """Back to Vec<Vec<(R, usize)>> on the host: per row a list of (limbs, column)."""
        host = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy().view(np.uint64)
        rp, ci, vals = host(self.row_ptr), host(self.col_idx), host(self.vals.data)
        w = self.config.limbs
        return [[(vals[e * w:(e + 1) * w].copy(), int(ci[e])) for e in range(int(rp[i]), int(rp[i + 1]))]
                for i in range(self.nrows)]

    def checked_mul_vec(self, v: RqNTT):
        """sparse_matrix.rs:201-212: None when ncols != v.len()."""
        cfg = self.config
        if v.config is not cfg:
            raise TypeError("vector is over a different ring")
        if self.ncols != len(v):
            return None
        pv, nv, loc, dev = _ptr_loc(v.data)
        prp, _, l1, _ = _ptr_loc(self.row_ptr)
        pci, _, l2, _ = _ptr_loc(self.col_idx) if len(self.col_idx) else (None, 0, loc, None)
        pvals, _, l3, _ = _ptr_loc(self.vals.data) if len(self.vals) else (None, 0, loc, None)
        if not (l1 == l2 == l3 == loc):
            raise ValueError("matrix arrays and vector must live in the same place")
        if loc == L.SR_DEVICE:
            out = torch.empty(self.nrows * cfg.limbs, dtype=v.data.dtype, device=v.data.device)
        else:
            out = np.empty(self.nrows * cfg.limbs, dtype=np.uint64)
        po = _ptr_loc(out)[0] if self.nrows else None
        c = self.ctx or v.ctx or default_context(0 if dev is None else dev)
        if dev is not None:
            c.use_torch_stream()
        rc = L.lib.sr_sparse_matvec(c.h, cfg.ring_id, self.nrows, self.ncols, prp, pci, pvals, pv, nv, po, loc)
        if rc == L.SR_ERR_BAD_LENGTH:
            return None
        c.check(rc, "sr_sparse_matvec")
        return RqNTT(cfg, out, c)

    def try_mul_vec(self, v: RqNTT) -> RqNTT:
        """sparse_matrix.rs:214-217: Err(AlgebraError::DifferentLengths(ncols, v.len()))."""
        out = self.checked_mul_vec(v)
        if out is None:
            raise DifferentLengths(self.ncols, len(v))
        return out

    def __matmul__(self, v: RqNTT) -> RqNTT:
        return self.try_mul_vec(v)

    __mul__ = __matmul__

    def __imul__(self, r: RqNTT) -> "SparseMatrix":
        """sparse_matrix.rs:298-302: every stored entry *= r."""
        if len(self.vals):
            _scale(self.vals, r, self.ctx)
        return self
