"""Mirror of stark-rings-linalg's dense Matrix for R = RqNTT (linear_algebra/src/matrix.rs).

  Matrix { nrows, ncols, vals: Vec<Vec<R>> }   matrix.rs:17-21  -> Matrix(rows)
  checked_mul_vec                              matrix.rs:168-178 -> Matrix.checked_mul_vec (None on mismatch)
  try_mul_vec                                  matrix.rs:180-183 -> Matrix.try_mul_vec (raises DifferentLengths)
  Mul<&[R]> for &Matrix<R>                     matrix.rs:199-205 -> Matrix.__matmul__ / __mul__
Only the ring mat-vec is on the hot path; mul_mat, padding etc. are out of scope.
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

    def partial_mul_vec(self, v: RqNTT) -> RqNTT:
        """One rank's share of a column-sharded commitment (sr_matvec_partial)."""
        out = self._call(L.lib.sr_matvec_partial, v)
        if out is None:
            raise DifferentLengths(self.ncols, len(v))
        return out
