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
  checked_mul_mat / try_mul_mat / Mul<&SparseMatrix>  sparse_matrix.rs:219-275 -> SparseMatrix.checked_mul_mat / try_mul_mat
  CanonicalSerialize / CanonicalDeserialize    matrix.rs:111-145, sparse_matrix.rs:157-200 -> serialize / deserialize
The ring mat-vec is the hot path; the rest are the callers' neighbouring linear maps.  Padding, hconcat and rand stay
with the caller.
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

    # -- CanonicalSerialize / CanonicalDeserialize (matrix.rs:111-145): self.vals: Vec<Vec<R>>, i.e. ark-serialize's
    # u64 little-endian length prefix of the outer Vec, then per row its own u64 length and the elements back to back
    # (each element = D field elements, standard-form little-endian, no prefix: coeff_form.rs:154-189).  The element
    # bytes are produced on the device for device-resident rows (sr_serialize_batch); the framing is host work.
    def serialize(self) -> np.ndarray:
        parts = [np.array([self.nrows], dtype="<u8").view(np.uint8)]
        for r in self.vals:
            parts.append(np.array([len(r)], dtype="<u8").view(np.uint8))
            b = r.serialize()
            parts.append(b if isinstance(b, np.ndarray) else b.cpu().numpy())
        return np.concatenate(parts)

    @classmethod
    def deserialize(cls, config, data: np.ndarray, device=None, ctx=None) -> "Matrix":
        """matrix.rs:131-145: nrows = vals.len(), ncols = the first row's length.  Raises LengthPanic on a truncated
        buffer and StarkRingsError on a non-canonical field element (SerializationError::InvalidData)."""
        from .errors import LengthPanic
        data = np.ascontiguousarray(data, dtype=np.uint8)
        per = int(L.lib.sr_serialized_bytes(config.ring_id, 1))
        pos = 0

        def u64():
            nonlocal pos
            if pos + 8 > data.size:
                raise LengthPanic("serialized matrix is truncated")
            v = int(data[pos:pos + 8].view("<u8")[0])
            pos += 8
            return v
        rows = []
        for _ in range(u64()):
            n = u64()
            if pos + n * per > data.size:
                raise LengthPanic("serialized matrix is truncated")
            chunk = data[pos:pos + n * per]
            pos += n * per
            src = torch.from_numpy(chunk.copy()).to(device) if device is not None else chunk.copy()
            rows.append(RqNTT.deserialize(config, src, ctx))
        m = cls.__new__(cls)
        m.vals, m.nrows, m.ctx, m.config = rows, len(rows), ctx, config
        m.ncols = len(rows[0]) if rows else 0
        return m

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

    def checked_mul_mat(self, m: "SparseMatrix"):
        """sparse_matrix.rs:219-275: None when self.ncols != m.nrows.  The structure pass (columns of m in row order,
        merge join of the index lists in stored order) runs on the host exactly as in the reference
        (sr_sparse_matmat_symbolic); the products and sums run on the device (sr_sparse_matmat_values); an output entry
        exists iff one of its products is not the zero element, and the rows come out ordered by column."""
        if self.ncols != m.nrows:
            return None
        cfg = self.config
        if m.config is not cfg:
            raise TypeError("matrices are over different rings")
        host = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy().view(np.uint64)
        arp, aci = np.ascontiguousarray(host(self.row_ptr)), np.ascontiguousarray(host(self.col_idx))
        mrp, mci = np.ascontiguousarray(host(m.row_ptr)), np.ascontiguousarray(host(m.col_idx))
        vp = lambda a: ctypes.c_void_p(a.ctypes.data) if a.size else None
        ncand, npairs = ctypes.c_size_t(0), ctypes.c_size_t(0)
        sym = lambda *outs: L.lib.sr_sparse_matmat_symbolic(self.nrows, vp(arp), vp(aci), m.nrows, m.ncols, vp(mrp), vp(mci),
                                                            ctypes.byref(ncand), ctypes.byref(npairs), *outs)
        if sym(None, None, None, None, None) != L.SR_OK:
            raise ValueError("malformed CSR arrays (the reference panics on an out-of-range column of m)")
        nc, npr = ncand.value, npairs.value
        cand_row, cand_col = np.empty(nc, dtype=np.uint64), np.empty(nc, dtype=np.uint64)
        pair_ptr = np.zeros(nc + 1, dtype=np.uint64)
        pair_a, pair_m = np.empty(max(npr, 1), dtype=np.uint64), np.empty(max(npr, 1), dtype=np.uint64)
        if nc:
            sym(vp(cand_row), vp(cand_col), vp(pair_ptr), vp(pair_a), vp(pair_m))
        pa, _, loc, dev = _ptr_loc(self.vals.data) if len(self.vals) else (None, 0, L.SR_HOST, None)
        pm, _, locm, _ = _ptr_loc(m.vals.data) if len(m.vals) else (None, 0, loc, None)
        if len(self.vals) and len(m.vals) and locm != loc:
            raise ValueError("both matrices must live in the same place")
        on_dev = loc == L.SR_DEVICE
        like = self.vals.data
        out_vals = (torch.empty(max(nc, 1) * cfg.limbs, dtype=like.dtype, device=like.device) if on_dev
                    else np.empty(max(nc, 1) * cfg.limbs, dtype=np.uint64))
        nz = np.zeros(max(nc, 1), dtype=np.int32)
        c = self.ctx or self.vals.ctx or default_context(0 if dev is None else dev)
        if nc:
            if on_dev:
                c.use_torch_stream()
            c.check(L.lib.sr_sparse_matmat_values(c.h, cfg.ring_id, pa, pm, nc, vp(pair_ptr), vp(pair_a), vp(pair_m),
                                                  _ptr_loc(out_vals)[0], ctypes.c_void_p(nz.ctypes.data), loc),
                    "sr_sparse_matmat_values")
        keep = np.nonzero(nz[:nc])[0]
        row_ptr = np.zeros(self.nrows + 1, dtype=np.uint64)
        np.add.at(row_ptr, cand_row[keep].astype(np.int64) + 1, 1)
        row_ptr = np.cumsum(row_ptr).astype(np.uint64)
        col_idx = cand_col[keep]
        w = cfg.limbs
        if on_dev:
            idx = torch.from_numpy(keep.astype(np.int64)).to(like.device)
            vals = out_vals.view(-1, w)[: max(nc, 1)].index_select(0, idx).reshape(-1).contiguous()
            to = lambda a: torch.from_numpy(a.view(np.int64)).to(like.device)
            return SparseMatrix(self.nrows, m.ncols, to(row_ptr), to(col_idx), RqNTT(cfg, vals, c), c)
        vals = out_vals.reshape(-1, w)[keep].reshape(-1).copy()
        return SparseMatrix(self.nrows, m.ncols, row_ptr, col_idx, RqNTT(cfg, vals, c), c)

    def try_mul_mat(self, m: "SparseMatrix") -> "SparseMatrix":
        """sparse_matrix.rs:271-275: Err(AlgebraError::DifferentLengths(self.ncols, m.nrows))."""
        out = self.checked_mul_mat(m)
        if out is None:
            raise DifferentLengths(self.ncols, m.nrows)
        return out

    # -- CanonicalSerialize / CanonicalDeserialize (sparse_matrix.rs:157-200): nrows and ncols as u64, then
    # coeffs: Vec<Vec<(R, usize)>>: u64 row count, per row its u64 length and the (element bytes, u64 column) tuples.
    def serialize(self) -> np.ndarray:
        host = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy().view(np.uint64)
        rp, ci = host(self.row_ptr), host(self.col_idx)
        b = self.vals.serialize() if len(self.vals) else np.empty(0, dtype=np.uint8)
        b = b if isinstance(b, np.ndarray) else b.cpu().numpy()
        per = int(L.lib.sr_serialized_bytes(self.config.ring_id, 1))
        nnz = len(ci)
        rec = np.empty((nnz, per + 8), dtype=np.uint8)
        rec[:, :per] = b.reshape(nnz, per)
        rec[:, per:] = np.ascontiguousarray(ci, dtype="<u8").view(np.uint8).reshape(nnz, 8)
        parts = [np.array([self.nrows, self.ncols, self.nrows], dtype="<u8").view(np.uint8)]
        for i in range(self.nrows):
            e0, e1 = int(rp[i]), int(rp[i + 1])
            parts.append(np.array([e1 - e0], dtype="<u8").view(np.uint8))
            parts.append(rec[e0:e1].reshape(-1))
        return np.concatenate(parts)

    @classmethod
    def deserialize(cls, config, data: np.ndarray, device=None, ctx=None) -> "SparseMatrix":
        from .errors import LengthPanic
        data = np.ascontiguousarray(data, dtype=np.uint8)
        per = int(L.lib.sr_serialized_bytes(config.ring_id, 1))
        pos = 0

        def u64():
            nonlocal pos
            if pos + 8 > data.size:
                raise LengthPanic("serialized sparse matrix is truncated")
            v = int(data[pos:pos + 8].view("<u8")[0])
            pos += 8
            return v
        nrows, ncols, outer = u64(), u64(), u64()
        row_ptr = np.zeros(outer + 1, dtype=np.uint64)
        elem_bytes, cols = [], []
        for i in range(outer):
            n = u64()
            if pos + n * (per + 8) > data.size:
                raise LengthPanic("serialized sparse matrix is truncated")
            rec = data[pos:pos + n * (per + 8)].reshape(n, per + 8)
            pos += n * (per + 8)
            elem_bytes.append(rec[:, :per].reshape(-1))
            cols.append(np.ascontiguousarray(rec[:, per:]).view("<u8").reshape(-1))
            row_ptr[i + 1] = row_ptr[i] + n
        eb = np.concatenate(elem_bytes) if elem_bytes else np.empty(0, dtype=np.uint8)
        col_idx = (np.concatenate(cols) if cols else np.empty(0, dtype=np.uint64)).astype(np.uint64)
        src = torch.from_numpy(eb.copy()).to(device) if device is not None else eb.copy()
        vals = RqNTT.deserialize(config, src, ctx)
        if device is not None:
            to = lambda a: torch.from_numpy(a.view(np.int64)).to(device)
            return cls(nrows, ncols, to(row_ptr), to(col_idx), vals, ctx)
        return cls(nrows, ncols, row_ptr, col_idx, vals, ctx)

    def __imul__(self, r: RqNTT) -> "SparseMatrix":
        """sparse_matrix.rs:298-302: every stored entry *= r."""
        if len(self.vals):
            _scale(self.vals, r, self.ctx)
        return self
