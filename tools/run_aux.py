#!/usr/bin/env python3
"""Throughput of the kernels next to the hot path on one GPU, CUDA events on the launching stream,
device-resident buffers: reduce / rot, gadget decompose / recompose, sparse mat-vec, dense mat-mat, scaling,
serialize / deserialize.  Prints one JSON line per measurement with the algorithmic bytes (each input and output byte
touched once) and the fraction of the measured HBM copy bandwidth.
Usage: tools/run_aux.py [--rings gl,bb,sp] [--log2n 20] > profiles/r01b_aux.jsonl"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import stark_rings_b200 as S
from bench import ELEM_BYTES, gen_raw_device, peaks

ap = argparse.ArgumentParser()
ap.add_argument("--rings", default="gl,bb,sp")
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--reps", type=int, default=5)
a_ = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = S.default_context(0)
ctx.use_torch_stream()
hbm, _ = peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(a_.reps):
        flush.fill_(1)
        ctx.use_torch_stream()
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop())
    return min(ts)


def emit(ring, op, n, ms, nbytes, **kw):
    print(json.dumps(dict(ring=ring, op=op, n=n, ms=round(ms, 5), units_per_s=n / ms * 1e3, GBps=nbytes / ms / 1e6,
                          frac_hbm=nbytes / ms / 1e6 / hbm, **kw)), flush=True)


n = 1 << a_.log2n
for ring in a_.rings.split(","):
    cfg = S.CONFIGS[ring]
    eb = ELEM_BYTES[ring]
    a = gen_raw_device(torch, ring, n, 1, dev)
    # coefficient-form helpers: 2D coefficients -> D (reduce), D -> D (rot)
    wide = torch.cat([a, gen_raw_device(torch, ring, n, 2, dev)])  # n/1 polynomials of 2D coefficients = 2n elements' worth
    emit(ring, "reduce_2D", n, timed(lambda: cfg.reduce_batch(wide, 2 * cfg.D, ctx)), 3 * n * eb)
    emit(ring, "rot", n, timed(lambda: cfg.rot_batch(a, ctx)), 2 * n * eb)
    del wide
    # canonical serialization
    x = S.RqPoly(cfg, a, ctx)
    sb = x.serialized_size()
    emit(ring, "serialize", n, timed(lambda: x.serialize()), n * eb + sb)
    blob = x.serialize()
    emit(ring, "deserialize", n, timed(lambda: S.RqPoly.deserialize(cfg, blob, ctx)), n * eb + sb)
    del blob
    # scalar scaling and slot-wise work on NTT form
    r = S.RqNTT(cfg, gen_raw_device(torch, ring, 1, 3, dev), ctx)
    row = S.Matrix([S.RqNTT(cfg, a.clone(), ctx)], ctx)

    def scale():
        global row
        row *= r
    emit(ring, "scale", n, timed(scale), 2 * n * eb)
    del row
    # gadget decomposition (Fp64 rings): basis 2^8, 8 digits (64-bit) / 4 digits (31-bit)
    if ring in ("gl", "bb"):
        pad, b = (8, 256) if ring == "gl" else (4, 256)
        m = n >> 3
        src = a[: m * cfg.limbs]
        emit(ring, "gadget_decompose_b256", m, timed(lambda: cfg.gadget_decompose(src, b, pad, ctx)), (1 + pad) * m * eb,
             padding_size=pad)
        digits = cfg.gadget_decompose(src, b, pad, ctx)
        emit(ring, "gadget_recompose_b256", m, timed(lambda: cfg.gadget_recompose(digits, b, pad, ctx)),
             (1 + pad) * m * eb, padding_size=pad)
        del digits
    # sparse mat-vec: constraint-matrix shape, nrows = ncols = n / 4, 4 entries per row at random columns
    nr = n >> 2
    nnz_row = 4
    rng = np.random.default_rng(7)
    row_ptr = torch.arange(0, (nr + 1) * nnz_row, nnz_row, dtype=torch.int64, device=dev)
    col_idx = torch.from_numpy(rng.integers(0, nr, size=nr * nnz_row).astype(np.int64)).to(dev)
    vals = S.RqNTT(cfg, a[: nr * nnz_row * cfg.limbs], ctx)  # nr * 4 = n elements
    v = S.RqNTT(cfg, gen_raw_device(torch, ring, nr, 4, dev), ctx)
    A = S.SparseMatrix(nr, nr, row_ptr, col_idx, vals, ctx)
    # bytes: every stored entry once, one gathered vector element per entry (L2 hits when the vector fits), the output
    emit(ring, "sparse_matvec_4_per_row", nr, timed(lambda: A.try_mul_vec(v)),
         (nr * nnz_row + nr * nnz_row + nr) * eb + nr * nnz_row * 8, nnz=nr * nnz_row,
         note="bytes count one gathered vector element per entry")
    # dense mat-mat: (32 x 256) * (256 x n/256)
    k, mc = 256, max(1, n >> 8)
    Am = S.Matrix([S.RqNTT(cfg, gen_raw_device(torch, ring, k, 10 + i, dev), ctx) for i in range(32)], ctx)
    Bm = S.Matrix([S.RqNTT(cfg, a[j * mc * cfg.limbs:(j + 1) * mc * cfg.limbs], ctx) for j in range(k)], ctx)
    emit(ring, "matmat_32x256_256xM", mc, timed(lambda: Am.try_mul_mat(Bm)), (32 * k + k * mc + 32 * mc) * eb, m_ncols=mc,
         ring_mul_adds=32 * k * mc)
    del Am, Bm, A, vals, v, a
    torch.cuda.empty_cache()
