#!/usr/bin/env python3
"""Run one hot-path op a few times on the GPU and print CUDA-event timings (used for tuning and as
the command profiled by ncu).  Usage: tools/run_op.py --ring bb --op ring_mul --log2n 22 --reps 5"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import stark_rings_b200 as S
from bench import gen_raw_device, ELEM_BYTES

ap = argparse.ArgumentParser()
ap.add_argument("--ring", default="bb"); ap.add_argument("--op", default="ring_mul")
ap.add_argument("--log2n", type=int, default=22); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--kappa", type=int, default=4)
a_ = ap.parse_args()
cfg = S.CONFIGS[a_.ring]; n = 1 << a_.log2n; dev = torch.device("cuda", 0)
ctx = S.default_context(0); ctx.use_torch_stream()
a = gen_raw_device(torch, a_.ring, n, 1, dev); b = gen_raw_device(torch, a_.ring, n, 2, dev)
out = torch.empty_like(a)
if a_.op == "matvec":
    rows = [S.RqNTT(cfg, gen_raw_device(torch, a_.ring, n, 10 + i, dev), ctx) for i in range(a_.kappa)]
    A = S.Matrix(rows, ctx); v = S.RqNTT(cfg, a, ctx)
    fn = lambda: A.try_mul_vec(v); nbytes = (a_.kappa * n + n + a_.kappa) * ELEM_BYTES[a_.ring]
elif a_.op == "ring_mul":
    fn = lambda: cfg.ring_mul_batch(a, b, out=out, ctx=ctx); nbytes = 3 * n * ELEM_BYTES[a_.ring]
elif a_.op == "ntt_mul":
    fn = lambda: cfg.ntt_mul_batch(a, b, ctx=ctx); nbytes = 3 * n * ELEM_BYTES[a_.ring]
elif a_.op == "crt":
    fn = lambda: cfg.crt_batch(a, ctx=ctx); nbytes = 2 * n * ELEM_BYTES[a_.ring]
else:
    fn = lambda: cfg.icrt_batch(a, ctx=ctx); nbytes = 2 * n * ELEM_BYTES[a_.ring]
for _ in range(2): fn()
torch.cuda.synchronize()
ts = []
for _ in range(a_.reps):
    ctx.timer_start(); fn(); ts.append(ctx.timer_stop())
ms = min(ts)
print(json.dumps({"ring": a_.ring, "op": a_.op, "log2n": a_.log2n, "ms_min": ms, "ms_all": [round(t, 4) for t in ts],
                  "units_per_s": n / ms * 1e3, "GBps": nbytes / ms / 1e6, "frac_hbm_6536": nbytes / ms / 1e6 / 6536.4}))
