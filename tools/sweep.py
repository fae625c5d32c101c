#!/usr/bin/env python3
"""BASELINE config 5: batch-size sweep 2^10 .. 2^26 for the three primes on one GPU, roofline fraction per
point.  Times with CUDA events on the launching stream; between launches of the small (L2-resident) points
an L2 flush (write of a 256 MiB buffer) is issued so that every point reads from HBM.
Usage: tools/sweep.py [--ops ring_mul,crt] [--max-log2n 26] > profiles/r01_sweep.jsonl"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import stark_rings_b200 as S
from bench import gen_raw_device, ELEM_BYTES, peaks

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="ring_mul,crt")
ap.add_argument("--rings", default="bb,gl,sp")
ap.add_argument("--min-log2n", type=int, default=10)
ap.add_argument("--max-log2n", type=int, default=26)
ap.add_argument("--step", type=int, default=2)
a_ = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = S.default_context(0); ctx.use_torch_stream()
hbm, src = peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
free, total = torch.cuda.mem_get_info()
for ring in a_.rings.split(","):
    cfg = S.CONFIGS[ring]
    for op in a_.ops.split(","):
        for l in range(a_.min_log2n, a_.max_log2n + 1, a_.step):
            n = 1 << l
            nbuf = 3 if op in ("ring_mul", "ntt_mul") else 1
            if n * ELEM_BYTES[ring] * nbuf > free * 0.9:
                continue
            a = gen_raw_device(torch, ring, n, 1, dev)
            b = gen_raw_device(torch, ring, n, 2, dev) if nbuf > 1 else None
            out = torch.empty_like(a) if nbuf > 1 else None
            if op == "ring_mul":
                fn = lambda: cfg.ring_mul_batch(a, b, out=out, ctx=ctx); mult = 3
            elif op == "ntt_mul":
                fn = lambda: cfg.ntt_mul_batch(a, b, ctx=ctx); mult = 3
            elif op == "crt":
                fn = lambda: cfg.crt_batch(a, ctx=ctx); mult = 2
            else:
                fn = lambda: cfg.icrt_batch(a, ctx=ctx); mult = 2
            for _ in range(3):
                fn()
            ts = []
            for _ in range(5):
                flush.fill_(1)
                ctx.timer_start(); fn(); ts.append(ctx.timer_stop())
            ms = min(ts); nbytes = mult * n * ELEM_BYTES[ring]
            print(json.dumps({"ring": ring, "op": op, "log2n": l, "ms": round(ms, 5), "units_per_s": n / ms * 1e3,
                              "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / hbm}), flush=True)
            del a, b, out
            torch.cuda.empty_cache()
