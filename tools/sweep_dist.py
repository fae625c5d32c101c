#!/usr/bin/env python3
"""BASELINE config 5 on N GPUs: batch-size sweep for the three primes with one process per GPU (torchrun), every
rank owning its own batch of 2^log2n elements (weak scaling, no data-path collective).  Per point: barrier +
synchronize, CUDA events on the launching stream, MAX over ranks (NCCL all-reduce of the times), L2 flushed before
every timed launch; rank 0 prints one JSON line per point with the aggregate rate and the per-GPU roofline fraction.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_dist.py \
      [--ops ring_mul,crt] [--rings bb,gl,sp] [--log2ns 10,14,18,22,26] > profiles/r02_sweep_nN.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import stark_rings_b200 as S
from bench import ELEM_BYTES, gen_raw_device, peaks

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="ring_mul")
ap.add_argument("--rings", default="bb,gl,sp")
ap.add_argument("--log2ns", default="10,14,18,22,26")
a_ = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = S.Context(local)
ctx.use_torch_stream()
hbm, src = peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
free, _ = torch.cuda.mem_get_info()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for ring in a_.rings.split(","):
    cfg = S.CONFIGS[ring]
    for op in a_.ops.split(","):
        for l in [int(x) for x in a_.log2ns.split(",")]:
            n = 1 << l
            nbuf = 3 if op in ("ring_mul", "ntt_mul") else 1
            if n * ELEM_BYTES[ring] * nbuf > free * 0.85:
                continue
            a = gen_raw_device(torch, ring, n, 1 + 100 * rank, dev)
            b = gen_raw_device(torch, ring, n, 2 + 100 * rank, dev) if nbuf > 1 else None
            out = torch.empty_like(a) if nbuf > 1 else None
            if op == "ring_mul":
                fn, mult = (lambda: cfg.ring_mul_batch(a, b, out=out, ctx=ctx)), 3
            elif op == "ntt_mul":
                fn, mult = (lambda: cfg.ntt_mul_batch(a, b, ctx=ctx)), 3
            elif op == "crt":
                fn, mult = (lambda: cfg.crt_batch(a, ctx=ctx)), 2
            else:
                fn, mult = (lambda: cfg.icrt_batch(a, ctx=ctx)), 2
            for _ in range(3):
                fn()
            ts = []
            for _ in range(5):
                flush.fill_(1)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(max_over_ranks(e0.elapsed_time(e1)))
            ms = min(ts)
            nbytes = mult * n * ELEM_BYTES[ring]
            if rank == 0:
                print(json.dumps({"n_gpus": world, "ring": ring, "op": op, "log2n_per_gpu": l, "ms_max_over_ranks": round(ms, 5),
                                  "units_per_s_all_gpus": world * n / ms * 1e3, "GBps_per_gpu": nbytes / ms / 1e6,
                                  "frac_hbm_per_gpu": nbytes / ms / 1e6 / hbm}), flush=True)
            del a, b, out
            torch.cuda.empty_cache()
barrier()
if world > 1:
    dist.destroy_process_group()
