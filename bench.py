#!/usr/bin/env python3
"""Benchmark of the stark-rings hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload ringmul|commit] [--ring bb|gl|sp] [--log2n L]

Default workload = BASELINE.json configs[1]: BabyBear ring, batched CRT -> slot mul -> ICRT on 2^24
elements per GPU (weak scaling: every rank owns its own 2^24-element shard, no data-path
collective).  A "step" is one pass of the fused ring-mul kernel over the rank's resident batch
(inputs 2 x 9.66 GB >> 126 MB L2, so no L2 flush is needed between steps).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RING_NAMES = {"bb": "babybear", "gl": "goldilocks", "sp": "stark_prime"}
ELEM_BYTES = {"bb": 576, "gl": 192, "sp": 512}
P = {"bb": 2013265921, "gl": 18446744069414584321}


def measured_traffic(tag, workload, log2n):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f).get("%s:%s:%d" % (tag, workload, log2n))
        return None if t is None else t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        """Starts nvidia-smi (20 ms period) and waits until it delivers samples, so that short timed
        regions are covered; mark_begin()/mark_end() bracket the timed region."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t_begin is not None and self.t_begin <= t <= (self.t_end or t) + 0.03]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: use the samples taken under the same load
            inside = [r for _, r in self.rows]
            window = "warm-up + timed region (timed region shorter than the 20 ms sampling period)"
        for r in inside:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def gen_raw_device(torch, tag, n, seed, device):
    """n elements of uniformly random canonical residues as raw limbs, generated on the device."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if tag == "bb":
        return torch.randint(0, P["bb"], (n * 72,), dtype=torch.int64, device=device, generator=g)
    if tag == "gl":
        # uniform 64-bit patterns, folded below p (p = 2^64 - 2^32 + 1: subtract p from the top 2^32-1 values)
        x = torch.randint(-(1 << 63), (1 << 63) - 1, (n * 24,), dtype=torch.int64, device=device, generator=g)
        # as unsigned, x >= p  <=>  high word == 0xFFFFFFFF and low word != 0
        hi_all = (x >> 32) == -1
        lo_nz = (x & 0xFFFFFFFF) != 0
        return torch.where(hi_all & lo_nz, x + 0xFFFFFFFF, x)  # x - p  ==  x + 2^32 - 1 (mod 2^64)
    x = torch.randint(-(1 << 63), (1 << 63) - 1, (n * 64,), dtype=torch.int64, device=device, generator=g)
    x[3::4] &= (1 << 59) - 1  # every field element < 2^251 < p
    return x


def make_config(args, world):
    """metric / unit / config shared by both arms (ours and --impl reference)."""
    tag = args.ring
    n = 1 << args.log2n
    if args.workload == "ringmul":
        cfg = {"workload": "%s ring: batched CRT->slot mul->ICRT on 2^%d elements per GPU" % (RING_NAMES[tag], args.log2n),
               "ring": RING_NAMES[tag], "log2_elements_per_gpu": args.log2n,
               "l2": "inputs >> L2 (no flush needed)" if n * ELEM_BYTES[tag] > (256 << 20) else "inputs fit L2",
               "layout": "reference layout: u64 Montgomery limbs, %d B per element" % ELEM_BYTES[tag]}
        return "ring muls/sec (CRT->mul->ICRT)", "ring_mul/s", cfg
    cfg = {"workload": "%s commit: %d x 2^%d ring matrix x vector, columns sharded over %d GPU(s)" % (
               RING_NAMES[tag], args.kappa, args.log2n, world),
           "ring": RING_NAMES[tag], "log2_columns": args.log2n, "kappa": args.kappa,
           "l2": "matrix >> L2" if args.kappa * n * ELEM_BYTES[tag] > (256 << 20) else "matrix fits L2",
           "layout": "reference layout: u64 Montgomery limbs, %d B per element" % ELEM_BYTES[tag]}
    return "commits/sec (kappa x m ring matrix x vector)", "commit/s", cfg


def cpu_reference_rate(tag, sample_elems, threads, repeats=1):
    """ring muls/s of the C restatement (oracle/sr_oracle.c, -march=native) on this host."""
    import numpy as np
    from oracle import c_oracle as C
    from tests.util import rand_raw
    name = RING_NAMES[tag]
    lib, kind = C.lib_native()
    a, b = rand_raw(name, sample_elems, 11, edge=False), rand_raw(name, sample_elems, 12, edge=False)
    C.ring_mul(name, a[: 1024 * C.words(name)], b[: 1024 * C.words(name)], threads=1, L=lib)  # warm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        C.ring_mul(name, a, b, threads=threads, L=lib)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return sample_elems / best, best, kind


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (C restatement: the Rust crate cannot be
    built here) on the box's host cores, all threads, same metric / unit / config as our arm.
    Each step is a bounded sample of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import numpy as np
    from oracle import c_oracle as C
    from tests.util import rand_raw
    tag = args.ring
    cores = os.cpu_count() or 1
    metric, unit, cfg = make_config(args, world)
    times = []
    if args.workload == "ringmul":
        sample = min(1 << args.log2n, args.cpu_sample)
        for i in range(args.warmup + args.steps):
            rate, dt, kind = cpu_reference_rate(tag, sample, cores)
            if i >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        value = sample / (ms / 1e3)
        what = "%d elements per step" % sample
    else:
        # commit: a bounded number of columns; the reference parallelises over the kappa rows only
        name = RING_NAMES[tag]
        lib, kind = C.lib_native()
        m = min(1 << args.log2n, 1 << 16)
        rows = [rand_raw(name, m, 40 + i, edge=False) for i in range(args.kappa)]
        v = rand_raw(name, m, 50, edge=False)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            C.matvec(name, rows, v, threads=min(cores, args.kappa), L=lib)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        ms_sample = 1e3 * sum(times) / len(times)
        ms = ms_sample * ((1 << args.log2n) / m)  # scaled to the full column count
        value = 1e3 / ms
        what = "%d of 2^%d columns per step (time scaled linearly), %d row threads" % (m, args.log2n, min(cores, args.kappa))
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if args.workload == "commit" else "weak", "vs_baseline": None,
        "dtype": "u32" if tag == "bb" else "u64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": "%s, %s build of oracle/sr_oracle.c (C restatement of the reference algorithm; "
                                   "the Rust crate cannot be built in this image: no cargo/rustc)" % (what, kind)},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner there) get stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ringmul", choices=["ringmul", "commit"])
    ap.add_argument("--ring", default=None, choices=["bb", "gl", "sp"])
    ap.add_argument("--log2n", type=int, default=None, help="log2 elements per GPU (ringmul) / columns (commit)")
    ap.add_argument("--kappa", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=1 << 20, help="elements in the CPU baseline sample")
    ap.add_argument("--e2e-log2n", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--commit-path", default="peer", choices=["peer", "nccl"],
                    help="commit workload, N > 1: how the per-rank partials reach rank 0")
    ap.add_argument("--graph", action="store_true", help="commit workload, 1 GPU: capture the step in a CUDA graph")
    args = ap.parse_args()
    if args.ring is None:
        args.ring = "gl" if args.workload == "commit" else "bb"
    if args.log2n is None:
        args.log2n = 20 if args.workload == "commit" else {"bb": 24, "gl": 24, "sp": 20}[args.ring]
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import stark_rings_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = S.Context(local)
    ctx.use_torch_stream()
    tag = args.ring
    cfg = S.CONFIGS[tag]
    hbm_peak, peak_src = peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = 1 << args.log2n
    line = {}
    metric, unit, config = make_config(args, world)
    if args.workload == "ringmul":
        a = gen_raw_device(torch, tag, n, 0x5EED ^ (1000 * rank + 1), dev)
        b = gen_raw_device(torch, tag, n, 0x5EED ^ (1000 * rank + 2), dev)
        out = torch.empty_like(a)
        step = lambda: cfg.ring_mul_batch(a, b, out=out, ctx=ctx)
        units_per_step = n
        alg_bytes_per_unit = 3 * ELEM_BYTES[tag]
    else:
        # Goldilocks Ajtai-style commit: kappa x m matrix times vector, columns sharded over ranks
        m_local = n // world
        rows = [S.RqNTT(cfg, gen_raw_device(torch, tag, m_local, 0x5EED ^ (1000 * rank + 10 + i), dev), ctx)
                for i in range(args.kappa)]
        v = S.RqNTT(cfg, gen_raw_device(torch, tag, m_local, 0x5EED ^ (1000 * rank + 3), dev), ctx)
        A = S.Matrix(rows, ctx)
        limbs = cfg.limbs
        gathered = torch.empty(world * args.kappa * limbs, dtype=torch.int64, device=dev)
        result = torch.empty(args.kappa * limbs, dtype=torch.int64, device=dev)
        import ctypes
        from stark_rings_b200 import _lib as L

        peer = None
        if dist is not None and args.commit_path == "peer":
            # partials go straight into rank 0's HBM over NVLink from the kernel that produces them; rank 0's
            # reduction kernel acquires the per-rank flags (stark_rings_b200/dist.py PeerCommit)
            from stark_rings_b200.dist import PeerCommit
            peer = PeerCommit(cfg, args.kappa, world, rank, ctx, device_epochs=args.graph)
        config["exchange"] = "none (1 GPU)" if dist is None else (
            "NVLink peer-memory mailbox fused into the producing / reducing kernels" if peer else
            "NCCL all_gather of raw limbs + rank-0 modular sum")

        def step():
            if peer is not None:
                return peer.commit(A, v, out=result)
            part = A.partial_mul_vec(v)
            if dist is not None:
                dist.all_gather_into_tensor(gathered, part.data)  # raw limbs; an NCCL sum cannot reduce mod p
                if rank == 0:
                    ctx.check(L.lib.sr_modsum_partials(ctx.h, cfg.ring_id, ctypes.c_void_p(gathered.data_ptr()),
                                                       world, args.kappa, ctypes.c_void_p(result.data_ptr()),
                                                       L.SR_DEVICE), "modsum")
            return part
        units_per_step = 1
        alg_bytes_per_unit = (args.kappa * m_local + m_local + args.kappa) * ELEM_BYTES[tag]

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = None
    if args.workload == "commit" and args.graph and (world == 1 or peer is not None):
        # optional: capture the step's kernels in a CUDA graph (the row table is already resident, see sr_capi.cu).
        # 1 GPU: gain 1% at kappa = 4, m = 2^20.  N GPUs: only the peer-memory path can be captured (device-resident
        # epochs, no collective call in the step); with NCCL in the step the capture hung in round 1.
        try:
            l0 = ctx.kernel_launches
            step()
            launches_per_step = ctx.kernel_launches - l0
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            step = graph.replay
            for _ in range(3):
                step()
            config["cuda_graph"] = True
        except Exception as ex:  # keep the eager path
            config["cuda_graph"] = "capture failed: %s" % str(ex)[:80]
            torch.cuda.synchronize()
        barrier()
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.kernel_launches - launches0
    if launches == 0 and launches_per_step:  # graph replays do not pass through the launch counter
        launches = launches_per_step * args.steps
    ms_per_step = total_ms / args.steps
    strong = args.workload == "commit"
    value = (units_per_step * (1 if strong else world)) / (ms_per_step / 1e3)

    # roofline of the dominant kernel: its own launches, timed alone with events on the same stream
    ctx.use_torch_stream()  # (a graph capture leaves the context on the capture stream)
    ktimes = []
    for _ in range(min(args.steps, 5)):
        ctx.timer_start()
        step()
        ktimes.append(ctx.timer_stop())
    k_ms = sum(ktimes) / len(ktimes)
    if args.workload == "ringmul":
        achieved = alg_bytes_per_unit * units_per_step / (k_ms / 1e3) / 1e9
    else:
        achieved = alg_bytes_per_unit / (k_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": measured_traffic(tag, args.workload, args.log2n), "peak_source": peak_src, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes_per_unit * (units_per_step if args.workload == "ringmul" else 1)}

    # ---- end to end through the C ABI with HOST buffers -----------------------------------------
    e2e = None
    if not args.no_e2e and args.workload == "ringmul":
        e2e_n = 1 << (args.e2e_log2n if args.e2e_log2n is not None else args.log2n)
        e2e = None
        while e2e is None and e2e_n >= 1 << 10:
            try:
                words = e2e_n * cfg.limbs
                ha = torch.empty(words, dtype=torch.int64).pin_memory()
                hb = torch.empty(words, dtype=torch.int64).pin_memory()
                ho = torch.empty(words, dtype=torch.int64).pin_memory()
                ha.copy_(a[:words])
                hb.copy_(b[:words])
                torch.cuda.synchronize()
                cfg.ring_mul_batch(ha, hb, out=ho, ctx=ctx)  # warm-up (allocates the staging ring)
                k = max(1, min(args.steps, 3))
                barrier()
                t0 = time.perf_counter()
                for _ in range(k):
                    cfg.ring_mul_batch(ha, hb, out=ho, ctx=ctx)  # synchronous: returns when `ho` is complete
                torch.cuda.synchronize()
                dt = max_over_ranks((time.perf_counter() - t0) / k)
                if rank == 0:
                    assert torch.equal(ho[: 64 * cfg.limbs].to(dev), out[: 64 * cfg.limbs])
                e2e = {"value": e2e_n * world / dt, "unit": unit, "h2d_bytes_per_step": 2 * words * 8,
                       "d2h_bytes_per_step": words * 8, "elements_per_gpu": e2e_n, "ms_per_step": dt * 1e3,
                       "note": "sr_ring_mul_batch(SR_HOST) on pinned host buffers: chunked H2D -> kernel -> D2H pipeline"}
                del ha, hb, ho
            except RuntimeError as ex:  # pinned allocation failed: halve
                e2e = None
                e2e_n //= 2
                last = str(ex)[:120]
        if e2e is None:
            e2e = {"value": None, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "note": "pinned host allocation failed: " + last}

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1 and args.workload == "ringmul":
        cores = os.cpu_count() or 1
        sample = min(n, args.cpu_sample)
        rate, dt, kind = cpu_reference_rate(tag, sample, cores)
        rate1, dt1, _ = cpu_reference_rate(tag, max(1024, sample // 16), 1)
        cpu = {"value": rate, "unit": unit, "cores": cores, "kind": "port",
               "sample": "%d elements (%.1f s wall on %d threads), %s build of oracle/sr_oracle.c" % (sample, dt, cores, kind),
               "value_1thread": rate1}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "u32" if tag == "bb" else "u64", "data": "synthetic",
            "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
