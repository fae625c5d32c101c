#!/usr/bin/env python3
"""Benchmark of the stark-rings hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload ringmul|commit] [--ring bb|gl|sp] [--log2n L] [--no-extra]

Headline = BASELINE.json configs[1]: BabyBear ring, batched CRT -> slot mul -> ICRT on 2^24 elements per GPU
(weak scaling: every rank owns its own 2^24-element shard, no data-path collective).  A "step" is one pass of the
fused ring-mul kernel over the rank's resident batch (inputs 2 x 9.66 GB >> 126 MB L2: no L2 flush needed).
The same run then measures the other BASELINE configs and reports them in the line's "extra" list, each with its own
roofline and clock record: Goldilocks 2^16 and 2^24 ring mul, Starknet-prime 2^20 ring mul against the integer
multiply-add peak measured IN THIS RUN (sr_imad_peak), and the Goldilocks 4 x 2^20 commitment strong-scaled over the
N ranks of the run (one kernel per rank per commitment, partials exchanged through the root's NVLink mailbox).
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
# one Starknet-prime ring mul = 120 modular multiplications x 88 (32 x 32 + 64 -> 64) multiply-adds
SP_IMAD_PER_RING_MUL = 10560
L2_BYTES = 126 << 20


def measured_traffic(tag, workload, log2n):
    """DRAM bytes per launch of the dominant kernel from the committed ncu captures (profiles/r0?_traffic.json)."""
    key = "%s:%s:%d" % (tag, workload, log2n)
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f).get(key)
            if t is not None:
                return t["dram_bytes_read"] + t["dram_bytes_write"]
        except Exception:
            pass
    return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions; one process for the whole run,
    mark_begin()/mark_end() bracket each region and return its own summary."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.t_begin = None

    def start(self):
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
        """Summary of the samples taken since mark_begin()."""
        t_end = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        inside = [r for t, r in self.rows if self.t_begin <= t <= t_end + 0.03]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: the nearest samples, taken under the same load
            near = sorted(self.rows, key=lambda tr: abs(tr[0] - t_end))[:3]
            inside = [r for _, r in near]
            window = "nearest samples (timed region shorter than the 20 ms sampling period)"
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
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

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()


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
               "l2": "inputs >> L2 (no flush needed)" if n * ELEM_BYTES[tag] > (256 << 20) else "inputs fit L2: L2 flushed between steps",
               "layout": "reference layout: u64 Montgomery limbs, %d B per element" % ELEM_BYTES[tag]}
        return "ring muls/sec (CRT->mul->ICRT)", "ring_mul/s", cfg
    cfg = {"workload": "%s commit: %d x 2^%d ring matrix x vector, columns sharded over %d GPU(s)" % (
               RING_NAMES[tag], args.kappa, args.log2n, world),
           "ring": RING_NAMES[tag], "log2_columns": args.log2n, "kappa": args.kappa,
           "layout": "reference layout: u64 Montgomery limbs, %d B per element" % ELEM_BYTES[tag]}
    return "commits/sec (kappa x m ring matrix x vector)", "commit/s", cfg


def cpu_reference_rate(tag, sample_elems, threads, repeats=1, inputs=None):
    """ring muls/s of the C restatement (oracle/sr_oracle.c, -march=native) on this host."""
    from oracle import c_oracle as C
    from tests.util import rand_raw
    name = RING_NAMES[tag]
    lib, kind = C.lib_native()
    if inputs is None:
        inputs = (rand_raw(name, sample_elems, 11, edge=False), rand_raw(name, sample_elems, 12, edge=False))
    a, b = inputs
    C.ring_mul(name, a[: 1024 * C.words(name)], b[: 1024 * C.words(name)], threads=1, L=lib)  # warm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        C.ring_mul(name, a, b, threads=threads, L=lib)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return sample_elems / best, best, kind


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (C restatement: the Rust crate cannot be built here) on
    the box's host cores, all threads, same metric / unit / config as our arm.  Rank 0 only.  The ring-mul arm runs
    the FULL configuration per step when the host has the memory for it (2 x 9.66 GB of inputs + 9.66 GB of output
    for BabyBear 2^24); otherwise, or with --cpu-sample, a bounded sample, stated in the line."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import c_oracle as C
    from tests.util import rand_raw
    tag = args.ring
    cores = os.cpu_count() or 1
    metric, unit, cfg = make_config(args, world)
    times = []
    if args.workload == "ringmul":
        full = 1 << args.log2n
        sample = full
        if args.cpu_sample:
            sample = min(full, args.cpu_sample)
        else:
            try:
                import psutil
                need = 3.3 * full * ELEM_BYTES[tag]
                if psutil.virtual_memory().available < need:
                    sample = min(full, 1 << 22)
            except Exception:
                sample = min(full, 1 << 22)
        name = RING_NAMES[tag]
        inputs = (rand_raw(name, sample, 11, edge=False), rand_raw(name, sample, 12, edge=False))
        for i in range(args.warmup + args.steps):
            rate, dt, kind = cpu_reference_rate(tag, sample, cores, inputs=inputs)
            if i >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        value = sample / (ms / 1e3)
        what = ("the full 2^%d elements per step" % args.log2n) if sample == full else \
            ("%d of 2^%d elements per step (rate-based metric)" % (sample, args.log2n))
        cfg["reference_arm_elements_per_step"] = sample
    else:
        # commit: the reference parallelises over the kappa rows only (matrix.rs:174)
        name = RING_NAMES[tag]
        lib, kind = C.lib_native()
        m = 1 << args.log2n
        rows = [rand_raw(name, m, 40 + i, edge=False) for i in range(args.kappa)]
        v = rand_raw(name, m, 50, edge=False)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            C.matvec(name, rows, v, threads=min(cores, args.kappa), L=lib)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        value = 1e3 / ms
        what = "the full 2^%d columns per step, %d row threads" % (args.log2n, min(cores, args.kappa))
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if args.workload == "commit" else "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": cfg,
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


class Env:
    """Per-process state shared by the measurements."""

    def __init__(self, args):
        import torch
        import stark_rings_b200 as S
        self.torch, self.S, self.args = torch, S, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.dev = torch.device("cuda", self.local)
        self.ctx = S.Context(self.local)
        self.ctx.use_torch_stream()
        self.hbm_peak, self.peak_src = peaks()
        self.sampler = ClockSampler(self.local)
        self.sampler.start()
        self.flush_buf = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self.flush_buf.fill_(1)

    def timed_steps(self, step, steps, warmup):
        """W untimed steps, then K steps between a barrier + synchronize on both sides, CUDA events on the launching
        stream, max over ranks.  Returns (ms_per_step, clocks)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        self.sampler.mark_begin()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        self.barrier()
        clocks = self.sampler.mark_end()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, clocks

    def timed_steps_flushed(self, step, steps, warmup):
        """For L2-resident inputs: an L2 flush (a 256 MiB write) before every step, each step timed on its own."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        self.sampler.mark_begin()
        ms = []
        for _ in range(steps):
            self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        self.barrier()
        clocks = self.sampler.mark_end()
        return self.max_over_ranks(sum(ms) / len(ms)), clocks


def measure_ringmul(env, tag, log2n, steps, warmup, imad_peaks=None):
    """Fused ring mul on 2^log2n elements resident in HBM.  Returns the record and (a, b, out, cfg)."""
    torch, S = env.torch, env.S
    cfg = S.CONFIGS[tag]
    n = 1 << log2n
    a = gen_raw_device(torch, tag, n, 0x5EED ^ (1000 * env.rank + 1), env.dev)
    b = gen_raw_device(torch, tag, n, 0x5EED ^ (1000 * env.rank + 2), env.dev)
    out = torch.empty_like(a)
    step = lambda: cfg.ring_mul_batch(a, b, out=out, ctx=env.ctx)
    in_l2 = 3 * n * ELEM_BYTES[tag] <= 2 * L2_BYTES
    launches0 = env.ctx.kernel_launches
    if in_l2:
        ms, clocks = env.timed_steps_flushed(step, steps, warmup)
    else:
        ms, clocks = env.timed_steps(step, steps, warmup)
    launches = env.ctx.kernel_launches - launches0 - warmup
    # the dominant (only) kernel timed alone with events on the same stream
    ktimes = []
    for _ in range(min(steps, 5)):
        if in_l2:
            env.flush_l2()
        env.ctx.timer_start()
        step()
        ktimes.append(env.ctx.timer_stop())
    k_ms = sum(ktimes) / len(ktimes)
    alg = 3 * ELEM_BYTES[tag] * n
    if tag == "sp" and imad_peaks is not None:
        ach = SP_IMAD_PER_RING_MUL * n / (k_ms / 1e3) / 1e12
        roofline = {"bound": "imad", "achieved": ach, "peak": imad_peaks[2], "unit": "T multiply-add/s",
                    "frac": ach / imad_peaks[2], "traffic": measured_traffic(tag, "ringmul", log2n),
                    "peak_source": "sr_imad_peak measured in this run: IMAD.WIDE.U32.X (32 x 32 + 64 -> 64 with carry), "
                                   "the instruction the multi-limb arithmetic is made of",
                    "imad32_peak_in_run": imad_peaks[0], "imad_wide_peak_in_run": imad_peaks[1],
                    "frac_of_imad32_peak": ach / imad_peaks[0],
                    "algorithmic_multiply_adds_per_launch": SP_IMAD_PER_RING_MUL * n,
                    "hbm_frac": alg / (k_ms / 1e3) / 1e9 / env.hbm_peak, "kernel_ms": k_ms}
    else:
        ach = alg / (k_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": env.hbm_peak, "unit": "GB/s", "frac": ach / env.hbm_peak,
                    "traffic": measured_traffic(tag, "ringmul", log2n), "peak_source": env.peak_src, "kernel_ms": k_ms,
                    "algorithmic_bytes_per_launch": alg}
    rec = {"workload": "%s ring: batched CRT->slot mul->ICRT on 2^%d elements per GPU" % (RING_NAMES[tag], log2n),
           "metric": "ring muls/sec (CRT->mul->ICRT)", "unit": "ring_mul/s", "scaling": "weak",
           "value": n * env.world / (ms / 1e3), "ms": ms, "steps": steps, "gpu_launches": launches,
           "l2": "inputs fit L2: L2 flushed (256 MiB write) before every step, steps timed one by one" if in_l2
                 else "inputs >> L2 (no flush needed)",
           "roofline": roofline, "clocks": clocks}
    return rec, (a, b, out, cfg)


def measure_e2e(env, tag, log2n, a, b, out, cfg, steps):
    """The same metric through sr_ring_mul_batch(SR_HOST) on pinned HOST buffers: H2D and D2H inside the timed region."""
    torch = env.torch
    e2e_n = 1 << log2n
    last = ""
    while e2e_n >= 1 << 10:
        try:
            words = e2e_n * cfg.limbs
            ha = torch.empty(words, dtype=torch.int64).pin_memory()
            hb = torch.empty(words, dtype=torch.int64).pin_memory()
            ho = torch.empty(words, dtype=torch.int64).pin_memory()
            ha.copy_(a[:words])
            hb.copy_(b[:words])
            torch.cuda.synchronize()
            cfg.ring_mul_batch(ha, hb, out=ho, ctx=env.ctx)  # warm-up (allocates the staging ring)
            k = max(1, min(steps, 3))
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(k):
                cfg.ring_mul_batch(ha, hb, out=ho, ctx=env.ctx)  # synchronous: returns when `ho` is complete
            torch.cuda.synchronize()
            dt = env.max_over_ranks((time.perf_counter() - t0) / k)
            if env.rank == 0:
                assert torch.equal(ho[: 64 * cfg.limbs].to(env.dev), out[: 64 * cfg.limbs])
            h2d, d2h = 2 * words * 8, words * 8
            gbs = (h2d + d2h) / dt / 1e9
            return {"value": e2e_n * env.world / dt, "unit": "ring_mul/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "elements_per_gpu": e2e_n, "ms_per_step": dt * 1e3,
                    "bound": "pcie_h2d",
                    "h2d_gbs": h2d / dt / 1e9, "d2h_gbs": d2h / dt / 1e9, "link_gbs_both_directions": gbs,
                    "frac_of_pcie": (h2d / dt / 1e9) / 64.0,
                    "pcie_reference": "PCIe Gen5 x16: 64 GB/s per direction nominal; H2D carries 2/3 of the bytes and "
                                      "bounds the step; both directions run concurrently",
                    "note": "sr_ring_mul_batch(SR_HOST) on pinned host buffers: chunked H2D -> kernel -> D2H pipeline; "
                            "at N > 1 all ranks share the host's memory and PCIe root complexes"}
        except RuntimeError as ex:  # pinned allocation failed: halve
            e2e_n //= 2
            last = str(ex)[:120]
    return {"value": None, "unit": "ring_mul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
            "note": "pinned host allocation failed: " + last}


def measure_commit(env, tag, kappa, log2m, steps, warmup, path="peer", graph=True):
    """kappa x 2^log2m commitment, columns sharded over the ranks of the run (strong scaling).  Rotates over enough
    distinct resident shards that the working set exceeds the L2 (one context per shard: each keeps its row table
    resident, so a step is kernel launches only)."""
    torch, S = env.torch, env.S
    import ctypes
    from stark_rings_b200 import _lib as L
    cfg = S.CONFIGS[tag]
    world, rank = env.world, env.rank
    m = 1 << log2m
    lo, hi = m * rank // world, m * (rank + 1) // world
    m_local = hi - lo
    shard_bytes = (kappa * m_local + m_local + kappa) * ELEM_BYTES[tag]
    nsets = max(1, min(8, -(-4 * L2_BYTES // shard_bytes)))  # working set >= 4 x L2
    ctxs = [env.ctx] + [S.Context(env.local) for _ in range(nsets - 1)]
    for c in ctxs:
        c.set_pipelined(not env.args.no_pipeline)  # resident inputs: the column loop of commitment i + 1 overlaps the tail of commitment i
    sets = []
    for k in range(nsets):
        ctxs[k].use_torch_stream()
        rows = [S.RqNTT(cfg, gen_raw_device(torch, tag, m_local, 0x5EED ^ (1000 * rank + 100 * k + 10 + i), env.dev), ctxs[k])
                for i in range(kappa)]
        v = S.RqNTT(cfg, gen_raw_device(torch, tag, m_local, 0x5EED ^ (1000 * rank + 100 * k + 3), env.dev), ctxs[k])
        sets.append((S.Matrix(rows, ctxs[k]), v))
    limbs = cfg.limbs
    result = torch.empty(kappa * limbs, dtype=torch.int64, device=env.dev)
    gathered = torch.empty(world * kappa * limbs, dtype=torch.int64, device=env.dev)
    peer = None
    if env.dist is not None and path == "peer":
        from stark_rings_b200.dist import PeerCommit
        peer = PeerCommit(cfg, kappa, world, rank, env.ctx, device_epochs=True)
    counter = [0]

    def step():
        k = counter[0] % nsets
        counter[0] += 1
        A, v = sets[k]
        if peer is not None:
            return peer.commit(A, v, out=result, ctx=ctxs[k])
        if env.dist is None:
            ctxs[k].use_torch_stream()
            return A.try_mul_vec(v)
        part = A.partial_mul_vec(v)
        env.dist.all_gather_into_tensor(gathered, part.data)  # raw limbs; an NCCL sum cannot reduce mod p
        if rank == 0:
            env.ctx.check(L.lib.sr_modsum_partials(env.ctx.h, cfg.ring_id, ctypes.c_void_p(gathered.data_ptr()), world,
                                                   kappa, ctypes.c_void_p(result.data_ptr()), L.SR_DEVICE), "modsum")
        return part

    # ---- correctness of the measured path, before any number is reported: one commitment against the oracle ----
    check = "not run"
    try:
        from oracle import c_oracle as C
        import numpy as np
        A, v = sets[0]
        counter[0] = 0
        y = step()
        torch.cuda.synchronize()
        # every rank's shard on the host -> the oracle's partial product -> modular sum via the oracle's own add
        rows_h = [r.data.cpu().numpy().view(np.uint64) for r in A.vals]
        v_h = v.data.cpu().numpy().view(np.uint64)
        sample = min(m_local, 4096)
        w = cfg.limbs
        want_part = C.matvec(RING_NAMES[tag], [r[: sample * w].copy() for r in rows_h], v_h[: sample * w].copy(), threads=kappa)
        A_s = S.Matrix([S.RqNTT(cfg, r.data[: sample * w].clone(), env.ctx) for r in A.vals], env.ctx)
        env.ctx.use_torch_stream()
        got_part = A_s.try_mul_vec(S.RqNTT(cfg, v.data[: sample * w].clone(), env.ctx)).data.cpu().numpy().view(np.uint64)
        ok_local = bool(np.array_equal(got_part, want_part))
        # whole-product consistency: the commit's result == the modular sum of the ranks' own full partial products
        full_part = A.partial_mul_vec(v).data if world > 1 else None
        ok_sum = True
        if world == 1:  # the whole product == the modular sum of the products of its two column halves
            half = (m_local // 2) * w
            halves = []
            for sl in (slice(0, half), slice(half, m_local * w)):
                Ah = S.Matrix([S.RqNTT(cfg, r.data[sl].clone(), env.ctx) for r in A.vals], env.ctx)
                halves.append(Ah.partial_mul_vec(S.RqNTT(cfg, v.data[sl].clone(), env.ctx)).data)
            parts = torch.cat(halves)
            ref = torch.empty(kappa * w, dtype=torch.int64, device=env.dev)
            env.ctx.check(L.lib.sr_modsum_partials(env.ctx.h, cfg.ring_id, ctypes.c_void_p(parts.data_ptr()), 2,
                                                   kappa, ctypes.c_void_p(ref.data_ptr()), L.SR_DEVICE), "modsum")
            ok_sum = bool(torch.equal(ref, y.data))
        if world > 1:
            parts = torch.empty(world * kappa * w, dtype=torch.int64, device=env.dev)
            env.dist.all_gather_into_tensor(parts, full_part)
            if rank == 0:
                ref = torch.empty(kappa * w, dtype=torch.int64, device=env.dev)
                env.ctx.use_torch_stream()
                env.ctx.check(L.lib.sr_modsum_partials(env.ctx.h, cfg.ring_id, ctypes.c_void_p(parts.data_ptr()), world,
                                                       kappa, ctypes.c_void_p(ref.data_ptr()), L.SR_DEVICE), "modsum")
                ok_sum = bool(torch.equal(ref, result))
        ok = env.max_over_ranks(0.0 if (ok_local and ok_sum) else 1.0) == 0.0
        if not ok:
            raise SystemExit("bench.py: the commitment does not match the oracle (local %s, sum %s)" % (ok_local, ok_sum))
        check = "one commitment verified: %d-column sample of every rank's shard against the C oracle, and the " \
                "delivered result against the modular sum of the ranks' partial products" % sample
    except ImportError:
        check = "oracle not importable"
    counter[0] = 0

    # ---- timing: the nsets commitments captured in one CUDA graph when the path allows it ----
    cuda_graph = False
    unit_steps = nsets
    one = lambda: [step() for _ in range(nsets)]
    for _ in range(2):
        one()
    env.barrier()
    if graph and (world == 1 or peer is not None):
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                one()
            one = g.replay
            cuda_graph = True
        except Exception as ex:
            cuda_graph = "capture failed: %s" % str(ex)[:80]
            torch.cuda.synchronize()
        for c in ctxs:
            c.use_torch_stream()
    reps = max(1, -(-steps // unit_steps))
    launches0 = sum(c.kernel_launches for c in ctxs)
    ms_rep, clocks = env.timed_steps(one, reps, max(1, -(-warmup // unit_steps)))
    launches = sum(c.kernel_launches for c in ctxs) - launches0
    ms = ms_rep / unit_steps
    if peer is not None:
        peer.check()  # raises if any in-kernel wait timed out: no number from a broken exchange
    # the product kernel alone (no mailbox): what the step costs without the exchange.  Captured in a graph as well,
    # so that the host's launch path is not what is measured.
    outs = [torch.empty(kappa * limbs, dtype=torch.int64, device=env.dev) for _ in range(nsets)]

    def products():
        for k in range(nsets):
            A, v = sets[k]
            pv, nv = ctypes.c_void_p(v.data.data_ptr()), v.data.numel()
            ptrs = (ctypes.c_void_p * kappa)(*[r.data.data_ptr() for r in A.vals])
            ctxs[k].use_torch_stream()
            ctxs[k].check(L.lib.sr_matvec_partial(ctxs[k].h, cfg.ring_id, ptrs, kappa, m_local, pv, nv,
                                                  ctypes.c_void_p(outs[k].data_ptr()), L.SR_DEVICE), "sr_matvec_partial")
    products()
    torch.cuda.synchronize()
    run_products = products
    if graph:
        try:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                products()
            run_products = g2.replay
        except Exception:
            torch.cuda.synchronize()
        for c in ctxs:
            c.use_torch_stream()
    k_rep, _ = env.timed_steps(run_products, max(3, reps), 2)
    kernel_ms = k_rep / nsets
    ach = shard_bytes / (ms / 1e3) / 1e9
    rec = {"workload": "%s commit: %d x 2^%d ring matrix x vector, columns sharded over %d GPU(s)" % (
               RING_NAMES[tag], kappa, log2m, world),
           "metric": "commits/sec (kappa x m ring matrix x vector)", "unit": "commit/s", "scaling": "strong",
           "value": 1e3 / ms, "ms": ms, "steps": reps * unit_steps,
           "gpu_launches": launches if launches else (reps * unit_steps * ((kappa + 3) // 4)),
           "kernels_per_commit_per_rank": (kappa + 3) // 4,
           "exchange": "none (1 GPU)" if env.dist is None else (
               "NVLink peer-memory mailbox: the tail of each rank's product kernel stores its partial rows into the "
               "root's HBM and publishes a flag; the tail of the root's kernel acquires the flags and adds mod p"
               if peer else "NCCL all_gather of raw limbs + rank-0 modular sum"),
           "cuda_graph": cuda_graph,
           "pipelined": "programmatic dependent launch: the column loop of a commitment starts while the previous one "
                        "is in its tail (cross-CTA reduction + NVLink hand-off); chunks are drawn from a device-wide "
                        "counter so a CTA that starts late draws fewer",
           "l2": "rotating over %d distinct resident shards per rank (%.0f MB each): working set >= 4 x L2" % (
               nsets, shard_bytes / 1e6),
           "product_kernel_us": kernel_ms * 1e3, "step_us": ms * 1e3,
           "exchange_us": max(0.0, (ms - kernel_ms) * 1e3),
           "verified": check,
           "roofline": {"bound": "hbm", "achieved": ach, "peak": env.hbm_peak, "unit": "GB/s per GPU",
                        "frac": ach / env.hbm_peak, "traffic": measured_traffic(tag, "commit", log2m),
                        "peak_source": env.peak_src, "algorithmic_bytes_per_launch_per_gpu": shard_bytes,
                        "product_kernel_frac": shard_bytes / (kernel_ms / 1e3) / 1e9 / env.hbm_peak},
           "clocks": clocks}
    if peer is not None:
        peer.close()
    env.ctx.set_pipelined(False)
    return rec


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
    ap.add_argument("--cpu-sample", type=int, default=0, help="elements in the CPU arm's sample (0: the full size if it fits)")
    ap.add_argument("--e2e-log2n", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (the line's `extra` list)")
    ap.add_argument("--commit-path", default="peer", choices=["peer", "nccl"],
                    help="commit workload, N > 1: how the per-rank partials reach rank 0")
    ap.add_argument("--no-pipeline", action="store_true", help="commit workload: no programmatic dependent launch")
    ap.add_argument("--no-graph", action="store_true", help="commit workload: do not capture the steps in a CUDA graph")
    args = ap.parse_args()
    default_run = args.ring is None and args.log2n is None and args.workload == "ringmul"
    if args.ring is None:
        args.ring = "gl" if args.workload == "commit" else "bb"
    if args.log2n is None:
        args.log2n = 20 if args.workload == "commit" else {"bb": 24, "gl": 24, "sp": 20}[args.ring]
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    env = Env(args)
    torch = env.torch
    tag = args.ring
    metric, unit, config = make_config(args, env.world)
    imad = None
    if tag == "sp" or (default_run and not args.no_extra):
        imad = env.ctx.imad_peak()

    if args.workload == "ringmul":
        head, (a, b, out, cfg) = measure_ringmul(env, tag, args.log2n, args.steps, args.warmup, imad)
        e2e = None
        if not args.no_e2e:
            e2e = measure_e2e(env, tag, args.e2e_log2n if args.e2e_log2n is not None else args.log2n, a, b, out, cfg,
                              args.steps)
        del a, b, out
        torch.cuda.empty_cache()
        strong = False
    else:
        head = measure_commit(env, tag, args.kappa, args.log2n, args.steps, args.warmup, args.commit_path,
                              not args.no_graph)
        for k in ("exchange", "cuda_graph", "l2", "product_kernel_us", "step_us", "exchange_us", "verified",
                  "kernels_per_commit_per_rank"):
            config[k] = head[k]
        e2e = None
        strong = True

    cpu = None
    if env.rank == 0 and not args.no_cpu and env.world == 1 and args.workload == "ringmul":
        cores = os.cpu_count() or 1
        sample = min(1 << args.log2n, args.cpu_sample or (1 << 20))
        rate, dt, kind = cpu_reference_rate(tag, sample, cores)
        rate1, dt1, _ = cpu_reference_rate(tag, max(1024, sample // 16), 1)
        cpu = {"value": rate, "unit": unit, "cores": cores, "kind": "port",
               "sample": "%d elements (%.1f s wall on %d threads), %s build of oracle/sr_oracle.c" % (sample, dt, cores, kind),
               "value_1thread": rate1}

    extra = []
    if default_run and not args.no_extra:
        # the other BASELINE configs, measured in this process
        for etag, elog in (("gl", 16), ("gl", 24), ("sp", 20)):
            rec, bufs = measure_ringmul(env, etag, elog, max(args.steps, 10), args.warmup, imad)
            del bufs
            torch.cuda.empty_cache()
            extra.append(rec)
        # (a commitment is 35-200 us: a few hundred of them, after a few dozen untimed ones, so that the start-up skew of
        # the ranks is not what is measured)
        extra.append(measure_commit(env, "gl", 4, 20, max(args.steps, 400), max(args.warmup, 40), args.commit_path,
                                    not args.no_graph))
    env.sampler.stop()

    if env.rank == 0:
        line = {
            "metric": metric, "value": head["value"], "unit": unit, "n_gpus": env.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": config,
            "clocks": head["clocks"], "e2e": e2e, "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "cpu_baseline": cpu,
        }
        if extra:
            line["extra"] = extra
        emit(line)
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
