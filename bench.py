#!/usr/bin/env python
"""bench.py -- headline benchmark of the flomat dense-FP64 hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 -- the configuration BASELINE.json's metric is quoted on.  One "step" = C = (times A B) followed by E = (expm X) at n = 8192,
A, B ~ N(0,1), X ~ N(0,1)/sqrt(n).  value = (2 n^3 + W_min) / t in GFLOP/s, W_min = (5+s) 2n^3 + 8/3 n^3 the fixed
algorithmic work of one expm, t from CUDA events over exactly K steps.  Inputs resident in HBM when the timed region starts.

N > 1 (torchrun, one rank per GPU) -- the workload that SHARDS, total work fixed ("scaling": "strong"): BASELINE config 4,
`times` at n = 16384 with B and C in column blocks and the all-gather of C fused into the GEMM epilogue over NVLink peer memory,
followed by config 5, 4096 independent expm of 256 x 256 matrices by item blocks (no collective).  value = total algorithmic GFLOP/s
of the job, time = max over ranks.  The results are VERIFIED after the timed region and the run exits non-zero if a check fails.
The same job on ONE GPU is timed in the N = 1 run under "sharded_job_n1" (T_1 of the strong-scaling efficiency T_1 / (N T_N));
the old weak-scaling number (one n = 8192 step per GPU) stays under "replicas".

e2e = the same step through the host API with HOST buffers: uploads from pinned memory, compute, downloads, every step.
cpu_baseline / --impl reference = the reference's own CPU path: its call sequence (oracle/flomat_oracle.py) on OpenBLAS with all host
threads.  At N = 1 the reference arm runs K REAL full steps (times n = 8192 + expm n = 8192: 11+s dgemms + dgesv, ~25 s each on 16
cores) whenever K of them fit --ref-budget-s (640 s; the driver's tightest per-run limit is 870 s); otherwise, and for the sharded
N > 1 job, its step is a BOUNDED SAMPLE at the full sizes: every distinct foreign call of the step executed once on the full-size
operands, step time = the sum over the call sequence the reference executes (multiplicities in `cpu_baseline.sample`).  The GPU
arm's own `cpu_baseline` at N = 1 runs ONE real full step and prints the sampled estimate beside it (measured: within 2-3 %).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP64 GFLOP/s for times (dgemm) and expm at n=8192; % of FP64 tensor peak"


def fp64_peak() -> dict:
    """The FP64 tensor (DMMA) peak: ONE tracked file with its clock record (MEASURED_PEAKS.json has no FP64 entry)."""
    with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
        return json.load(f)


def dgemm_traffic() -> dict:
    """DRAM bytes per launch of the dominant kernel, imported from the committed `ncu --set full` capture (not measured by this run)."""
    path = os.path.join(ROOT, "profiles", "dgemm_n8192_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except OSError:
        return {"dram_bytes_per_launch": None, "source": "missing " + path}


def expm_flops_min(n: int, s: float) -> float:
    return (5 + max(s, 0)) * 2.0 * n ** 3 + (8.0 / 3.0) * n ** 3


def expm_flops_ref(n: int, s: float) -> float:
    return (11 + max(s, 0)) * 2.0 * n ** 3 + (8.0 / 3.0) * n ** 3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the busier half of the samples = clocks under load
        busy = sorted(s for s, p in zip(sm, power) if p >= 0.5 * max(power)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# workload descriptions (the same dict is printed by both arms)
# ------------------------------------------------------------------------------------------------
def workload_config(args, gpus: int) -> dict:
    if gpus == 1:
        n = args.n
        return {"workload": f"flomat (times A B) + (expm X), n={n}, FP64, A,B~N(0,1), X~N(0,1)/sqrt(n)", "n": n,
                "expm": "Pade[8/8] scaling+squaring (expm.rkt:203): W_min = (5+s) 2n^3 + 8/3 n^3 counted",
                "l2": "inputs (3 x 512 MiB at n=8192) exceed the 126 MB L2; no flush needed", "parallelism": "1 GPU"}
    return {"workload": f"sharded job: (times A B) n={args.colshard_n} column-sharded with the all-gather of C fused into the GEMM epilogue "
                        f"(BASELINE config 4) + {args.batch_items} independent expm of {args.batch_n}x{args.batch_n} item-sharded (config 5); "
                        "FP64, A,B~N(0,1), items~N(0,1)/16; total work fixed",
            "n": args.colshard_n, "items": args.batch_items, "item_n": args.batch_n,
            "l2": "A (2 GiB) and every rank's block of B/C exceed the 126 MB L2; no flush needed",
            "parallelism": f"column blocks of B/C x{gpus} + item blocks x{gpus}"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's call sequence on OpenBLAS (oracle) -- test infrastructure used as the reported baseline
# ------------------------------------------------------------------------------------------------
def _cpu_flomat():
    from oracle.flomat_oracle import Flomat, default_provider

    prov = default_provider()
    if hasattr(prov, "_set"):  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses all the host threads it can
        prov._set(os.cpu_count() or 1)
    return Flomat(prov), prov


def _t(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


class CpuMetricStep:
    """N = 1 config on the CPU.  full(): one real step.  sampled(): every distinct foreign call once, summed over the sequence
    expm.rkt executes: (12+s) x [zero C + dgemm], 11 x [copy + n daxpy], 9 x [c .* I], 2 x [eye + zeros], ./, dlange, mldivide."""

    def __init__(self, n: int):
        import numpy as np

        self.np, self.n = np, n
        self.F, self.prov = _cpu_flomat()
        rng = np.random.default_rng(6001)
        self.a = np.asfortranarray(rng.standard_normal((n, n)))
        self.b = np.asfortranarray(rng.standard_normal((n, n)))
        self.x = np.asfortranarray(rng.standard_normal((n, n)) / math.sqrt(n))
        self.s = self.F.scaling_exponent(self.x)
        self.flops = 2.0 * n ** 3 + expm_flops_min(n, self.s)
        self.F.times(self.a[:256, :256].copy(order="F"), self.b[:256, :256].copy(order="F"))  # spin the BLAS threads up

    def full(self) -> float:
        t0 = time.perf_counter()
        self.F.times(self.a, self.b)
        self.F.expm(self.x)
        return time.perf_counter() - t0

    def sampled(self) -> tuple[float, dict]:
        F, n, s = self.F, self.n, max(int(self.s), 0)
        xs = F.dot_divide(self.x, 2.0 ** self.s)
        t_gemm, aa = _t(lambda: F.times(xs, xs))              # make-flomat zeros + cblas_dgemm alpha = beta = 1 (flomat.rkt:905-914)
        t_eye, ident = _t(lambda: F.eye(n))
        t_zero, _ = _t(lambda: F.make(n, n))
        t_ci, ci = _t(lambda: F.dot_times(0.5, ident))        # (.* c I)
        t_plus, even = _t(lambda: F.plus(ci, aa))             # copy + n daxpy (flomat.rkt:805-819)
        t_div, _ = _t(lambda: F.dot_divide(self.x, 2.0 ** self.s))
        t_norm, _ = _t(lambda: F.norm(self.x, 1))
        den = F.minus(even, xs)
        t_solve, _ = _t(lambda: F.mldivide(den, even))        # 2 copies + dgesv with n right-hand sides
        mult = {"gemm": 1 + 11 + s, "plus": 11, "c.*I": 9, "eye+zeros": 2, "./": 1, "dlange": 1, "mldivide": 1}
        t = (mult["gemm"] * t_gemm + 11 * t_plus + 9 * t_ci + 2 * (t_eye + t_zero) + t_div + t_norm + t_solve)
        return t, {"seconds_once": {"gemm": t_gemm, "plus": t_plus, "c.*I": t_ci, "eye": t_eye, "zeros": t_zero, "./": t_div,
                                    "dlange": t_norm, "mldivide": t_solve}, "multiplicity": mult}

    def describe(self, t: float) -> str:
        return (f"reference call sequence (flomat.rkt:888, expm.rkt:193-210) on {self.prov.name}, times n={self.n} + expm n={self.n} "
                f"(s={int(self.s)}): each distinct foreign call executed once per step on the full-size operands, step time = sum over the "
                f"{12 + max(int(self.s), 0)} dgemm, 11 plus/minus, 9 c.*I, ./, dlange and mldivide the reference executes; W_min counted; "
                f"{t:.1f} s per step")


class CpuShardedStep:
    """N > 1 config on the CPU (the CPU does the whole job): times n x n sampled by ONE column block of n / 8 columns (x 8) and the
    batch of expm by `sample_items` whole items (x items / sample_items)."""

    def __init__(self, n: int, items: int, n_item: int, sample_items: int = 32, blocks: int = 8):
        import numpy as np

        self.np, self.n, self.items, self.n_item = np, n, items, n_item
        self.F, self.prov = _cpu_flomat()
        self.blocks = blocks if n >= 8 * blocks else 1
        self.k = max(1, min(sample_items, items))
        rng = np.random.default_rng(4001)
        self.a = np.asfortranarray(rng.standard_normal((n, n)))
        self.b = np.asfortranarray(rng.standard_normal((n, n // self.blocks)))
        rng = np.random.default_rng(5001)
        self.xs = [np.asfortranarray(rng.standard_normal((n_item, n_item)) / 16.0) for _ in range(self.k)]
        w_item = sum(expm_flops_min(n_item, self.F.scaling_exponent(x)) for x in self.xs) / self.k
        self.flops = 2.0 * n ** 3 + items * w_item
        self.F.times(self.xs[0], self.xs[0])

    def sampled(self) -> tuple[float, dict]:
        t_blk, _ = _t(lambda: self.F.times(self.a, self.b))
        # small products do not scale to every core: give the reference the better of all threads and 8 threads for the items
        t_items, _ = _t(lambda: [self.F.expm(x) for x in self.xs])
        all_threads = self.prov.threads()
        if hasattr(self.prov, "_set") and all_threads > 8:
            self.prov._set(8)
            t8, _ = _t(lambda: [self.F.expm(x) for x in self.xs])
            self.prov._set(all_threads)
            t_items = min(t_items, t8)
        t = self.blocks * t_blk + (self.items / self.k) * t_items
        return t, {"seconds_once": {"times_column_block": t_blk, f"expm_{self.k}_items": t_items},
                   "multiplicity": {"times_column_block": self.blocks, f"expm_{self.k}_items": self.items / self.k}}

    def describe(self, t: float) -> str:
        return (f"reference call sequence on {self.prov.name}: times n={self.n} sampled by one block of {self.n // self.blocks} columns of B/C "
                f"(x{self.blocks}), {self.items} expm of {self.n_item}x{self.n_item} sampled by {self.k} whole items (x{self.items / self.k:g}); "
                f"W_min counted; {t:.1f} s per step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    job = CpuMetricStep(args.n) if world == 1 else CpuShardedStep(args.colshard_n, args.batch_items, args.batch_n)
    t_est, detail = job.sampled()  # doubles as the warm-up (BLAS threads, page faults); more warm-up buys nothing on the CPU
    # N = 1: K REAL full steps whenever they fit the budget (a real step is ~25 s on 16 cores: K = 20 fits the driver's per-run limit
    # of 870 s); otherwise, and for the sharded job (thousands of identical independent pieces), the bounded sample at full size
    real = world == 1 and args.steps * t_est <= args.ref_budget_s
    times = []
    for _ in range(args.steps):
        if real:
            times.append(job.full())
        else:
            t, detail = job.sampled()
            times.append(t)
    t = sum(times) / len(times)
    value = job.flops / t / 1e9
    sample = (f"{args.steps} REAL full steps of the reference call sequence (flomat.rkt:888, expm.rkt:193-210) on {job.prov.name}: times n={args.n} + "
              f"expm n={args.n} (s={int(job.s)}, {11 + max(int(job.s), 0)} dgemms + dgesv executed, W_min counted), {t:.1f} s per step"
              if real else job.describe(t))
    cb = {"value": value, "unit": "GFLOP/s", "cores": job.prov.threads(), "kind": "port", "sample": sample, "real_full_steps": bool(real),
          "warmup_steps_run": 1, "sampled_estimate_s": t_est, "detail": detail}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Gpu:
    def __init__(self):
        import torch
        import torch.distributed as dist

        import sci_b200 as fm

        self.torch, self.dist, self.fm = torch, dist, fm
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        # one process per GPU: keep this rank's host threads and pinned buffers on the GPU's NUMA node (before anything allocates)
        self.affinity = None
        self.cpus_before = os.sched_getaffinity(0)
        if self.world > 1 and not os.environ.get("FLOMAT_BENCH_NO_AFFINITY"):
            from sci_b200.multigpu import bind_host_near_device

            self.affinity = bind_host_near_device(self.local)
        torch.cuda.set_device(self.local)
        fm.init(self.local)
        self.lib = fm.load()
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def best_ms(self, fn, reps=3):
        fn(); self.torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a0, a1 = self.ev(), self.ev()
            a0.record(); fn(); a1.record(); self.torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(a1))
        return best

    def dev_randn(self, rows, cols, seed, scale=1.0):
        g = self.torch.Generator(device=self.dev); g.manual_seed(seed)
        t = self.torch.randn(rows * cols, dtype=self.torch.float64, device=self.dev, generator=g) * scale
        out = self.fm.Flomat.alloc(rows, cols)
        self.fm._lib.check(self.lib.fm_copy2d(out.a, rows, t.data_ptr(), rows, rows, cols), "fm_copy2d")
        self.fm.sync()
        return out


def timed_steps(G: Gpu, step, steps: int, sampler: ClockSampler | None):
    """Exactly `steps` steps between barriers, CUDA events on the launching (default) stream; returns (ms per step max over ranks,
    kernel launches of this rank, clocks)."""
    launches0 = G.lib.fm_kernel_launches()
    if sampler:
        sampler.start()
    G.barrier()
    e0, e1 = G.ev(), G.ev()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    G.barrier()
    clocks = sampler.stop() if sampler else None
    launches = G.lib.fm_kernel_launches() - launches0
    (ms,) = G.max_over_ranks(e0.elapsed_time(e1))
    return ms / steps, int(launches), clocks


def timed_e2e(G: Gpu, step, steps: int):
    step()
    G.barrier()
    t0 = time.perf_counter()
    g0, g1 = G.ev(), G.ev()
    g0.record()
    for _ in range(steps):
        step()
    g1.record()
    G.barrier()
    (ms,) = G.max_over_ranks(max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3))
    return ms / steps


def metric_step_n8192(G: Gpu, args, steps: int, warmup: int, with_e2e: bool, sampler: ClockSampler | None, seed_offset: int = 0):
    """The metric's configuration on THIS rank's GPU: times + expm at n = args.n."""
    import numpy as np

    fm, lib, torch, n = G.fm, G.lib, G.torch, args.n
    rng = np.random.default_rng(6001 + 17 * seed_offset)
    pin = (lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()) if with_e2e else (lambda shape: np.empty(shape))
    hA, hB, hX = pin((n, n)), pin((n, n)), pin((n, n))  # C-order host buffers hold the column-major matrices' bytes
    for h, scale in ((hA, 1.0), (hB, 1.0), (hX, 1.0 / math.sqrt(n))):
        h[...] = rng.standard_normal((n, n)) * scale
    hC, hE = (pin((n, n)), pin((n, n))) if with_e2e else (None, None)

    def upload(h):
        out = fm.Flomat.alloc(n, n)
        fm._lib.check(lib.fm_upload(out.a, n, h.ctypes.data, n, n, n), "fm_upload")
        return out

    A, B, X = upload(hA), upload(hB), upload(hX)
    fm.sync()
    C = fm.Flomat.alloc(n, n)
    mode = fm.EXPM_MIN_PRODUCTS if args.expm_mode == 1 else fm.EXPM_REFERENCE_ORDER
    s_val = 0.0
    for _ in range(warmup):
        fm.flomat_times(A, B, C, 1.0, 0.0)
        _, s_val = fm.expm(X, mode, return_s=True)
    G.barrier()
    flops_step = 2.0 * n ** 3 + expm_flops_min(n, s_val)
    et = [(G.ev(), G.ev()) for _ in range(steps)]
    k = [0]

    def step():
        a, b = et[k[0]]
        k[0] += 1
        a.record()
        fm.flomat_times(A, B, C, 1.0, 0.0)
        b.record()
        fm.expm(X, mode)

    ms_step, launches, clocks = timed_steps(G, step, steps, sampler)
    (ms_times,) = G.max_over_ranks(sum(a.elapsed_time(b) for a, b in et) / steps)
    out = {"ms_step": ms_step, "ms_times": ms_times, "s": s_val, "flops_step": flops_step, "launches": launches, "clocks": clocks,
           "mode": "5+s" if args.expm_mode == 1 else "7+s", "host": (hA, hB, hX)}
    if with_e2e:
        def e2e_step():
            # `times` on HOST operands through fm_dgemm_host (what cblas_dgemm does with the reference's host matrices): A goes up, then
            # B and C move in column blocks beside the product.  X goes up behind B's blocks while `times` runs; `expm` delivers its
            # result to the host buffer block by block during the last squaring.
            x_ = fm.Flomat.alloc(n, n)
            fm.times_host(hA.T, hB.T, hC.T, 1.0, 0.0, wait=False)
            x_.upload_async(hX.reshape(-1))
            fm.expm_to_host(x_, hE.T, mode)
            fm.transfer_sync()

        e2e_steps = max(1, min(steps, 3))
        out["e2e_ms"] = timed_e2e(G, e2e_step, e2e_steps)
        out["e2e_steps"] = e2e_steps
        out["e2e_bytes"] = (3 * n * n * 8, 2 * n * n * 8)
        out["e2e_api"] = ("per step, pinned host buffers: times_host (fm_dgemm_host: A up, B/C in column blocks beside the product) -> X "
                          "upload_async beside it -> expm with a host result buffer (last squaring in column blocks, downloads beside it)")
    del A, B, C, X
    return out


def sharded_job(G: Gpu, args, steps: int, warmup: int, with_e2e: bool, sampler: ClockSampler | None, compare_nccl: bool):
    """BASELINE configs 4 + 5 sharded over all ranks; verified after timing."""
    from sci_b200.multigpu import ShardedJob

    fm = G.fm
    mode = fm.EXPM_MIN_PRODUCTS if args.expm_mode == 1 else fm.EXPM_REFERENCE_ORDER
    job = ShardedJob(args.colshard_n, args.batch_items, args.batch_n, G.world, G.rank, mode)
    for _ in range(warmup):
        job.step()
    G.barrier()
    (w_total,) = G.sum_over_ranks(job.flops(expm_flops_min))
    ms_step, launches, clocks = timed_steps(G, job.step, steps, sampler)
    checks = job.verify()
    out = {"ms_step": ms_step, "flops_step": w_total, "launches": launches, "clocks": clocks, "checks": checks,
           "columns_per_gpu": job.cs.nloc, "items_per_gpu": job.nitems, "s_max": float(job.s_out.max())}
    # the two halves on their own (same inputs, not part of `value`)
    ms_t = job.cs._time(job.cs.step_fused, max(2, min(steps, 3)), 1)
    out["times"] = {"n": args.colshard_n, "fused_ms": ms_t, "tflops_total": 2.0 * args.colshard_n ** 3 / ms_t / 1e9,
                    "tflops_per_gpu": 2.0 * args.colshard_n ** 3 / ms_t / 1e9 / G.world}
    if compare_nccl and G.world > 1:
        job.cs.c_tensor().zero_()
        ms_n = job.cs._time(job.cs.step_nccl, max(2, min(steps, 3)), 1)
        out["times"].update({"nccl_allgather_ms": ms_n, "nccl_ok": job.cs.verify()})
        checks["all_ranks_ok"] = checks["all_ranks_ok"] and out["times"]["nccl_ok"]
    if job.nitems:
        ms_e = G.best_ms(lambda: fm.expm_batched(job.src.data_ptr(), job.n_item, job.nitems, job.dst.data_ptr(), mode, None))
        (ms_e,) = G.max_over_ranks(ms_e)
        w_e = w_total - 2.0 * args.colshard_n ** 3
        out["expm_batched"] = {"n": args.batch_n, "items_total": args.batch_items, "items_per_gpu": job.nitems, "ms": ms_e,
                               "tflops_wmin_total": w_e / ms_e / 1e9, "tflops_wmin_per_gpu": w_e / ms_e / 1e9 / G.world}
    if with_e2e:
        e2e_steps = max(1, min(steps, 3))
        out["e2e_ms"] = timed_e2e(G, job.e2e_step, e2e_steps)
        out["e2e_steps"] = e2e_steps
        b = job.e2e_bytes()
        h2d, d2h = G.sum_over_ranks(b["h2d"], b["d2h"])
        out["e2e_bytes"] = (int(h2d), int(d2h))
        out["e2e_api"] = ("per step and rank, pinned host buffers: own column slice of A up (fm_upload_async) -> A all-gathered over NVLink "
                          "(NCCL, in place) -> own block of B / C in sub-blocks: upload i+1 | fused column-sharded GEMM i | download i-1; own "
                          "items up beside the GEMMs -> fm_expm_batched -> results down; bytes are totals over all ranks")
    job.close()
    return out


def single_process_mg(G: Gpu, args) -> dict:
    """The single-process entry points (fm_mg_*: ONE process, ONE calling thread, every device of the box -- the form `flomat*` can
    use, flomat.rkt:905-914) driven from rank 0 while the other ranks wait on the host: config 4 with device operands on device 0,
    with pinned and with pageable host operands, and config 5 with device operands."""
    import numpy as np

    fm, lib, torch, n = G.fm, G.lib, G.torch, args.colshard_n
    os.sched_setaffinity(0, G.cpus_before)  # this process now drives every device: its worker threads must not inherit rank 0's NUMA binding
    ndev = fm.mg_init(G.world)
    out = {"devices": ndev, "n": n}
    A, B = G.dev_randn(n, n, 4001), G.dev_randn(n, n, 4002)
    C = fm.Flomat.alloc(n, n)
    ms = G.best_ms(lambda: fm.times_mg(A, B, C, 1.0, 0.0))
    ok = fm.norm(fm.minus(C, fm.times(A, B)), "max") == 0.0  # bit for bit what one device computes
    out["dgemm_device_operands"] = {"ms": ms, "tflops_total": 2.0 * n ** 3 / ms / 1e9, "equals_single_device": bool(ok),
                                    "includes": "A and the blocks of B travel from device 0 over NVLink inside the call; C is assembled on device 0 by the peer-store epilogue"}
    hA, hB = A.to_numpy(), B.to_numpy()  # pageable
    hC = np.empty((n, n), order="F")
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter(); fm.times_host_mg(hA, hB, hC, 1.0, 0.0); best = min(best, (time.perf_counter() - t0) * 1e3)
    samp = np.random.default_rng(1).integers(0, n, 64)
    want = C.to_numpy()
    ok_h = bool(np.array_equal(hC[:, samp], want[:, samp]))
    out["dgemm_host_pageable"] = {"ms": best, "tflops_end_to_end": 2.0 * n ** 3 / best / 1e9, "equals_device_result": ok_h,
                                  "h2d_bytes": 2 * n * n * 8, "d2h_bytes": n * n * 8}
    del want
    pin = lambda: torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy().T  # noqa: E731  (F-contiguous views)
    pA, pB, pC = pin(), pin(), pin()
    pA[...] = hA; pB[...] = hB
    del hA, hB, hC
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); fm.times_host_mg(pA, pB, pC, 1.0, 0.0); best = min(best, (time.perf_counter() - t0) * 1e3)
    out["dgemm_host_pinned"] = {"ms": best, "tflops_end_to_end": 2.0 * n ** 3 / best / 1e9}
    del pA, pB, pC, A, B, C
    ni, items = args.batch_n, args.batch_items
    g = torch.Generator(device=G.dev); g.manual_seed(5001)
    src = torch.randn(ni * ni * items, dtype=torch.float64, device=G.dev, generator=g) / 16.0
    dst = torch.empty_like(src)
    s_out = np.zeros(items)
    mode = fm.EXPM_MIN_PRODUCTS if args.expm_mode == 1 else fm.EXPM_REFERENCE_ORDER
    ms = G.best_ms(lambda: fm.expm_batched_mg(src.data_ptr(), ni, items, dst.data_ptr(), mode, s_out))
    w = float(sum(expm_flops_min(ni, sv) for sv in s_out))
    out["expm_batched_device_operands"] = {"items": items, "n": ni, "ms": ms, "tflops_wmin_total": w / ms / 1e9}
    del src, dst
    # the metric's expm (n = 8192, ONE matrix) over every device: fm_mg_expm against fm_expm on device 0, same input
    ne = args.n
    X = G.dev_randn(ne, ne, 6003, 1.0 / math.sqrt(ne))
    t_one = G.best_ms(lambda: fm.expm(X, mode), 2)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); E, s_e = fm.expm_mg(X, mode=mode, return_s=True); best = min(best, (time.perf_counter() - t0) * 1e3)
    E1 = fm.expm(X, mode)
    diff = fm.norm(fm.minus(E, E1), 1) / fm.norm(E1, 1)
    gate = 64.0 * ne * 2.0 ** -52
    out["expm_one_matrix"] = {"n": ne, "s": s_e, "ms": best, "one_device_ms": t_one, "speedup": t_one / best,
                              "tflops_wmin_total": expm_flops_min(ne, s_e) / best / 1e9, "rel_diff_vs_one_device": diff, "gate_64_n_eps": gate,
                              "what": "fm_mg_expm: products column-sharded with the fused all-gather epilogue, LU replicated, solves sharded by right-hand sides; wall clock of the blocking call"}
    out["ok"] = bool(ok and ok_h and diff <= gate)
    return out


def run_gpu(args):
    import numpy as np

    G = Gpu()
    fm, lib, torch, world, rank = G.fm, G.lib, G.torch, G.world, G.rank
    peak = fp64_peak()
    PEAK = float(peak["fp64_dmma_tflops"])
    warm = max(args.warmup, 3)
    sampler = ClockSampler(G.local) if rank == 0 else None
    line = {"metric": METRIC, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warm, "higher_is_better": True,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world)}
    failed = []

    if world == 1:
        m = metric_step_n8192(G, args, args.steps, warm, True, sampler)
        n = args.n
        value = m["flops_step"] / (m["ms_step"] * 1e-3) / 1e9
        times_tflops = 2.0 * n ** 3 / (m["ms_times"] * 1e-3) / 1e12
        expm_ms = m["ms_step"] - m["ms_times"]
        expm_gflops = expm_flops_min(n, m["s"]) / (expm_ms * 1e-3) / 1e9
        traffic = dgemm_traffic()
        line.update({
            "value": value, "ms_per_step": m["ms_step"], "scaling": "weak",
            "pct_of_fp64_tensor_peak": 100.0 * value / 1e3 / PEAK,
            "times": {"n": n, "ms": m["ms_times"], "tflops": times_tflops, "pct_of_peak": 100.0 * times_tflops / PEAK},
            "expm": {"n": n, "s": m["s"], "ms": expm_ms, "gflops_wmin": expm_gflops, "gflops_wref": expm_flops_ref(n, m["s"]) / (expm_ms * 1e-3) / 1e9,
                     "pct_of_peak_wmin": 100.0 * expm_gflops / 1e3 / PEAK, "mode": m["mode"]},
            "roofline": {"bound": "tensor", "kernel": "dgemm_tma_kernel (DMMA.8x8x4 + TMA)", "achieved": times_tflops, "peak": PEAK,
                         "unit": "TFLOP/s", "frac": times_tflops / PEAK, "flops_per_launch": 2.0 * n ** 3,
                         "traffic": traffic.get("dram_bytes_per_launch"),
                         "traffic_source": "IMPORTED, not measured by this run: " + str(traffic.get("source")),
                         "algorithmic_bytes_per_launch": 3.0 * n * n * 8,
                         "peak_source": f"profiles/fp64_peak.json ({peak['how']}; {peak['sm_mhz']} MHz)"},
            "e2e": {"value": m["flops_step"] / (m["e2e_ms"] * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": m["e2e_bytes"][0],
                    "d2h_bytes_per_step": m["e2e_bytes"][1], "ms_per_step": m["e2e_ms"], "steps": m["e2e_steps"], "api": m["e2e_api"]},
            "gpu_launches": m["launches"], "clocks": m["clocks"]})
        hA, hB, hX = m.pop("host")
    else:
        j = sharded_job(G, args, args.steps, warm, True, sampler, compare_nccl=True)
        value = j["flops_step"] / (j["ms_step"] * 1e-3) / 1e9
        if not j["checks"]["all_ranks_ok"]:
            failed.append(f"sharded job verification failed: {j['checks']}")
        line.update({
            "value": value, "ms_per_step": j["ms_step"], "scaling": "strong",
            "pct_of_fp64_tensor_peak_per_gpu": 100.0 * value / world / 1e3 / PEAK,
            "verified": j["checks"], "columns_per_gpu": j["columns_per_gpu"], "items_per_gpu": j["items_per_gpu"],
            "times_colshard": j["times"], "expm_batched": j.get("expm_batched"),
            "roofline": {"bound": "tensor", "kernel": "dgemm_tma_kernel (DMMA.8x8x4 + TMA), column block with peer-store epilogue",
                         "achieved": j["times"]["tflops_per_gpu"], "peak": PEAK, "unit": "TFLOP/s", "frac": j["times"]["tflops_per_gpu"] / PEAK,
                         "flops_per_launch": 2.0 * args.colshard_n ** 3 / world, "traffic": None,
                         "peak_source": f"profiles/fp64_peak.json ({peak['how']}; {peak['sm_mhz']} MHz)"},
            "e2e": {"value": j["flops_step"] / (j["e2e_ms"] * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": j["e2e_bytes"][0],
                    "d2h_bytes_per_step": j["e2e_bytes"][1], "ms_per_step": j["e2e_ms"], "steps": j["e2e_steps"], "api": j["e2e_api"]},
            "gpu_launches": j["launches"], "clocks": j["clocks"], "host_affinity_rank0": G.affinity})
        if not args.no_replicas:  # the round-1 weak-scaling number: one independent n = 8192 step per GPU, no collective
            r = metric_step_n8192(G, args, max(2, min(args.steps, 3)), 3, False, None, seed_offset=rank)
            r.pop("host")
            line["replicas"] = {"workload": f"one (times + expm) n={args.n} instance per GPU, no collective (weak scaling)",
                                "gflops_total": world * r["flops_step"] / (r["ms_step"] * 1e-3) / 1e9, "ms_per_step": r["ms_step"]}

    if world > 1 and not args.no_single_process:
        # the other ranks wait on the HOST (a store key, no NCCL kernel spinning on their GPUs) while rank 0 drives every device
        store = G.dist.distributed_c10d._get_default_store()
        G.barrier()
        if rank == 0:
            try:
                sp = single_process_mg(G, args)
                line["single_process_mg"] = sp
                if not sp["ok"]:
                    failed.append(f"single-process multi-GPU verification failed: {sp}")
            finally:
                store.set("single_process_mg_done", "1")
        else:
            store.wait(["single_process_mg_done"])
        G.barrier()

    extra = {}
    if world == 1 and not args.no_extra_configs:
        mode = fm.EXPM_MIN_PRODUCTS if args.expm_mode == 1 else fm.EXPM_REFERENCE_ORDER
        # T_1 of the strong-scaling job, same code path as N > 1 (a column "block" of all columns, no peers)
        if args.colshard_n > 0:
            j1 = sharded_job(G, args, max(2, min(args.steps, 3)), 2, False, None, compare_nccl=False)
            if not j1["checks"]["all_ranks_ok"]:
                failed.append(f"sharded job (N=1) verification failed: {j1['checks']}")
            line["sharded_job_n1"] = {"ms_per_step": j1["ms_step"], "gflops": j1["flops_step"] / (j1["ms_step"] * 1e-3) / 1e9,
                                      "verified": j1["checks"], "times": j1["times"], "expm_batched": j1.get("expm_batched"),
                                      "note": "T_1 for the N > 1 lines: strong-scaling efficiency = T_1 / (N T_N)"}
        ns, nrhs = 8192, 64                                   # config 3: solve n=8192, 64 right-hand sides (mldivide: copies + dgesv)
        As, Bs = G.dev_randn(ns, ns, 3001), G.dev_randn(ns, nrhs, 3002)
        ms = G.best_ms(lambda: fm.mldivide(As, Bs))
        w = (2.0 / 3.0) * ns ** 3 + 2.0 * ns * ns * nrhs
        extra["solve"] = {"n": ns, "nrhs": nrhs, "ms": ms, "tflops": w / ms / 1e9, "pct_of_peak": 100.0 * w / ms / 1e9 / PEAK,
                          "api": "mldivide (flomat.rkt:3289): copy A, copy B, LU with partial pivoting + look-ahead, 2 triangular sweeps"}
        lu = fm.copy_flomat(As)
        piv = fm.Pivots(ns)
        ms = G.best_ms(lambda: (lib.fm_copy2d(lu.a, ns, As.a, ns, ns, ns), lib.fm_getrf(ns, ns, lu.a, ns, piv.ptr, piv.info_ptr)))
        extra["getrf"] = {"n": ns, "ms": ms, "tflops": (2.0 / 3.0) * ns ** 3 / ms / 1e9, "includes": "the copy of A (0.16 ms)"}
        ms = G.best_ms(lambda: fm.inv(As))
        extra["inv"] = {"n": ns, "ms": ms, "tflops": 2.0 * ns ** 3 / ms / 1e9}
        del As, Bs, lu, piv
        ne = 512                                              # config 2: expm 512x512
        Xe = G.dev_randn(ne, ne, 2001, 1.0 / math.sqrt(ne))
        _, se = fm.expm(Xe, mode, return_s=True)
        ms = G.best_ms(lambda: fm.expm(Xe, mode), reps=5)
        extra["expm512"] = {"n": ne, "s": se, "ms": ms, "gflops_wmin": expm_flops_min(ne, se) / ms / 1e6}
        del Xe
        nw = 8192                                             # elementwise rows of SURVEY 8(d): HBM-bound, bit-exact ops
        Aw, Bw = G.dev_randn(nw, nw, 7001), G.dev_randn(nw, nw, 7002)
        Cw = fm.Flomat.alloc(nw, nw)
        hbm = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm = float(json.load(f)["hbm_gbs"])
        except (OSError, KeyError, ValueError):
            pass
        for name, fn in (("plus", lambda: fm.dot_plus(Aw, Bw, Cw)), ("dot_times", lambda: fm.dot_times(Aw, Bw, Cw))):
            ms = G.best_ms(fn, reps=5)
            extra[name] = {"n": nw, "ms": ms, "gbs": 24.0 * nw * nw / ms / 1e6, "bytes_per_elem": 24,
                           "frac_of_measured_hbm": (24.0 * nw * nw / ms / 1e6 / hbm) if hbm else None}
        del Aw, Bw, Cw
        # ---- the compat tier as the unmodified reference calls it: PAGEABLE host operands, wall clock ----
        nh = args.n
        pa, pb = np.asfortranarray(hA.T), np.asfortranarray(hB.T)
        pc = np.zeros((nh, nh), order="F")
        # flomat* (flomat.rkt:905-914) allocates a ZERO matrix and calls dgemm with alpha = beta = 1: C travels both ways
        call = lambda: lib.cblas_dgemm(102, 111, 111, nh, nh, nh, 1.0, pa.ctypes.data, nh, pb.ctypes.data, nh, 1.0, pc.ctypes.data, nh)  # noqa: E731
        call()
        best = 1e30
        for _ in range(3):
            pc.fill(0.0)
            t0 = time.perf_counter(); call(); best = min(best, (time.perf_counter() - t0) * 1e3)
        extra["cblas_dgemm_host"] = {"n": nh, "ms": best, "tflops_end_to_end": 2.0 * nh ** 3 / best / 1e9, "host_memory": "pageable",
                                     "alpha_beta": [1.0, 1.0], "h2d_bytes": 3 * nh * nh * 8, "d2h_bytes": nh * nh * 8,
                                     "path": "fm_dgemm_host: column-block pipeline, pageable memory staged by the library's copy threads"}
        del pa, pb, pc
        from ctypes import byref, c_int
        rs = np.random.default_rng(3001)
        a0, b0 = np.asfortranarray(rs.standard_normal((ns, ns))), np.asfortranarray(rs.standard_normal((ns, nrhs)))
        ipv, inf = np.zeros(ns, dtype=np.int32), c_int(0)
        best = 1e30
        for _ in range(3):
            a1, b1 = a0.copy(order="F"), b0.copy(order="F")
            t0 = time.perf_counter()
            lib.dgesv_(byref(c_int(ns)), byref(c_int(nrhs)), a1.ctypes.data, byref(c_int(ns)), ipv.ctypes.data, b1.ctypes.data, byref(c_int(ns)), byref(inf))
            best = min(best, (time.perf_counter() - t0) * 1e3)
        resid = float(np.abs(a0[:256] @ b1 - b0[:256]).max())
        extra["dgesv_host"] = {"n": ns, "nrhs": nrhs, "ms": best, "tflops_end_to_end": w / best / 1e9, "host_memory": "pageable", "info": inf.value,
                               "h2d_bytes": (ns * ns + ns * nrhs) * 8, "d2h_bytes": (ns * ns + ns * nrhs) * 8, "max_residual_rows_0_255": resid}
        del a0, b0, a1, b1
        if not args.no_compat_expm:
            # expm n = 8192 with ZERO changes to the Racket sources: every foreign call of expm.rkt through the compat symbols, host buffers
            from sci_b200.racket_front import RacketFrontEnd
            front = RacketFrontEnd()
            px = np.asfortranarray(hX.T)
            front.expm(px[:512, :512].copy(order="F"))
            front.calls.clear()
            t0 = time.perf_counter()
            front.expm(px)
            dt = time.perf_counter() - t0
            extra["expm_compat"] = {"n": nh, "seconds": dt, "gflops_wmin": expm_flops_min(nh, m["s"]) / dt / 1e9, "host_memory": "pageable",
                                    "foreign_calls": dict(front.calls),
                                    "what": "sci_b200/racket_front.py: the unmodified expm.rkt / flomat.rkt call sequence (11+s cblas_dgemm, 11 x n "
                                            "cblas_daxpy, cblas_dcopy, dlange_, dgesv_) on libflomat_cuda's compat symbols; Racket's own loops on the host"}
            del px
        lib.fm_trim_cache()
    line["other_configs"] = extra

    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # CPU baseline: rank 0 at N=1 only -- ONE real full step
        cpu = CpuMetricStep(args.n)
        t_full = cpu.full()
        t_s, detail = cpu.sampled()
        line["cpu_baseline"] = {"value": cpu.flops / t_full / 1e9, "unit": "GFLOP/s", "cores": cpu.prov.threads(), "kind": "port",
                                "sample": f"ONE real step of the reference call sequence on {cpu.prov.name}: times n={args.n} + expm n={args.n} "
                                          f"(s={int(cpu.s)}, {11 + int(cpu.s)} dgemms + dgesv executed, W_min counted): {t_full:.1f} s",
                                "seconds_full_step": t_full,
                                "sampled_estimate": {"value": cpu.flops / t_s / 1e9, "seconds": t_s, "rel_diff_vs_full": t_s / t_full - 1.0,
                                                     "what": "the bounded sample `--impl reference` times (see its cpu_baseline.sample)", **detail}}
    else:
        line["cpu_baseline"] = None

    if rank == 0:
        line["failed"] = failed
        print(json.dumps(line), flush=True)
    if world > 1:
        G.dist.destroy_process_group()
    if failed:
        print("bench.py: " + "; ".join(failed), file=sys.stderr)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=8192, help="size of the metric's times + expm step (N = 1)")
    ap.add_argument("--expm-mode", type=int, default=1, help="1: 5+s products (default), 0: reference Horner order 7+s")
    ap.add_argument("--colshard-n", type=int, default=16384, help="size of the column-sharded times of the sharded job (config 4)")
    ap.add_argument("--batch-items", type=int, default=4096, help="items of the batched expm of the sharded job (config 5)")
    ap.add_argument("--batch-n", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip sharded_job_n1 and the solve / expm512 / elementwise / compat sub-measurements")
    ap.add_argument("--no-compat-expm", action="store_true")
    ap.add_argument("--no-replicas", action="store_true")
    ap.add_argument("--no-single-process", action="store_true", help="skip the fm_mg_* (one process, all devices) sub-measurement at N > 1")
    ap.add_argument("--ref-budget-s", type=float, default=640.0, help="--impl reference at N=1 runs real full steps if K of them fit this many seconds")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
