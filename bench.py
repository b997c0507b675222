#!/usr/bin/env python
"""bench.py -- headline benchmark of the flomat dense-FP64 hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 8192]

One "step" = one pass of the hot path over one batch of synthetic input:  C = (times A B)  followed by  E = (expm X)
at n = 8192 (the configuration BASELINE.json's metric is quoted on), with A, B ~ N(0,1) and X ~ N(0,1)/sqrt(n)
(SURVEY 8d).  value = algorithmic FP64 GFLOP/s of the whole job = N * (2 n^3 + W_min) / t, where
W_min = (5+s) 2n^3 + 8/3 n^3 is the fixed algorithmic work of one expm and t is the max over ranks of the
CUDA-event time of the K timed steps.  Inputs are resident in HBM when the timed region starts.

N > 1 (launched by torchrun, one rank per GPU): the path shards by independent items -- every rank runs the step on
its own matrices (batch items per GPU, no data-path collective): "scaling": "weak".  The column-sharded `times`
n = 16384 with the all-gather fused into the GEMM epilogue over NVLink peer memory (BASELINE config 4) is measured
in the same run and reported under "times_colshard" (total work fixed: that sub-number is strong scaling).

e2e = the same step through the public host API with HOST buffers: upload from pinned memory, compute, download.
cpu_baseline / --impl reference = the reference's own CPU path (its exact call sequence on OpenBLAS, all host
threads; oracle/flomat_oracle.py) on a bounded sample.
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

FP64_DMMA_PEAK_TFLOPS = 37.0  # tools/peak_fp64.cu on this pool's B200 @1965 MHz (profiles/r01_peak_fp64.csv)


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
# CPU reference arm: the reference's call sequence on OpenBLAS (oracle) -- bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_times: int, n_expm: int, steps: int, warmup: int):
    import numpy as np
    from oracle.flomat_oracle import Flomat, default_provider

    prov = default_provider()
    if hasattr(prov, "_set"):  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses all the host threads it can
        prov._set(os.cpu_count() or 1)
    F = Flomat(prov)
    rng = np.random.default_rng(6001)
    a = np.asfortranarray(rng.standard_normal((n_times, n_times)))
    b = np.asfortranarray(rng.standard_normal((n_times, n_times)))
    x = np.asfortranarray(rng.standard_normal((n_expm, n_expm)) / math.sqrt(n_expm))
    s = F.scaling_exponent(x)
    flops = 2.0 * n_times ** 3 + expm_flops_min(n_expm, s)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        F.times(a, b)
        F.expm(x)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return {"value": flops / t / 1e9, "unit": "GFLOP/s", "cores": prov.threads(), "kind": "port",
            "sample": f"reference call sequence (flomat.rkt:888, expm.rkt:193-210) on {prov.name}: times n={n_times} + expm "
                      f"n={n_expm} (s={int(s)}; W_min counted, the reference executes {11 + int(s)} dgemms), "
                      f"{steps} timed steps of {t:.2f} s",
            "ms_per_step": t * 1e3, "provider": prov.name}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_sample(args.ref_n_times, args.ref_n_expm, args.steps, max(args.warmup, 1) if args.warmup else 1)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.n, args.gpus),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


METRIC = "FP64 GFLOP/s for times (dgemm) and expm at n=8192; % of FP64 tensor peak"


def workload_config(n, gpus):
    return {"workload": f"flomat (times A B) + (expm X), n={n}, FP64, A,B~N(0,1), X~N(0,1)/sqrt(n); one independent instance per GPU",
            "n": n, "expm": "Pade[8/8] scaling+squaring (expm.rkt:203), 5+s products + LU solve with n rhs",
            "l2": "inputs (3 x 512 MiB at n=8192) exceed the 126 MB L2; no flush needed", "parallelism": f"items x{gpus}"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import sci_b200 as fm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    fm.init(local)
    lib = fm.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs, generated on the host with the SURVEY 8d seeds (+rank), resident in HBM before timing ----
    rng = np.random.default_rng(6001 + 17 * rank)
    pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    hA, hB, hX = pin((n, n)), pin((n, n)), pin((n, n))  # C-order host buffers hold the column-major matrices' bytes
    for h, scale in ((hA, 1.0), (hB, 1.0), (hX, 1.0 / math.sqrt(n))):
        h[...] = rng.standard_normal((n, n)) * scale
    hC, hE = pin((n, n)), pin((n, n))

    def upload(h):
        out = fm.Flomat.alloc(n, n)
        fm._lib.check(lib.fm_upload(out.a, n, h.ctypes.data, n, n, n), "fm_upload")
        return out

    A, B, X = upload(hA), upload(hB), upload(hX)
    fm.sync()
    C = fm.Flomat.alloc(n, n)
    mode = fm.EXPM_MIN_PRODUCTS if args.expm_mode == 1 else fm.EXPM_REFERENCE_ORDER

    def step():
        fm.flomat_times(A, B, C, 1.0, 0.0)
        return fm.expm(X, mode, return_s=True)

    # ---- warm-up (also sizes the workspaces) ----
    s_val = 0.0
    for _ in range(max(args.warmup, 3)):
        _, s_val = step()
    barrier()
    flops_step = 2.0 * n ** 3 + expm_flops_min(n, s_val)

    # ---- timed region: exactly K steps, CUDA events on the launching (default) stream ----
    sampler = ClockSampler(local)
    launches0 = lib.fm_kernel_launches()
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    et = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record()
    for k in range(args.steps):
        et[k][0].record()
        fm.flomat_times(A, B, C, 1.0, 0.0)
        et[k][1].record()
        fm.expm(X, mode)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.fm_kernel_launches() - launches0
    ms_total = e0.elapsed_time(e1)
    ms_times = sum(a.elapsed_time(b) for a, b in et) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tt = torch.tensor([ms_times], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, ms_times = float(t.item()), float(tt.item())
    ms_step = ms_total / args.steps
    value = world * flops_step / (ms_step * 1e-3) / 1e9
    times_tflops = 2.0 * n ** 3 / (ms_times * 1e-3) / 1e12
    expm_ms = ms_step - ms_times
    expm_gflops = expm_flops_min(n, s_val) / (expm_ms * 1e-3) / 1e9

    # ---- e2e: host buffers in, host buffers out, through the public API, every step ----
    def e2e_step():
        # `times` on HOST operands through fm_dgemm_host (what cblas_dgemm does with the reference's host matrices): A goes up, then
        # B and C move in column blocks beside the product.  X goes up behind B's blocks while `times` runs; `expm` delivers its
        # result to the host buffer block by block during the last squaring.
        x_ = fm.Flomat.alloc(n, n)
        fm.times_host(hA.T, hB.T, hC.T, 1.0, 0.0, wait=False)
        x_.upload_async(hX.reshape(-1))
        fm.expm_to_host(x_, hE.T, mode)  # the last squaring's column blocks come down while the remaining ones are computed
        fm.transfer_sync()

    e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(e2e_steps):
        e2e_step()
    g1.record()
    barrier()
    e2e_ms = max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e = {"value": world * flops_step / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 3 * n * n * 8,
           "d2h_bytes_per_step": 2 * n * n * 8, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "per step, pinned host buffers: times_host (fm_dgemm_host: A up, B/C in column blocks beside the product) -> X upload_async beside it -> expm with a host result buffer (last squaring in column blocks, downloads beside it)"}

    # ---- column-sharded times (BASELINE config 4): fused P2P all-gather epilogue when N > 1 ----
    colshard = None
    if args.colshard_n > 0:
        from sci_b200.multigpu import ColShardTimes
        del A, B, C, X
        cs = ColShardTimes(args.colshard_n, world, rank, seed=4001)
        colshard = cs.bench(steps=max(2, min(args.steps, 3)), warmup=2)
        cs.close()

    # ---- the other BASELINE.json configurations (device-resident, CUDA events, best of 3): reported, not part of `value` ----
    def best_ms(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); fn(); a1.record(); torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(a1))
        return best

    def dev_randn(rows, cols, seed, scale=1.0):
        g = torch.Generator(device=dev); g.manual_seed(seed)
        t = torch.randn(rows * cols, dtype=torch.float64, device=dev, generator=g) * scale
        out = fm.Flomat.alloc(rows, cols)
        fm._lib.check(lib.fm_copy2d(out.a, rows, t.data_ptr(), rows, rows, cols), "fm_copy2d")
        fm.sync()
        return out

    extra = {}
    if not args.no_extra_configs:
        ns, nrhs = 8192, 64                                   # config 3: solve n=8192, 64 right-hand sides (mldivide: copies + dgesv)
        As, Bs = dev_randn(ns, ns, 3001 + rank), dev_randn(ns, nrhs, 3002 + rank)
        ms = best_ms(lambda: fm.mldivide(As, Bs))
        w = (2.0 / 3.0) * ns ** 3 + 2.0 * ns * ns * nrhs
        extra["solve"] = {"n": ns, "nrhs": nrhs, "ms": ms, "tflops": w / ms / 1e9, "pct_of_peak": 100.0 * w / ms / 1e9 / FP64_DMMA_PEAK_TFLOPS,
                          "api": "mldivide (flomat.rkt:3289): copy A, copy B, LU with partial pivoting + look-ahead, 2 triangular sweeps"}
        del As, Bs
        ne = 512                                              # config 2: expm 512x512
        Xe = dev_randn(ne, ne, 2001 + rank, 1.0 / math.sqrt(ne))
        _, se = fm.expm(Xe, mode, return_s=True)
        ms = best_ms(lambda: fm.expm(Xe, mode))
        extra["expm512"] = {"n": ne, "s": se, "ms": ms, "gflops_wmin": expm_flops_min(ne, se) / ms / 1e6}
        del Xe
        nb_, per_gpu = 256, 4096 // 8                         # config 5: 4096 items of 256x256 over 8 GPUs = 512 items per GPU
        g = torch.Generator(device=dev); g.manual_seed(5001 + rank)
        src = torch.randn(nb_ * nb_ * per_gpu, dtype=torch.float64, device=dev, generator=g) / 16.0
        dst = torch.empty_like(src)
        s_out = np.zeros(per_gpu)
        barrier()
        ms = best_ms(lambda: fm.expm_batched(src.data_ptr(), nb_, per_gpu, dst.data_ptr(), mode, s_out))
        wb = float(sum(expm_flops_min(nb_, sv) for sv in s_out))
        tb = torch.tensor([ms, wb], dtype=torch.float64, device=dev)
        if world > 1:  # config 5 proper: the batch is sharded by items, no collective; job time = the slowest rank
            tmax, wsum = tb.clone(), tb.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(wsum, op=dist.ReduceOp.SUM)
            ms_job, w_job = float(tmax[0].item()), float(wsum[1].item())
        else:
            ms_job, w_job = ms, wb
        extra["expm_batched"] = {"n": nb_, "items_per_gpu": per_gpu, "items_total": per_gpu * world, "ms": ms_job,
                                 "tflops_wmin_total": w_job / ms_job / 1e9, "tflops_wmin_per_gpu": w_job / ms_job / 1e9 / world,
                                 "pct_of_peak": 100.0 * w_job / ms_job / 1e9 / world / FP64_DMMA_PEAK_TFLOPS, "s_max": float(s_out.max()),
                                 "sharding": "items (batch index) per GPU, no data-path collective; time = max over ranks"}
        del src, dst
        nw = 8192                                             # elementwise rows of SURVEY 8(d): HBM-bound, bit-exact ops
        Aw, Bw = dev_randn(nw, nw, 7001 + rank), dev_randn(nw, nw, 7002 + rank)
        Cw = fm.Flomat.alloc(nw, nw)
        ms = best_ms(lambda: fm.dot_plus(Aw, Bw, Cw), reps=5)
        extra["plus"] = {"n": nw, "ms": ms, "gbs": 24.0 * nw * nw / ms / 1e6, "bytes_per_elem": 24}
        ms = best_ms(lambda: fm.dot_times(Aw, Bw, Cw), reps=5)
        extra["dot_times"] = {"n": nw, "ms": ms, "gbs": 24.0 * nw * nw / ms / 1e6, "bytes_per_elem": 24}
        del Aw, Bw, Cw
        if rank == 0 and world == 1:
            # the compat symbol as the unmodified reference calls it (flomat.rkt:888): PAGEABLE host operands, alpha = 1, wall clock
            nh = n
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
            # config 3 the same way: dgesv_ on pageable host operands (flomat.rkt:2140 after its two copies); A comes back as L\U
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
        fm.load().fm_trim_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # CPU baseline: rank 0 at N=1 only
        cpu = cpu_reference_sample(args.ref_n_times, args.ref_n_expm, 1, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(n, world),
                "pct_of_fp64_tensor_peak": 100.0 * value / world / 1e3 / FP64_DMMA_PEAK_TFLOPS,
                "times": {"n": n, "ms": ms_times, "tflops": times_tflops, "pct_of_peak": 100.0 * times_tflops / FP64_DMMA_PEAK_TFLOPS},
                "expm": {"n": n, "s": s_val, "ms": expm_ms, "gflops_wmin": expm_gflops, "gflops_wref": expm_flops_ref(n, s_val) / (expm_ms * 1e-3) / 1e9,
                         "pct_of_peak_wmin": 100.0 * expm_gflops / 1e3 / FP64_DMMA_PEAK_TFLOPS, "mode": "5+s" if args.expm_mode == 1 else "7+s"},
                "roofline": {"bound": "tensor", "kernel": "dgemm_tma_kernel (DMMA.8x8x4 + TMA)", "achieved": times_tflops,
                             "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": times_tflops / FP64_DMMA_PEAK_TFLOPS,
                             "traffic": 6.98e9, "traffic_unit": "bytes/launch (dram read+write, ncu --set full, profiles/r01_dgemm_tma_n8192_ncu_full.txt); algorithmic 1.61e9",
                             "peak_source": "measured DMMA issue rate, tools/peak_fp64.cu (MEASURED_PEAKS.json has no FP64 entry)",
                             "flops_per_launch": 2.0 * n ** 3},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "times_colshard": colshard,
                "other_configs": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--expm-mode", type=int, default=1, help="1: 5+s products (default), 0: reference Horner order 7+s")
    ap.add_argument("--colshard-n", type=int, default=16384, help="0 disables the column-sharded times sub-benchmark")
    ap.add_argument("--ref-n-times", type=int, default=8192)
    ap.add_argument("--ref-n-expm", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the solve / expm512 / batched expm / elementwise sub-measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
