"""CPU: the benchmark's reference arm (`bench.py --impl reference`, the reference's call sequence on OpenBLAS) runs without a
GPU and prints one JSON line with the keys the driver reads -- for the N = 1 configuration and for the sharded N > 1 job."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))


@pytest.mark.parametrize("flags,scaling", [(["--gpus", "1", "--n", "96"], "weak"),
                                           (["--gpus", "2", "--colshard-n", "256", "--batch-items", "8", "--batch-n", "32"], "strong")])
def test_reference_arm_prints_contract_line(flags, scaling):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", *flags],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["dtype"] == "f64" and line["vs_baseline"] is None and line["scaling"] == scaling
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == int(flags[1])
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=60, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_sampled_cpu_step_matches_a_real_step():
    """The bounded sample the reference arm times (every distinct foreign call once, summed over the reference's call sequence) against
    one real step, at a size where the real step takes a second: the two agree closely enough to stand in for each other."""
    sys.path.insert(0, ROOT)
    import bench

    step = bench.CpuMetricStep(768)
    step.full()
    t_full = min(step.full() for _ in range(2))
    t_s = min(step.sampled()[0] for _ in range(2))
    assert abs(t_s / t_full - 1.0) < 0.35, (t_s, t_full)
    assert step.flops == 2.0 * 768 ** 3 + bench.expm_flops_min(768, step.s)
