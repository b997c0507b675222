"""fm_mg_expm (one matrix exponential over every visible device, one process) against fm_expm on one device: wall clock around the
blocking call, device-resident and pageable-host operands, n = 4096 / 8192 (/ 16384 with --big)."""
import json, math, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
ndev = fm.mg_init(0)
out = {"devices": ndev}
def best(fn, reps=3):
    fn(); b = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t0)
    return b * 1e3
sizes = (4096, 8192) + ((16384,) if "--big" in sys.argv else ())
for n in sizes:
    a = np.asfortranarray(np.random.default_rng(6003).standard_normal((n, n)) / math.sqrt(n))
    A = fm.matrix(a)
    def one():
        fm.expm(A, 1); fm.sync()
    def mg():
        fm.expm_mg(A, mode=1)
    def mg_host():
        fm.expm_mg(a, mode=1)
    t1, tm, th = best(one), best(mg), best(mg_host, 2)
    _, s = fm.expm_mg(A, mode=1, return_s=True)
    wmin = ((5 + s) * 2 + 8 / 3) * n ** 3
    got, ref1 = fm.expm_mg(A, mode=1).to_numpy(), fm.expm(A, 1).to_numpy()
    err = float(np.linalg.norm(got - ref1, 1) / np.linalg.norm(ref1, 1))
    out[f"n{n}"] = {"s": s, "one_device_ms": round(t1, 2), "mg_device_operands_ms": round(tm, 2), "mg_host_operands_ms": round(th, 2),
                    "speedup": round(t1 / tm, 2), "tflops_wmin_total": round(wmin / tm / 1e9, 1), "rel_diff_vs_one_device": err, "gate": 64 * n * 2.0 ** -52}
    del A
print(json.dumps(out))
