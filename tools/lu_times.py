"""CUDA-event times (best of 5) of the LU-bound calls: getrf at several n, mldivide n=8192/64, inv n=4096, expm n=512, batched expm 512 x 256^2."""
import json, math, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
dev = torch.device("cuda", 0)
def best(fn, reps=5):
    fn(); torch.cuda.synchronize(); b = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
def randn(m, n, seed, scale=1.0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.randn(m * n, dtype=torch.float64, device=dev, generator=g) * scale
    out = fm.Flomat.alloc(m, n); lib.fm_copy2d(out.a, m, t.data_ptr(), m, m, n); fm.sync(); return out
out = {}
for n in (512, 1024, 2048, 4096, 8192):
    A = randn(n, n, 3); LU = fm.Flomat.alloc(n, n); piv = fm.Pivots(n)
    def f():
        lib.fm_copy2d(LU.a, n, A.a, n, n, n); lib.fm_getrf(n, n, LU.a, n, piv.ptr, piv.info_ptr)
    t_copy = best(lambda: lib.fm_copy2d(LU.a, n, A.a, n, n, n))
    ms = best(f) - t_copy
    out[f"getrf_{n}"] = {"ms": round(ms, 4), "tflops": round((2 / 3) * n ** 3 / ms / 1e9, 2)}
    del A, LU
A, B = randn(8192, 8192, 3001), randn(8192, 64, 3002)
ms = best(lambda: fm.mldivide(A, B), 3); out["mldivide_8192_64"] = {"ms": round(ms, 3), "tflops": round(((2 / 3) * 8192 ** 3 + 2 * 8192 ** 2 * 64) / ms / 1e9, 2)}
del A, B
A = randn(4096, 4096, 5); ms = best(lambda: fm.inv(A), 3); out["inv_4096"] = {"ms": round(ms, 3)}; del A
X = randn(512, 512, 2001, 1 / math.sqrt(512)); ms = best(lambda: fm.expm(X, 1), 10); out["expm_512"] = {"ms": round(ms, 4)}; del X
g = torch.Generator(device=dev); g.manual_seed(5001)
src = torch.randn(256 * 256 * 512, dtype=torch.float64, device=dev, generator=g) / 16.0; dst = torch.empty_like(src)
ms = best(lambda: fm.expm_batched(src.data_ptr(), 256, 512, dst.data_ptr(), 1, None)); out["expm_batched_512x256"] = {"ms": round(ms, 4)}
del src, dst
for n in (1024, 2048, 4096, 8192):
    X = randn(n, n, 6003, 1 / math.sqrt(n)); ms = best(lambda: fm.expm(X, 1), 3); out[f"expm_{n}"] = {"ms": round(ms, 3)}; del X
print(json.dumps(out))
