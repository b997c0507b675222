"""CUDA-event times of small square products (the sizes split-K serves) and of expm n = 512 / 1024; FLOMAT_GEMM_SPLITK_MIN_K sets the
shortest k slice."""
import json, math, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
dev = torch.device("cuda", 0)
def best(fn, reps=20, inner=10):
    fn(); torch.cuda.synchronize(); b = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1) / inner)
    return b
out = {"splitk_min_k": os.environ.get("FLOMAT_GEMM_SPLITK_MIN_K", "default")}
for n in (256, 384, 512, 640, 768, 1024, 1100):
    a = torch.randn(n * n, dtype=torch.float64, device=dev); b = torch.randn(n * n, dtype=torch.float64, device=dev); c = torch.zeros(n * n, dtype=torch.float64, device=dev)
    f = lambda: lib.fm_dgemm(111, 111, n, n, n, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0, c.data_ptr(), n)
    us = best(f) * 1e3
    want = b.view(n, n) @ a.view(n, n)  # row-major views of column-major data: C^T = B^T A^T
    err = float((c.view(n, n) - want).abs().max() / want.abs().max())
    out[f"dgemm_{n}"] = {"us": round(us, 2), "tflops": round(2 * n ** 3 / us / 1e6, 2), "err": err}
for n in (512, 1024):
    t = torch.randn(n * n, dtype=torch.float64, device=dev) / math.sqrt(n)
    X = fm.Flomat.alloc(n, n); lib.fm_copy2d(X.a, n, t.data_ptr(), n, n, n); fm.sync()
    out[f"expm_{n}"] = {"ms": round(best(lambda: fm.expm(X, 1), 10, 1), 4)}
print(json.dumps(out))
