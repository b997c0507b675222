"""expm_batched 512 x 256^2 and the batched 256^3 product alone (CUDA events, best of 10)."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load(); dev = torch.device("cuda", 0)
def best(fn, reps=10):
    fn(); torch.cuda.synchronize(); b = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
g = torch.Generator(device=dev); g.manual_seed(5001)
n, batch = 256, 512
src = torch.randn(n * n * batch, dtype=torch.float64, device=dev, generator=g) / 16.0; dst = torch.empty_like(src); c = torch.empty_like(src)
t_e = best(lambda: fm.expm_batched(src.data_ptr(), n, batch, dst.data_ptr(), 1, None))
t_g = best(lambda: lib.fm_dgemm_batched(n, n, n, 1.0, src.data_ptr(), n, n * n, dst.data_ptr(), n, n * n, 0.0, c.data_ptr(), n, n * n, 0.0, batch))
print(json.dumps({"stagger_all": os.environ.get("FLOMAT_GEMM_STAGGER_ALL"), "stagger_ns": os.environ.get("FLOMAT_GEMM_STAGGER_NS"), "expm_batched_ms": round(t_e, 4),
                  "gemm256_batched_ms": round(t_g, 4), "gemm_tflops": round(2.0 * n ** 3 * batch / t_g / 1e9, 2)}))
