"""Products whose 128x128 tile count leaves a partial last wave on 148 SMs: 128x128 tiles (one CTA per SM) against 64x128 tiles (two CTAs
per SM; FLOMAT_GEMM_FORCE_HALF=1).  CUDA events, best of 5."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
dev = torch.device("cuda", 0)
def best(fn, reps=5):
    fn(); torch.cuda.synchronize(); b = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
out = {"force_half": os.environ.get("FLOMAT_GEMM_FORCE_HALF") is not None, "tail_rule": os.environ.get("FLOMAT_GEMM_NO_TAIL_RULE") is None}
for m, n, k in ((8192, 1024, 8192), (8192, 2048, 8192), (2560, 2560, 2560), (3072, 3072, 3072), (4096, 4096, 4096), (8192, 8192, 8192), (5120, 1024, 5120), (8192, 640, 8192), (2304, 2304, 2304), (4096, 2432, 4096), (4096, 3072, 4096), (6144, 1536, 6144), (4608, 4608, 4608), (10624, 3072, 4096), (8192, 1152, 8192), (1920, 1920, 1920)):
    a = torch.randn(m * k, dtype=torch.float64, device=dev); b = torch.randn(k * n, dtype=torch.float64, device=dev); c = torch.empty(m * n, dtype=torch.float64, device=dev)
    ms = best(lambda: lib.fm_dgemm(111, 111, m, n, k, 1.0, a.data_ptr(), m, b.data_ptr(), k, 0.0, c.data_ptr(), m))
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    out[f"{m}x{n}x{k}"] = {"ms": round(ms, 4), "tflops": round(2.0 * m * n * k / ms / 1e9, 2), "tiles128": tiles, "waves": round(tiles / 148, 2)}
    del a, b, c
print(json.dumps(out))
