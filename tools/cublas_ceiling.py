"""Sanity ceiling only (never on the product path): what cuBLAS DGEMM reaches on this B200.
Used once to calibrate the FP64 roofline denominator in DESIGN.md."""
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
out = {}
for n in (2048, 4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[n] = {"ms": best, "tflops": 2 * n**3 / best / 1e9}
    print(n, out[n], flush=True)
    del a, b, c
# LU ceiling
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
torch.linalg.lu_factor(a); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); torch.linalg.lu_factor(a); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("getrf8192", ms, (2/3) * n**3 / ms / 1e9, flush=True)
out["getrf8192"] = {"ms": ms, "tflops": (2/3) * n**3 / ms / 1e9}
json.dump(out, open("gpurun_out/cublas_ceiling.json", "w"))
