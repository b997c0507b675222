"""Time the batched dgemm of the batched-expm shape (BATCH x n^3) under the current FLOMAT_GEMM_* switches (developer tool)."""
import sys, os, math
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("N", "256")); batch = int(os.environ.get("BATCH", "512")); beta = float(os.environ.get("BETA", "0"))
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.randn(n * n * batch, dtype=torch.float64, device="cuda", generator=g)
b = torch.randn(n * n * batch, dtype=torch.float64, device="cuda", generator=g)
c = torch.zeros(n * n * batch, dtype=torch.float64, device="cuda")
def run():
    rc = lib.fm_dgemm_batched(n, n, n, 1.0, a.data_ptr(), n, n * n, b.data_ptr(), n, n * n, beta, c.data_ptr(), n, n * n, 0.0, batch)
    assert rc == 0
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); t = ts[len(ts) // 2]
fl = 2.0 * n ** 3 * batch
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("FLOMAT_"))
print(f"batched dgemm {batch} x {n}^3 beta={beta}: {t*1e3:.1f} us  {fl/t/1e9:.2f} TFLOP/s ({fl/t/1e9/37.0*100:.1f}%)  [{tag}]", flush=True)
