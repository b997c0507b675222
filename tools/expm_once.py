"""One warm-up + one measured expm at n=EXPM_N (BATCH>0: batched 256x256) for ncu launch lists."""
import sys, os, math
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("EXPM_N", "512")); batch = int(os.environ.get("BATCH", "0")); mode = int(os.environ.get("MODE", "1"))
g = torch.Generator(device="cuda"); g.manual_seed(5)
if batch:
    src = torch.randn(n * n * batch, dtype=torch.float64, device="cuda", generator=g) / math.sqrt(n)
    dst = torch.empty_like(src); s_out = np.zeros(batch)
    for _ in range(2): fm.expm_batched(src.data_ptr(), n, batch, dst.data_ptr(), mode, s_out)
else:
    t = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g) / math.sqrt(n)
    X = fm.Flomat.alloc(n, n); lib.fm_copy2d(X.a, n, t.data_ptr(), n, n, n)
    for _ in range(2): fm.expm(X, mode)
fm.sync()
print("launches", lib.fm_kernel_launches())
