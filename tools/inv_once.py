"""One warm-up + one measured inv (copy + getrf + getri) at n = INV_N, for ncu launch lists (developer tool)."""
import sys, os, math
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("INV_N", "8192"))
g = torch.Generator(device="cuda"); g.manual_seed(5)
t = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g)
A = fm.Flomat.alloc(n, n); lib.fm_copy2d(A.a, n, t.data_ptr(), n, n, n)
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); Z = fm.inv(A); e1.record(); fm.sync()
print(f"inv n={n}: {e0.elapsed_time(e1):.2f} ms, launches {lib.fm_kernel_launches()}")
