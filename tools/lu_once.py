"""One warm-up + one measured getrf (+ getrs with LU_NRHS right-hand sides) at n=LU_N (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("LU_N", "8192")); nrhs = int(os.environ.get("LU_NRHS", "64"))
g = torch.Generator(device="cuda"); g.manual_seed(3)
t = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g)
b = torch.randn(n * nrhs, dtype=torch.float64, device="cuda", generator=g)
A = fm.Flomat.alloc(n, n); piv = fm.Pivots(n)
reps = int(os.environ.get("LU_REPS", "2"))
for _ in range(reps):
    lib.fm_copy2d(A.a, n, t.data_ptr(), n, n, n)
    lib.fm_getrf(n, n, A.a, n, piv.ptr, piv.info_ptr)
    if nrhs:
        lib.fm_getrs(n, nrhs, A.a, n, piv.ptr, b.data_ptr(), n)
fm.sync()
print("info", piv.info(), "launches", lib.fm_kernel_launches())
