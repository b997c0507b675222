"""Two fm_getrf_nopiv calls at n = NP_N on a column diagonally dominant matrix (for ncu launch lists) + CUDA-event time."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("NP_N", "8192"))
g = torch.Generator(device="cuda"); g.manual_seed(3)
t = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g) * (0.5 / n)
t.view(n, n).diagonal().add_(1.0)
A = fm.Flomat.alloc(n, n); piv = fm.Pivots(n + 4)
for rep in range(2):
    lib.fm_copy2d(A.a, n, t.data_ptr(), n, n, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = lib.fm_getrf_nopiv(n, A.a, n, piv.ptr, piv.info_ptr, piv.ptr + 4 * n); e1.record(); torch.cuda.synchronize()
h = np.zeros(1, dtype=np.int32); lib.fm_download_i32(h.ctypes.data, piv.ptr + 4 * n, 1)
print("rc", rc, "viol", int(h[0]), "ms", round(e0.elapsed_time(e1), 3))
