"""Residual check of mldivide at sizes above the unit tests (exercises the 1024-rows-per-CTA panel configuration)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
for n in [int(x) for x in (sys.argv[1:] or ["12288", "16384"])]:
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    a = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)   # row-major torch = column-major A^T: still N(0,1)
    x0 = torch.randn(n, 8, dtype=torch.float64, device="cuda", generator=g)
    At = a.t().contiguous()                                                  # At[j, i] = a[i, j]: buffer of column-major `a`
    b = a @ x0
    A = fm.Flomat.alloc(n, n); lib.fm_copy2d(A.a, n, At.data_ptr(), n, n, n)
    B = fm.Flomat.alloc(n, 8); lib.fm_copy2d(B.a, n, b.t().contiguous().data_ptr(), n, n, 8)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); X = fm.mldivide(A, B); e1.record(); torch.cuda.synchronize()
    x = torch.empty(8, n, dtype=torch.float64, device="cuda"); lib.fm_copy2d(x.data_ptr(), n, X.a, n, n, 8); fm.sync()
    x = x.t()
    res = (a @ x - b).abs().sum(0).max() / (a.abs().sum(0).max() * x.abs().sum(0).max())
    err = (x - x0).abs().max() / x0.abs().max()
    print(f"n={n}: mldivide {e0.elapsed_time(e1):.1f} ms, backward error {res.item():.2e} (64 n eps = {64*n*2**-52:.2e}), forward error {err.item():.2e}", flush=True)
    del a, At, b, A, B, X, x
    lib.fm_trim_cache(); torch.cuda.empty_cache()
