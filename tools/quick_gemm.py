"""Quick GPU sanity + timing for the dgemm kernel (developer tool)."""
import sys, time, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np
import sci_b200 as fm
fm.init(0)
lib = fm.load()
import ctypes

def bench(n, reps=5):
    a = fm.Flomat.alloc(n, n); b = fm.Flomat.alloc(n, n); c = fm.Flomat.alloc(n, n)
    lib.fm_fill(a.a, n, n, n, 1.0); lib.fm_fill(b.a, n, n, n, 0.5)
    import torch
    for _ in range(2):
        fm.flomat_times(a, b, c, 1.0, 0.0)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fm.flomat_times(a, b, c, 1.0, 0.0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"dgemm n={n}: {best:.3f} ms  {2*n**3/best/1e9:.2f} TFLOP/s  ({2*n**3/best/1e9/37.0*100:.1f}% of 37.0)", flush=True)

rng = np.random.default_rng(0)
for (m, n, k) in [(8, 8, 4), (128, 128, 128), (130, 258, 70), (512, 384, 1000)]:
    a = np.asfortranarray(rng.standard_normal((m, k))); b = np.asfortranarray(rng.standard_normal((k, n)))
    got = fm.times(fm.matrix(a), fm.matrix(b)).to_numpy()
    err = np.abs(got - a @ b).max() / np.abs(a @ b).max()
    print("check", (m, n, k), err, flush=True)
    assert err < 1e-12
for n in [int(x) for x in (sys.argv[1:] or ["1024", "2048", "4096", "8192"])]:
    bench(n)
