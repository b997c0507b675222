"""Times a single 16-column leaf factorisation (getrf of an m x 16 matrix) and related small kernels."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
def t_ms(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for m in (256, 1024, 2048, 4096, 8192):
    for w in (1, 16, 32, 128):
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        t = torch.randn(m * w, dtype=torch.float64, device="cuda", generator=g)
        piv = fm.Pivots(16)
        ms = t_ms(lambda: lib.fm_getrf(m, w, t.data_ptr(), m, piv.ptr, piv.info_ptr))
        ms0 = t_ms(lambda: lib.fm_fill(piv.ptr, 2, 2, 1, 0.0))
        print(f"leaf m={m} w={w}: {ms*1e3:.1f} us  (an empty-ish launch: {ms0*1e3:.1f} us)", flush=True)
