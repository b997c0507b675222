"""Time of ONE panel factorization (m x 128) per configuration, CUDA events over 20 launches (input restored before each)."""
import sys, os, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
out = {}
for m in [int(x) for x in (sys.argv[1:] or ["128", "256", "512", "1024", "2048", "4096", "8192"])]:
    w = 128
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    t = torch.randn(m * w, dtype=torch.float64, device="cuda", generator=g)
    A = fm.Flomat.alloc(m, w); piv = fm.Pivots(w)
    reps = 20
    tot = 0.0
    for rep in range(reps + 2):
        lib.fm_copy2d(A.a, m, t.data_ptr(), m, m, w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.fm_getrf(m, w, A.a, m, piv.ptr, piv.info_ptr); e1.record(); torch.cuda.synchronize()
        if rep >= 2: tot += e0.elapsed_time(e1)
    out[m] = round(tot / reps * 1e3, 1)  # us (includes zero_info + permute of the panel's own columns, ~6 us)
print(json.dumps(out))
