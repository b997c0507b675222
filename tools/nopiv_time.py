import sys, os, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
out = {}
for n in (128, 256, 512):
    rng = np.random.default_rng(n); a = rng.standard_normal((n, n)); a = np.asfortranarray(a / (np.abs(a).sum(axis=0) * 1.5) + np.eye(n))
    A = fm.matrix(a); W = fm.Flomat.alloc(n, n); piv = fm.Pivots(n + 4)
    tot = 0.0
    for rep in range(22):
        lib.fm_copy2d(W.a, n, A.a, n, n, n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.fm_getrf_nopiv(n, W.a, n, piv.ptr, piv.info_ptr, piv.ptr + 4 * n); e1.record(); torch.cuda.synchronize()
        if rep >= 2: tot += e0.elapsed_time(e1)
    out[n] = round(tot / 20 * 1e3, 1)
print(json.dumps(out))
