"""CUDA-event time of fm_getrf_chunked (the factorisation behind the host-operand dgesv_ pipeline) against fm_getrf, device-resident A."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
dev = torch.device("cuda", 0)
def best(fn, reps=4):
    fn(); torch.cuda.synchronize(); b = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
out = {}
for n in (8192, 16384):
    t = torch.randn(n * n, dtype=torch.float64, device=dev)
    LU = fm.Flomat.alloc(n, n); piv = fm.Pivots(n)
    t_copy = best(lambda: lib.fm_copy2d(LU.a, n, t.data_ptr(), n, n, n))
    def plain():
        lib.fm_copy2d(LU.a, n, t.data_ptr(), n, n, n); lib.fm_getrf(n, n, LU.a, n, piv.ptr, piv.info_ptr)
    row = {"fm_getrf": round(best(plain) - t_copy, 3)}
    for cw in (512, 1024, 2048, 4096):
        def chunked():
            lib.fm_copy2d(LU.a, n, t.data_ptr(), n, n, n); rc = lib.fm_getrf_chunked(n, LU.a, n, piv.ptr, piv.info_ptr, cw); assert rc == 0
        row[f"chunk_{cw}"] = round(best(chunked) - t_copy, 3)
    out[f"n{n}"] = row
    del t, LU
print(json.dumps(out))
