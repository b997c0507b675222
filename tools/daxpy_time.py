import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
n = 8192
a = np.asfortranarray(np.random.default_rng(1).standard_normal((n, 256))); b = a.copy(order="F"); want = b + 2.0 * a
pa, pb = a.ctypes.data, b.ctypes.data
lib.cblas_daxpy(n, 2.0, pa, 1, pb, 1); b[:, 0] -= 2.0 * a[:, 0]
t0 = time.perf_counter()
for j in range(256): lib.cblas_daxpy(n, 2.0, pa + 8 * j * n, 1, pb + 8 * j * n, 1)
dt = time.perf_counter() - t0
print("us per daxpy(n=8192, host):", round(dt / 256 * 1e6, 1), "exact:", np.array_equal(b, want))
from sci_b200.racket_front import RacketFrontEnd
f = RacketFrontEnd(); x = np.asfortranarray(np.random.default_rng(2).standard_normal((2048, 2048)) / 45.0)
f.expm(x[:256, :256].copy(order="F")); t0 = time.perf_counter(); f.expm(x); print("front expm n=2048 s:", round(time.perf_counter() - t0, 3))
