"""One dgesv_ with pageable host operands under FLOMAT_HOST_LU_TRACE=1 (per-chunk upload / factor / download times on stderr)."""
import os, sys, time
os.environ.setdefault("FLOMAT_HOST_LU_TRACE", "1")
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
from ctypes import byref, c_int
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192; nrhs = 64
rng = np.random.default_rng(n)
a = np.asfortranarray(rng.standard_normal((n, n))); b = np.asfortranarray(rng.standard_normal((n, nrhs)))
ipiv = np.zeros(n, dtype=np.int32); info = c_int(-99)
for rep in range(3):
    A, B = a.copy(order="F"), b.copy(order="F")
    t0 = time.perf_counter()
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
