"""One launch each of the kernels whose ncu summaries are kept under profiles/: times n=8192, plus n=8192, getrf n=8192."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
n = 8192
iseed = np.array([1, 2, 3, 5], dtype=np.int32)
A, B = fm.randn(n, n, iseed=iseed), fm.randn(n, n, iseed=iseed)
C = fm.Flomat.alloc(n, n)
for _ in range(2):
    fm.flomat_times(A, B, C, 1.0, 0.0)
    fm.dot_plus(A, B, C)
piv = fm.Pivots(n)
lib.fm_copy2d(C.a, n, A.a, n, n, n)
lib.fm_getrf(n, n, C.a, n, piv.ptr, piv.info_ptr)
fm.sync()
print("ok", piv.info())
