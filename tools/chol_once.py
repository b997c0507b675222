"""One potrf at n=CHOL_N for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("CHOL_N", "8192"))
M = fm.randn(n, n, iseed=np.array([1, 2, 3, 5], dtype=np.int32))
A = fm.plus(fm.times(M, fm.transpose(M)), fm.dot_times(fm.eye(n), float(n)))
info = fm.Pivots(0)
lib.fm_potrf(b"L", n, A.a, n, info.info_ptr); fm.sync()
print("info", info.info())
