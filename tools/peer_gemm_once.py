"""One column-block product with the fused all-gather epilogue on ONE GPU: the 7 "peers" are other buffers of the same device, so the
epilogue code of the 8-GPU run executes (stores go to local HBM instead of NVLink).  For `ncu --set full` captures of the two epilogues:
default = forwarding warps, FLOMAT_GEMM_NO_COPIER=1 = the consumer warps store to the peers themselves."""
import ctypes, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
n, nloc, npeers = 16384, 2048, 7
A = fm.randn(n, n, iseed=np.array([1, 2, 3, 5], dtype=np.int32)); B = fm.randn(n, nloc, iseed=np.array([2, 2, 3, 5], dtype=np.int32))
C = fm.Flomat.alloc(n, nloc + 0); peers = [fm.Flomat.alloc(n, nloc) for _ in range(npeers)]
arr = (ctypes.c_void_p * npeers)(*[p.a for p in peers])
for _ in range(3):
    fm._lib.check(lib.fm_dgemm_colshard(n, nloc, n, A.a, n, B.a, n, C.a, n, 0, arr, npeers), "fm_dgemm_colshard")
fm.sync()
print("ok", np.array_equal(C.to_numpy()[:, :4], peers[6].to_numpy()[:, :4]))
