"""Phase breakdown (clock64, thread 0 of each block-row CTA, critical section) of the wavefront triangular solve:
FLOMAT_CUDA_LIB=sci_b200/lib/libflomat_cuda_prof.so python tools/wave_prof.py [n] [nrhs]"""
import sys, os, ctypes
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
lib.fm_lu_prof.argtypes = [ctypes.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nrhs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
names = {17: "off-diag: landed + barrier (x2)", 18: "off-diag: prefetch issue (x2)", 20: "off-diag: barrier + x load issue", 19: "off-diag: update DMMAs (x2)",
         21: "diag: wait_group + barrier (x2)", 22: "diag: cp.async issue (x2)", 23: "diag: solve + updates (x2)", 24: "x stores", 25: "barrier", 26: "fence + flag"}
g = torch.Generator(device="cuda"); g.manual_seed(3)
t = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g)
A = fm.Flomat.alloc(n, n); lib.fm_copy2d(A.a, n, t.data_ptr(), n, n, n)
piv = fm.Pivots(n)
lib.fm_getrf(n, n, A.a, n, piv.ptr, piv.info_ptr)
B = fm.Flomat.alloc(n, nrhs); lib.fm_fill(B.a, n, n, nrhs, 1.0)
buf = (ctypes.c_longlong * 32)()
for rep in range(2):
    fm.sync(); lib.fm_lu_prof(buf)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); lib.fm_getrs(n, nrhs, A.a, n, piv.ptr, B.a, n); e1.record()
    fm.sync(); lib.fm_lu_prof(buf)
cnt = max(buf[27], 1)
print(f"n={n} nrhs={nrhs}: getrs {e0.elapsed_time(e1)*1e3:.0f} us, {cnt} block-row CTAs profiled (both triangles)")
tot = 0
for k, nm in names.items():
    print(f"   {nm:45s} {buf[k]//cnt:8d} cycles per step"); tot += buf[k] // cnt
print(f"   sum {tot} cycles = {tot/1965:.2f} us per step (flag observed -> flag published)")
