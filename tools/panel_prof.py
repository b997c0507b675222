"""Phase breakdown (clock64, warp 0 of CTA 0) of the LU panel kernel: FLOMAT_CUDA_LIB=sci_b200/lib/libflomat_cuda_prof.so."""
import sys, os, ctypes
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
lib.fm_lu_prof.argtypes = [ctypes.c_void_p]
names = {0: "step: mbarrier wait", 1: "step: winner scan + pm loads + bookkeeping", 2: "step: l, next column, thread best", 3: "step: warp argmax + rest of update",
         4: "step: publish (winner STS)", 5: "step: __syncthreads", 6: "step: CTA argmax + push", 8: "leaf: load", 9: "leaf: all column steps", 10: "leaf: store",
         11: "leaf: cluster wait + pending U12", 12: "leaf: U12 (pivot-row loads + solve)", 13: "leaf: rank-16 update"}
buf = (ctypes.c_longlong * 32)()
for m in [int(x) for x in (sys.argv[1:] or ["1024", "4096", "8192"])]:
    w = 128
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    t = torch.randn(m * w, dtype=torch.float64, device="cuda", generator=g)
    A = fm.Flomat.alloc(m, w); piv = fm.Pivots(w)
    for rep in range(2):
        lib.fm_copy2d(A.a, m, t.data_ptr(), m, m, w)
        fm.sync(); lib.fm_lu_prof(buf)
        lib.fm_getrf(m, w, A.a, m, piv.ptr, piv.info_ptr)
        fm.sync(); lib.fm_lu_prof(buf)
    tot = buf[9] + buf[8] + buf[10] + buf[11] + buf[12] + buf[13]
    print(f"--- m={m} w={w}: {tot} cycles in leaves ({tot/1965:.1f} us @1965MHz)")
    for k, nm in names.items():
        per = 128 if k < 8 else 8
        print(f"   {nm:50s} {buf[k]:9d} cyc  ({buf[k]//per:7d} per {'step' if k < 8 else 'leaf'})")
