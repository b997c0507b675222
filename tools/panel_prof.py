"""Phase breakdown (clock64) of the LU panel kernel; needs libflomat_cuda_prof.so (tools/build_prof.sh)."""
import sys, os, ctypes
os.environ["FLOMAT_CUDA_LIB"] = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir, "sci_b200", "lib", "libflomat_cuda_prof.so"))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
lib.fm_lu_prof.argtypes = [ctypes.c_void_p]
names = ["leaf load", "col: pm+swap+l + next candidate/publish/sync/push", "(unused)", "col: mbarrier wait", "col: winner scan", "col: rest of update",
         "leaf store", "swap (cta0)", "cluster barrier 1", "trsm (cta0)", "cluster barrier 2", "stage U", "rest update"]
buf = (ctypes.c_longlong * 32)()
for m, w in [(1024, 128), (8192, 128)]:
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    t = torch.randn(m * w, dtype=torch.float64, device="cuda", generator=g)
    piv = fm.Pivots(128)
    lib.fm_getrf(m, w, t.data_ptr(), m, piv.ptr, piv.info_ptr)
    lib.fm_lu_prof(buf)
    t = torch.randn(m * w, dtype=torch.float64, device="cuda", generator=g)
    lib.fm_getrf(m, w, t.data_ptr(), m, piv.ptr, piv.info_ptr)
    lib.fm_lu_prof(buf)
    tot = sum(buf[:13])
    print(f"--- m={m} w={w}: total {tot} cycles ({tot/1.965e3:.1f} us @1965MHz)")
    for i, nme in enumerate(names):
        per = buf[i] / (w if nme.startswith("col") else max(1, (w + 15) // 16))
        print(f"   {nme:32s} {buf[i]:9d} cyc  ({per:8.0f} per {'column' if nme.startswith('col') else 'leaf'})")
