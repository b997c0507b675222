"""Times LU-style update GEMMs C -= A*B (alpha=-1, beta=1) for a list of m,n,k shapes (developer tool).
FLOMAT_GEMM_NO_HALF=1 disables the 64x128 two-CTA/SM configuration for A/B runs."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
shapes = [tuple(int(x) for x in s.split(",")) for s in sys.argv[1:]] or [(8192, 8192, 128), (8192, 8192, 256), (8192, 8192, 512),
          (4096, 4096, 128), (8064, 128, 128), (4096, 128, 128), (2048, 2048, 128), (8192, 64, 128), (1024, 1024, 1024), (512, 512, 512), (256, 256, 256)]
for (m, n, k) in shapes:
    ld = max(m, k)
    a = fm.Flomat.alloc(ld, max(k, n)); b = fm.Flomat.alloc(ld, n); c = fm.Flomat.alloc(ld, n)
    lib.fm_fill(a.a, ld, ld, max(k, n), 1e-3); lib.fm_fill(b.a, ld, ld, n, 0.5); lib.fm_fill(c.a, ld, ld, n, 0.0)
    def run(): lib.fm_dgemm(111, 111, m, n, k, -1.0, a.a, ld, b.a, ld, 1.0, c.a, ld)
    for _ in range(2): run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"gemm m={m} n={n} k={k}: {best*1e3:.1f} us  {2*m*n*k/best/1e9:.2f} TFLOP/s ({2*m*n*k/best/1e9/37.0*100:.1f}%)", flush=True)
