"""Single GPU: run fm_dgemm_colshard with a second local buffer as the 'peer' and compare both copies with an independent product."""
import os, sys, ctypes
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = int(os.environ.get("CS_N", "16384")); nloc = n // 2
g = torch.Generator(device="cuda"); g.manual_seed(1)
A = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g)
B = torch.randn(n * nloc, dtype=torch.float64, device="cuda", generator=g)
C = torch.full((n * n,), -7.0, dtype=torch.float64, device="cuda")
P = torch.full((n * n,), -7.0, dtype=torch.float64, device="cuda")
arr = (ctypes.c_void_p * 1)(P.data_ptr())
a = A.view(n, n).t()
for it in range(3):
    rc = lib.fm_dgemm_colshard(n, nloc, n, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n, 0, arr, 1)
    torch.cuda.synchronize()
    for jj in (0, 1, 2, 3, 4, 200):
        want = a @ B[jj * n:(jj + 1) * n]
        for name, buf in (("local", C), ("peer", P)):
            got = buf[jj * n:(jj + 1) * n]
            bad = ((got - want).abs() > 1e-6).nonzero().flatten()
            if bad.numel():
                i = bad[0].item()
                print(f"it{it} col {jj} {name}: bad={bad.numel()} first rows={bad[:12].tolist()} got={got[i].item():.6f} want={want[i].item():.6f}", flush=True)
    print(f"it{it} rc={rc} equal copies: {torch.equal(C[:n*nloc], P[:n*nloc])}", flush=True)
