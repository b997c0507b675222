"""Repeat few-right-hand-side solves (wavefront triangular solves) and check every result: looks for rare ordering bugs in the
flag hand-over between CTAs (developer tool)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
worst = 0.0
for (n, nrhs) in [(2048, 64), (4100, 17), (1000, 200), (8192, 64), (640, 1)]:
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    a = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(n * nrhs, dtype=torch.float64, device="cuda", generator=g)
    A = fm.Flomat.alloc(n, n); lib.fm_copy2d(A.a, n, a.data_ptr(), n, n, n)
    LU = fm.Flomat.alloc(n, n); lib.fm_copy2d(LU.a, n, a.data_ptr(), n, n, n)
    piv = fm.Pivots(n); lib.fm_getrf(n, n, LU.a, n, piv.ptr, piv.info_ptr)
    X = fm.Flomat.alloc(n, nrhs); B = fm.Flomat.alloc(n, nrhs); lib.fm_copy2d(B.a, n, b.data_ptr(), n, n, nrhs)
    na, nb = fm.norm(A, 1), fm.norm(B, 1)
    first = None
    for r in range(reps if n < 8192 else max(20, reps // 5)):
        lib.fm_copy2d(X.a, n, b.data_ptr(), n, n, nrhs)
        assert lib.fm_getrs(n, nrhs, LU.a, n, piv.ptr, X.a, n) == 0
        R = fm.minus(fm.times(A, X), B)
        res = fm.norm(R, 1) / (na * fm.norm(X, 1))
        worst = max(worst, res)
        assert res <= 64 * n * 2.0 ** -52, (n, nrhs, r, res)
        if r % 10 == 0:  # bit-identical run to run
            x = X.to_numpy()
            if first is None: first = x
            else: assert np.array_equal(first, x), (n, nrhs, r)
    print(f"n={n} nrhs={nrhs}: ok", flush=True)
print("worst backward error", worst)
