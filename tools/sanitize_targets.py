"""Small instances of every kernel family in one short script (written for `compute-sanitizer --tool memcheck`, which this pool does
not allow; still a quick all-families check against numpy):
products through all three GEMM paths (TMA big / half / split-K, generic fallback with odd leading dimensions), LU with and without
interchanges (cluster panels, diagonal-block kernel, leaf and wavefront solves), getri, Cholesky, expm (LU and Neumann solves, batched),
elementwise / norm / level-1-2 kernels, the staged host copies.  Prints OK when the results also agree with numpy."""
import math, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
rng = np.random.default_rng(7)
def rnd(m, n, scale=1.0): return np.asfortranarray(rng.standard_normal((m, n)) * scale)
def close(got, want, tol=1e-9): assert np.linalg.norm(got - want, 1) <= tol * max(1.0, np.linalg.norm(want, 1)), (np.linalg.norm(got - want, 1), np.linalg.norm(want, 1))
which = set(sys.argv[1:]) or {"gemm", "lu", "expm", "ewise", "host"}
if "gemm" in which:
    for m, n, k in ((256, 256, 256), (300, 200, 520), (129, 257, 64), (640, 384, 1024), (77, 33, 19)):
        a, b, c = rnd(m, k), rnd(k, n), rnd(m, n)
        close(fm.times(fm.matrix(a), fm.matrix(b)).to_numpy(), a @ b)
        C = fm.matrix(c); fm.flomat_times(fm.matrix(a), fm.matrix(b), C, 0.5, -2.0); close(C.to_numpy(), 0.5 * a @ b - 2.0 * c)
    a = rnd(200, 150); close(fm.flomat_times(fm.matrix(a), fm.matrix(a), None, 1.0, 1.0, True, False).to_numpy(), a.T @ a)
if "lu" in which:
    for n in (5, 128, 130, 300, 600, 1100):
        a, b = rnd(n, n), rnd(n, 3)
        close(fm.mldivide(fm.matrix(a), fm.matrix(b)).to_numpy(), np.linalg.solve(a, b), 1e-7)
    a = rnd(260, 260); close(fm.inv(fm.matrix(a)).to_numpy(), np.linalg.inv(a), 1e-7)
    d = rnd(300, 300, 0.001) + np.eye(300)
    D = fm.matrix(d); piv = fm.Pivots(300); viol = fm.Pivots(1)
    assert lib.fm_getrf_nopiv(300, D.a, D.lda, piv.ptr, piv.info_ptr, viol.ptr) == 0
    import scipy.linalg as sl
    lu, p = sl.lu_factor(d); close(D.to_numpy(), lu, 1e-10)
    a = rnd(700, 700); A = fm.matrix(a); piv = fm.Pivots(700)
    assert lib.fm_getrf_chunked(700, A.a, A.lda, piv.ptr, piv.info_ptr, 256) == 0
    lu, p = sl.lu_factor(a); close(A.to_numpy(), lu, 1e-8)
if "expm" in which:
    for n in (3, 64, 200, 520):
        a = rnd(n, n, 1.0 / math.sqrt(n))
        want = sl.expm(a) if "sl" in dir() else __import__("scipy.linalg").linalg.expm(a)
        for mode in (0, 1):
            close(fm.expm(fm.matrix(a), mode).to_numpy(), want, 1e-9)
    import torch
    src = torch.randn(64 * 64 * 20, dtype=torch.float64, device="cuda") / 8.0; dst = torch.empty_like(src)
    fm.expm_batched(src.data_ptr(), 64, 20, dst.data_ptr(), 1, None); fm.sync()
if "ewise" in which:
    a, b = rnd(333, 77), rnd(333, 77)
    A, B = fm.matrix(a), fm.matrix(b)
    assert np.array_equal(fm.plus(A, B).to_numpy(), a + b) and np.array_equal(fm.dot_times(A, B).to_numpy(), a * b)
    assert abs(fm.norm(A, 1) - np.linalg.norm(a, 1)) < 1e-9 and abs(fm.norm(A, "inf") - np.linalg.norm(a, np.inf)) < 1e-9
    close(fm.transpose(A).to_numpy(), a.T)
if "host" in which:
    n = 1024
    a, b = rnd(n, n), rnd(n, n)
    c = np.zeros((n, n), order="F")
    lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, a.ctypes.data, n, b.ctypes.data, n, 1.0, c.ctypes.data, n)
    close(c, a @ b)
    os.environ["FLOMAT_CUDA_LU_PIPELINE_MIN_N"] = "512"; os.environ["FLOMAT_CUDA_LU_PIPELINE_CHUNK"] = "256"
    from ctypes import byref, c_int
    A = a.copy(order="F"); B = b[:, :2].copy(order="F"); ipiv = np.zeros(n, dtype=np.int32); info = c_int(-1)
    lib.dgesv_(byref(c_int(n)), byref(c_int(2)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    close(B, np.linalg.solve(a, b[:, :2]), 1e-6)
fm.sync()
print("OK", sorted(which))
