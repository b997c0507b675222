"""GPU: parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, the reference's
golden vectors, committed fixtures, and size-independent properties at larger sizes.

Tolerances: elementwise ops are bit-exact; dgemm / LU / expm use the north_star gate
||X_gpu - X_ref||_1 / ||X_ref||_1 <= 64 * n * eps  (eps = 2^-52), written as tol(n) below."""
import ctypes
import math
from ctypes import byref, c_double, c_int

import numpy as np
import pytest

from oracle.flomat_oracle import fortran, normwise_rel_err, tol
from parity import gate, gate_inverse, gate_solve

pytestmark = pytest.mark.gpu

EPS13 = 1e-13  # flomat's epsilon (flomat.rkt:160)


def up(fm, x):
    return fm.matrix(np.asfortranarray(x, dtype=np.float64))


def rnd(seed, m, n, scale=1.0):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)) * scale)


# ---------------------------------------------------------------- reference golden vectors through the ABI
def test_golden_gemm_device_and_compat(fm, golden):
    lib = fm.load()
    for case in golden["gemm"]:
        A, B, AB = fortran(case["A"]), fortran(case["B"]), np.array(case["AB"], dtype=float)
        assert np.array_equal(fm.times(up(fm, A), up(fm, B)).to_numpy(), AB), case["cite"]
        # compat tier with HOST pointers, exactly the call flomat.rkt:888 makes (alpha = beta = 1 into zeros)
        C = np.zeros(AB.shape, order="F")
        lib.cblas_dgemm(102, 111, 111, A.shape[0], B.shape[1], A.shape[1], 1.0, A.ctypes.data, A.shape[0], B.ctypes.data,
                        B.shape[0], 1.0, C.ctypes.data, C.shape[0])
        assert np.array_equal(C, AB), case["cite"]


def test_golden_expt_plus_minus_scale(fm, golden):
    A = up(fm, fortran(golden["expt"]["A"]))
    for k, want in golden["expt"]["powers"].items():
        assert np.array_equal(fm.power(A, int(k)).to_numpy(), np.array(want, dtype=float)), k
    g = golden["plus_minus"]
    A, B = up(fm, fortran(g["A"])), up(fm, fortran(g["B"]))
    assert np.array_equal(fm.plus(A, B).to_numpy(), np.array(g["A+B"], dtype=float))
    assert np.array_equal(fm.minus(A, B).to_numpy(), np.array(g["A-B"], dtype=float))
    assert np.array_equal(fm.minus(A).to_numpy(), np.array(g["-A"], dtype=float))
    assert np.array_equal(fm.flomat_minus(A).to_numpy(), np.array(g["-A"], dtype=float))
    s = golden["scale"]
    assert np.array_equal(fm.flomat_scale(s["s"], up(fm, fortran(s["A"]))).to_numpy(), np.array(s["sA"], dtype=float))


def test_golden_solve_inverse_det_plu(fm, golden):
    M, b = fortran(golden["solve"]["M"]), fortran(golden["solve"]["b"])
    x = fm.mldivide(up(fm, M), up(fm, b))
    assert np.max(np.abs(fm.times(up(fm, M), x).to_numpy() - b)) <= EPS13
    M = fortran(golden["inverse"]["M"])
    Mi = fm.inv(up(fm, M))
    assert np.max(np.abs(fm.times(up(fm, M), Mi).to_numpy() - np.eye(2))) <= EPS13
    assert np.max(np.abs(fm.times(Mi, up(fm, M)).to_numpy() - np.eye(2))) <= EPS13
    for case in golden["det"]:
        d = fm.det(up(fm, fortran(case["M"])))
        if case["tol"] == 0.0:
            assert d == case["det"], (case["cite"], d)
        else:
            assert abs(d - case["det"]) < case["tol"], (case["cite"], d)
    M = fortran(golden["plu"]["M"])
    lu = up(fm, M)
    piv, info = fm.flomat_lu_bang(lu)
    assert info == 0
    LU, ip = lu.to_numpy(), piv.to_numpy()
    n = 4
    P = np.eye(n)
    for i in range(n - 1, -1, -1):
        if ip[i] - 1 != i:
            P[[i, ip[i] - 1], :] = P[[ip[i] - 1, i], :]
    assert np.max(np.abs(P @ (np.tril(LU, -1) + np.eye(n)) @ np.triu(LU) - M)) <= EPS13


def test_identity_and_const(fm):
    for n in (1, 2, 3, 17):
        assert np.array_equal(fm.eye(n).to_numpy(), np.eye(n))
    assert np.array_equal(fm.eye(3, 5, 1).to_numpy(), np.eye(3, 5, 1))
    assert np.array_equal(fm.make_flomat(2, 3, 0.0).to_numpy(), np.zeros((2, 3)))
    assert np.array_equal(fm.ones(5, 2).to_numpy(), np.ones((5, 2)))


# ---------------------------------------------------------------- elementwise: bit exact
@pytest.mark.parametrize("m,n", [(1, 1), (2, 3), (7, 5), (64, 64), (129, 33), (1000, 17), (2048, 512)])
def test_ewise_bit_exact(fm, ref, m, n):
    a, b = rnd(1, m, n), rnd(2, m, n)
    A, B = up(fm, a), up(fm, b)
    assert np.array_equal(fm.dot_plus(A, B).to_numpy(), ref.dot_plus(a, b))
    assert np.array_equal(fm.dot_minus(A, B).to_numpy(), ref.dot_minus(a, b))
    assert np.array_equal(fm.dot_times(A, B).to_numpy(), ref.dot_times(a, b))
    assert np.array_equal(fm.dot_div(A, B).to_numpy(), ref.dot_divide(a, b))
    assert np.array_equal(fm.dot_times(A, 3.25).to_numpy(), ref.dot_times(a, 3.25))
    assert np.array_equal(fm.dot_minus(1.5, B).to_numpy(), ref.dot_minus(1.5, b))
    assert np.array_equal(fm.dot_div(A).to_numpy(), ref.dot_divide(a))
    assert np.array_equal(fm.dot_minus(A).to_numpy(), -a)
    assert np.array_equal(fm.plus(A, B).to_numpy(), ref.plus(a, b))
    assert np.array_equal(fm.minus(A, B).to_numpy(), ref.minus(a, b))
    assert np.array_equal(fm.plus(A, B, 2.0, A).to_numpy(), ref.plus(a, b, 2.0, a))
    assert np.array_equal(fm.flomat_scale(-0.5, A).to_numpy(), ref.scale(-0.5, a))
    assert np.array_equal(fm.copy_flomat(A).to_numpy(), a)


def test_ewise_views_and_inplace(fm, ref):
    a, b = rnd(3, 40, 30), rnd(4, 40, 30)
    A, B = up(fm, a), up(fm, b)
    va, vb = A.sub(3, 5, 21, 11), B.sub(3, 5, 21, 11)  # odd offsets: 8-byte aligned only, lda > m
    assert np.array_equal(fm.dot_times(va, vb).to_numpy(), a[3:24, 5:16] * b[3:24, 5:16])
    fm.dot_plus(va, vb, va)  # the `!` form writes through the view
    want = a.copy()
    want[3:24, 5:16] += b[3:24, 5:16]
    assert np.array_equal(A.to_numpy(), want)
    fm.flomat_plus_bang(vb, va)
    want[3:24, 5:16] = b[3:24, 5:16] + want[3:24, 5:16]
    assert np.array_equal(A.to_numpy(), want)


def test_empty_matrices(fm):
    z = fm.zeros(0, 5)
    assert fm.dot_plus(z, z).to_numpy().shape == (0, 5)
    assert fm.times(fm.zeros(3, 0), fm.zeros(0, 4)).to_numpy().tolist() == np.zeros((3, 4)).tolist()
    assert fm.norm(z, 1) == 0.0


@pytest.mark.parametrize("m,n", [(5, 7), (64, 64), (300, 129), (1024, 77)])
def test_norms(fm, ref, m, n):
    a = rnd(5, m, n)
    A = up(fm, a)
    for kind in (1, "inf", "frob", "max"):
        got, want = fm.norm(A, kind), ref.norm(a, kind)
        assert abs(got - want) <= 4 * max(m, n) * 2.0 ** -52 * abs(want), kind
    assert fm.norm(A, "max") == np.abs(a).max()
    v = A.sub(1, 1, m - 2, n - 2)
    assert abs(fm.norm(v, 1) - np.abs(a[1:-1, 1:-1]).sum(axis=0).max()) <= 1e-12 * m


# ---------------------------------------------------------------- dgemm
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (8, 8, 4), (16, 16, 16), (127, 65, 33), (128, 128, 128), (130, 258, 70), (256, 256, 256),
                                   (512, 384, 1000), (1024, 1024, 1024), (100, 2000, 48), (2000, 100, 48)])
def test_dgemm_nn_parity(fm, ref, m, n, k):
    a, b = rnd(10, m, k), rnd(11, k, n)
    got = fm.times(up(fm, a), up(fm, b)).to_numpy()
    want = ref.gemm(a, b)
    assert normwise_rel_err(got, want) <= tol(max(m, n, k))


def test_dgemm_alpha_beta_trans_and_views(fm, ref):
    rng = np.random.default_rng(12)
    for (ta, tb) in [(False, False), (True, False), (False, True), (True, True)]:
        m, n, k = 150, 90, 70
        a = rnd(13, k, m) if ta else rnd(13, m, k)
        b = rnd(14, n, k) if tb else rnd(14, k, n)
        c = rnd(15, m, n)
        C = up(fm, c)
        fm.flomat_times(up(fm, a), up(fm, b), C, alpha=-0.75, beta=0.5, transpose_a=ta, transpose_b=tb)
        want = ref.gemm(a, b, c.copy(order="F"), alpha=-0.75, beta=0.5, ta=ta, tb=tb)
        assert normwise_rel_err(C.to_numpy(), want) <= tol(k), (ta, tb)
    # views: odd row offsets (not 16-byte aligned) and even ones (TMA path with lda > m)
    big_a, big_b = rnd(16, 300, 300), rnd(17, 300, 300)
    A, B = up(fm, big_a), up(fm, big_b)
    for (i, j) in [(1, 3), (2, 4), (16, 32)]:
        va, vb = A.sub(i, j, 200, 180), B.sub(j, i, 180, 150)
        got = fm.times(va, vb).to_numpy()
        want = ref.gemm(np.asfortranarray(big_a[i:i + 200, j:j + 180]), np.asfortranarray(big_b[j:j + 180, i:i + 150]))
        assert normwise_rel_err(got, want) <= tol(200), (i, j)


def test_dgemm_exact_integers_large(fm):
    """Integer-valued inputs make every partial sum exact, so any summation order must give identical results."""
    rng = np.random.default_rng(18)
    a = np.asfortranarray(rng.integers(-8, 9, (640, 520)).astype(float))
    b = np.asfortranarray(rng.integers(-8, 9, (520, 700)).astype(float))
    assert np.array_equal(fm.times(up(fm, a), up(fm, b)).to_numpy(), a @ b)


@pytest.mark.parametrize("m,n,k", [(256, 256, 256), (384, 200, 1536), (512, 512, 512), (130, 70, 2000)])
def test_dgemm_split_k_alpha_beta_diag(fm, ref, m, n, k):
    """Few tiles and a long inner dimension take the split-K path (partials + fixed-order reduce): same result contract,
    including alpha/beta and the fused `+ c I` of mpoly (expm.rkt:176), and exact on integer data."""
    lib = fm.load()
    a, b, c = rnd(60, m, k), rnd(61, k, n), rnd(62, m, n)
    C = up(fm, c)
    fm.flomat_times(up(fm, a), up(fm, b), C, alpha=-0.75, beta=0.5)
    want = ref.gemm(a, b, c.copy(order="F"), alpha=-0.75, beta=0.5)
    assert normwise_rel_err(C.to_numpy(), want) <= tol(max(m, n, k))
    A, B, D = up(fm, a), up(fm, b), fm.Flomat.alloc(m, n)
    fm._lib.check(lib.fm_dgemm_diag(m, n, k, 1.0, A.a, A.lda, B.a, B.lda, 0.0, D.a, D.lda, 2.5), "fm_dgemm_diag")
    assert normwise_rel_err(D.to_numpy(), a @ b + 2.5 * np.eye(m, n)) <= tol(max(m, n, k))
    rng = np.random.default_rng(63)
    ai = np.asfortranarray(rng.integers(-8, 9, (m, k)).astype(float))
    bi = np.asfortranarray(rng.integers(-8, 9, (k, n)).astype(float))
    assert np.array_equal(fm.times(up(fm, ai), up(fm, bi)).to_numpy(), ai @ bi)


def test_dgemm_linearity_at_full_size(fm):
    """BASELINE config size n=4096..8192: (A1+A2)B == A1 B + A2 B within tolerance; no CPU product needed."""
    n = 4096
    a1, a2, b = rnd(20, n, n), rnd(21, n, n), rnd(22, n, n)
    A1, A2, B = up(fm, a1), up(fm, a2), up(fm, b)
    lhs = fm.times(fm.plus(A1, A2), B)
    rhs = fm.plus(fm.times(A1, B), fm.times(A2, B))
    diff = fm.norm(fm.minus(lhs, rhs), 1)
    assert diff / fm.norm(rhs, 1) <= tol(n)
    # spot-check 64 random entries against float64 dot products
    got = lhs.to_numpy()
    rng = np.random.default_rng(23)
    for _ in range(64):
        i, j = rng.integers(0, n, 2)
        want = float((a1[i, :] + a2[i, :]) @ b[:, j])
        assert abs(got[i, j] - want) <= 64 * n * 2.0 ** -52 * (np.abs(a1[i, :] + a2[i, :]) @ np.abs(b[:, j]))


# ---------------------------------------------------------------- LU family
@pytest.mark.parametrize("n,nrhs", [(1, 1), (2, 1), (5, 3), (16, 16), (17, 2), (64, 7), (130, 5), (256, 256), (500, 33), (1024, 64),
                                    (1500, 10)])
def test_solve_parity(fm, ref, n, nrhs):
    a, b = rnd(30, n, n), rnd(31, n, nrhs)
    x = fm.mldivide(up(fm, a), up(fm, b)).to_numpy()
    want = ref.mldivide(a, b)
    gate_solve(f"mldivide n={n} nrhs={nrhs}", a, b, x, want, n)  # 1 x 64 n eps; cond-scaled only with measured cond (tests/parity.py)
    # the other names of the same path (flomat.rkt:2126-2159): identical bytes
    assert np.array_equal(fm.flomat_solve_many(up(fm, a), up(fm, b)).to_numpy(), x)
    if nrhs <= 3:
        cols = [up(fm, b[:, j:j + 1]) for j in range(nrhs)]
        assert np.array_equal(fm.flomat_left_divide(up(fm, a), cols).to_numpy(), x)
        assert np.array_equal(fm.flomat_solve(up(fm, a), list(b[:, 0])).to_numpy(), fm.mldivide(up(fm, a), up(fm, b[:, :1])).to_numpy())


@pytest.mark.parametrize("n", [8, 33, 64, 130])
def test_lu_fixtures_ipiv_and_factors(fm, fixtures, n):
    m = np.asfortranarray(fixtures[f"lu_in_{n}"])
    lu = up(fm, m)
    piv, info = fm.flomat_lu_bang(lu)
    assert info == 0
    assert np.array_equal(piv.to_numpy(), fixtures[f"oracle_ipiv_{n}"])  # same pivot sequence as the reference's LAPACK
    LU = lu.to_numpy()
    P = np.arange(n)
    for i, p in enumerate(piv.to_numpy()):
        P[[i, p - 1]] = P[[p - 1, i]]
    rec = (np.tril(LU, -1) + np.eye(n)) @ np.triu(LU)
    assert normwise_rel_err(rec, m[P, :]) <= tol(n)
    gate_inverse(f"inv fixture n={n}", m, fm.inv(up(fm, m)).to_numpy(), fixtures[f"oracle_inv_out_{n}"], n)
    d, want = fm.det(up(fm, m)), float(fixtures[f"oracle_det_out_{n}"])
    assert abs(d - want) <= 1e-10 * abs(want)
    rhs = np.asfortranarray(fixtures[f"rhs_in_{n}"])
    x = fm.mldivide(up(fm, m), up(fm, rhs)).to_numpy()
    gate_solve(f"mldivide fixture n={n}", m, rhs, x, fixtures[f"oracle_solve_out_{n}"], n)


def test_lu_rectangular_and_singular(fm, ref):
    for (m, n) in [(40, 20), (20, 40), (300, 130)]:
        a = rnd(32, m, n)
        lu = up(fm, a)
        piv, info = fm.flomat_lu_bang(lu)
        want_lu, want_piv, want_info = ref.lu(a)
        assert info == want_info == 0
        assert np.array_equal(piv.to_numpy(), want_piv[: min(m, n)])
        gate(f"getrf factors {m}x{n}", lu.to_numpy(), want_lu, max(m, n))
    # exactly singular: a zero column -> info = its 1-based index (dgetrf semantics; the reference ignores it)
    a = rnd(33, 6, 6)
    a[:, 2] = 0.0
    piv, info = fm.flomat_lu_bang(up(fm, a))
    assert info == 3 == ref.lu(a)[2]


@pytest.mark.parametrize("n", [6, 130, 300])
def test_singular_matrix_follows_dgesv_and_dgetri(fm, ref, n):
    """An exactly singular A (a zero column): LAPACK's dgesv skips the solve and leaves B untouched, dgetri returns info = i and leaves
    the LU factors; the reference ignores info (flomat.rkt:2135-2142, :1770-1779), so `mldivide` returns the copy of B and `inv` the
    factors.  Same bytes as the oracle, device tier and compat tier."""
    lib = fm.load()
    a, b = rnd(40 + n, n, n), rnd(41 + n, n, 3)
    a[:, n // 2] = 0.0
    want_x = ref.mldivide(a, b)
    assert ref.last_info == n // 2 + 1 and np.array_equal(want_x, b)
    assert np.array_equal(fm.mldivide(up(fm, a), up(fm, b)).to_numpy(), b)
    want_inv = ref.inv(a)
    assert ref.last_info == n // 2 + 1
    got_inv = fm.inv(up(fm, a)).to_numpy()  # the LU factors: same pivots, so the same matrix up to GEMM summation order
    assert np.isfinite(got_inv).all()
    gate(f"inv(singular) = LU factors n={n}", got_inv, want_inv, n)
    A, B = a.copy(order="F"), b.copy(order="F")
    ipiv, info = np.zeros(n, dtype=np.int32), c_int(-99)
    lib.dgesv_(byref(c_int(n)), byref(c_int(3)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    assert info.value == n // 2 + 1 and np.array_equal(B, b)
    lu_before = A.copy(order="F")
    lib.dgetri_(byref(c_int(n)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, None, byref(c_int(n * n)), byref(info))
    assert info.value == n // 2 + 1 and np.array_equal(A, lu_before)


def test_dgetri_singular_not_masked_by_overflow(fm):
    """a zero pivot behind an overflowing diagonal product (inf * 0 = NaN != 0) must still give info = i."""
    lib = fm.load()
    n = 400
    u = np.asfortranarray(np.triu(rnd(44, n, n)) + np.diag(np.full(n, 1e3)))  # prod of the first 120 diagonal entries overflows
    u[300, 300] = 0.0
    ipiv, info = np.arange(1, n + 1, dtype=np.int32), c_int(-99)
    before = u.copy(order="F")
    lib.dgetri_(byref(c_int(n)), u.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, None, byref(c_int(n * n)), byref(info))
    assert info.value == 301 and np.array_equal(u, before)


def test_compat_lapack_symbols_host_pointers(fm, ref):
    """dgesv_/dgetrf_/dgetri_/dlange_ called exactly as the Racket FFI calls them: scalars by reference, host buffers."""
    lib = fm.load()
    n, nrhs = 50, 3
    a, b = rnd(34, n, n), rnd(35, n, nrhs)
    A, B = a.copy(order="F"), b.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    info = c_int(-99)
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)),
               byref(info))
    assert info.value == 0
    gate_solve("dgesv_(host)", a, b, B, ref.mldivide(a, b), n)
    want_lu, want_piv, _ = ref.lu(a)
    assert np.array_equal(ipiv, want_piv)
    # inverse: dgetrf_ + dgetri_ (flomat.rkt:1770-1779)
    A = a.copy(order="F")
    lib.dgetrf_(byref(c_int(n)), byref(c_int(n)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, byref(info))
    assert info.value == 0
    work = np.zeros(n * n)
    lib.dgetri_(byref(c_int(n)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, work.ctypes.data, byref(c_int(n * n)), byref(info))
    assert info.value == 0
    gate_inverse("dgetri_(host)", a, A, ref.inv(a), n)
    got = lib.dlange_(b"1", byref(c_int(n)), byref(c_int(n)), a.ctypes.data, byref(c_int(n)), None)
    assert abs(got - ref.norm(a, 1)) <= 1e-12 * got
    # daxpy / dscal / dcopy incl. the incX = 0 broadcast make-flomat relies on (flomat.rkt:544)
    x, y = rnd(36, 100, 1).ravel().copy(), rnd(37, 100, 1).ravel().copy()
    want = y + 2.0 * x
    lib.cblas_daxpy(100, 2.0, x.ctypes.data, 1, y.ctypes.data, 1)
    assert np.array_equal(y, want)
    lib.cblas_dscal(100, -3.0, y.ctypes.data, 1)
    assert np.array_equal(y, -3.0 * want)
    one = np.array([7.5])
    lib.cblas_dcopy(100, one.ctypes.data, 0, y.ctypes.data, 1)
    assert np.array_equal(y, np.full(100, 7.5))


def test_solve_roundtrip_at_config_size(fm):
    """BASELINE config 3 shape at n=4096, 64 rhs: residual property (the oracle comparison at the full n = 8192 is
    tests/test_gpu_full_size_oracle.py)."""
    n, nrhs = 4096, 64
    a, b = rnd(38, n, n), rnd(39, n, nrhs)
    A, B = up(fm, a), up(fm, b)
    X = fm.mldivide(A, B)
    R = fm.minus(fm.times(A, X), B)
    assert fm.norm(R, 1) / (fm.norm(A, 1) * fm.norm(X, 1)) <= tol(n)


@pytest.mark.parametrize("n,nrhs", [(8192, 64), (9100, 8)])
def test_solve_roundtrip_full_size_on_device(fm, n, nrhs):
    """BASELINE config 3 at its full size (n = 8192, 64 right-hand sides) and one size past 8192 rows (the 1024-rows-per-CTA
    panel configuration): inputs from LAPACK's generator on the device, backward error ||AX - B||_1 / (||A||_1 ||X||_1)."""
    iseed = np.array([n % 4096, 7, 11, 13], dtype=np.int32)
    A, B = fm.randn(n, n, iseed=iseed), fm.randn(n, nrhs, iseed=iseed)
    X = fm.mldivide(A, B)
    R = fm.minus(fm.times(A, X), B)
    assert fm.norm(R, 1) / (fm.norm(A, 1) * fm.norm(X, 1)) <= tol(n)
    P, L, U = fm.flomat_plu(A)  # A = P L U at full size: ||P L U - A||_1 / ||A||_1
    R2 = fm.minus(fm.times(P, fm.times(L, U)), A)
    assert fm.norm(R2, 1) / fm.norm(A, 1) <= tol(n)


def test_lu_tall_panel_fallback(fm, ref, monkeypatch):
    """Panels of more than 16384 rows take an unblocked fallback (explicit interchanges, one IDAMAX per column).  The test
    hook FLOMAT_LU_TALL_MIN sends small panels down that path: same pivots and factors as the oracle, same solve."""
    monkeypatch.setenv("FLOMAT_LU_TALL_MIN", "200")  # panels of > 200 rows use the fallback, the last ones the cluster kernel
    for m, n in [(700, 700), (900, 300), (300, 900)]:
        a = rnd(90 + m, m, n)
        lu = up(fm, a)
        piv, info = fm.flomat_lu_bang(lu)
        want_lu, want_piv, _ = ref.lu(a)
        assert info == 0 and np.array_equal(piv.to_numpy(), want_piv[: min(m, n)])
        gate(f"getrf (tall fallback) {m}x{n}", lu.to_numpy(), want_lu, max(m, n))
    a, b = rnd(97, 600, 600), rnd(98, 600, 5)
    x = fm.mldivide(up(fm, a), up(fm, b)).to_numpy()
    assert np.abs(a @ x - b).sum(axis=0).max() / (np.abs(a).sum(axis=0).max() * np.abs(x).sum(axis=0).max()) <= tol(600)


@pytest.mark.parametrize("n", [130, 257, 384, 700, 1153, 1500, 4096])
def test_inverse_parity_and_property(fm, ref, n):
    """`inv` = dgetrf + dgetri (flomat.rkt:1770-1781): structured L^-1, U sweep, column scatter.  Diagonally shifted random
    matrices keep cond(A) small, so ||A inv(A) - I||_1 and the distance to the oracle's inverse stay inside 64 n eps."""
    a = rnd(70 + n, n, n) + np.sqrt(n) * 2.0 * np.eye(n)
    A = up(fm, a)
    Ai = fm.inv(A)
    R = fm.minus(fm.times(A, Ai), fm.eye(n))
    assert fm.norm(R, 1) <= tol(n)
    if n <= 1500:
        gate(f"inv (shifted) n={n}", Ai.to_numpy(), ref.inv(a), n)


@pytest.mark.parametrize("n", [200, 640, 1000, 2304])
def test_inverse_general_matrix_with_interchanges(fm, ref, n):
    """Unshifted random matrices: every panel interchanges rows, so the column scatter of dgetri and the ragged levels of the
    batched structured inverse of L are exercised.  Residual ||A inv(A) - I||_1 <= 64 n eps ||A||_1 ||inv(A)||_1 (backward-stable bound)."""
    a = rnd(300 + n, n, n)
    A = up(fm, a)
    Ai = fm.inv(A)
    R = fm.minus(fm.times(A, Ai), fm.eye(n))
    assert fm.norm(R, 1) <= tol(n) * fm.norm(A, 1) * fm.norm(Ai, 1)
    if n <= 1000:
        gate_inverse(f"inv (general) n={n}", a, Ai.to_numpy(), ref.inv(a), n)


@pytest.mark.parametrize("n", [1, 7, 16, 33, 128, 130, 257, 512, 1000, 2304, 6144, 6401])
def test_lu_without_interchanges_is_dgetrf_on_dominant_matrices(fm, ref, n):
    """fm_getrf_nopiv (the solve inside expm): on a strictly column diagonally dominant matrix partial pivoting never interchanges, so
    the verified no-interchange LU must reproduce the oracle's dgetrf -- ipiv = 1..n, same factors within 64 n eps -- and report no
    violation; on a general matrix it must raise the violation flag instead of returning wrong factors."""
    a = rnd(500 + n, n, n)
    a = np.asfortranarray(a / (np.abs(a).sum(axis=0) * 1.5) + np.eye(n))  # off-diagonal column sums <= 2/3 < diagonal
    lu = up(fm, a)
    piv, info, violated = fm.flomat_lu_nopivot_bang(lu)
    want_lu, want_piv, want_info = ref.lu(a)
    assert not violated and info == want_info == 0
    assert np.array_equal(piv.to_numpy(), np.arange(1, n + 1)) and np.array_equal(want_piv[:n], np.arange(1, n + 1))
    gate(f"getrf_nopiv n={n}", lu.to_numpy(), want_lu, n)
    b = rnd(600 + n, n, n)  # a general matrix: pivoting needed
    if n > 1:
        _piv, _info, violated = fm.flomat_lu_nopivot_bang(up(fm, b))
        assert violated
    z = np.asfortranarray(np.eye(n)); z[n // 2, n // 2] = 0.0  # exact zero pivot
    _piv, _info, violated = fm.flomat_lu_nopivot_bang(up(fm, z))
    assert violated


def test_expm_falls_back_to_pivoting_when_the_check_fails(fm, ref):
    """The expm pipeline tries the no-interchange LU first and repeats itself with the pivoting LU when the check fails.  A finite
    nonzero input can never fail it (the Pade denominator is dominant), so the fallback is exercised through the NaN case the
    reference produces for the zero matrix (expm.rkt:205: A / 2^-inf), and both paths must agree on ordinary inputs."""
    assert np.isnan(fm.expm(fm.zeros(5, 5)).to_numpy()).all()
    a = rnd(77, 300, 300, 1.0 / math.sqrt(300))
    E = fm.expm(up(fm, a), fm.EXPM_MIN_PRODUCTS).to_numpy()
    assert normwise_rel_err(E, ref.expm(a)) <= tol(300)


@pytest.mark.parametrize("n", [2, 7, 64, 130, 512, 1000, 1024, 1025])
def test_expm_small_matrices_minimal_work_mode_vs_oracle(fm, ref, n):
    """EXPM_MIN_PRODUCTS on ONE matrix of n <= 1024 forms den^-1 num by the factored Neumann product (I+E)(I+E^2)...(I+E^16), E = I - den,
    ||E||_1 <= 0.282 checked on the device (expm.cu) instead of an LU -- nine products, no latency chain.  Against the oracle (the
    reference's dgesv path) at 1x the gate; n = 1025 is the first size back on the LU."""
    a = rnd(900 + n, n, n, 1.0 / math.sqrt(n))
    E, s = fm.expm(up(fm, a), fm.EXPM_MIN_PRODUCTS, return_s=True)
    assert s == ref.scaling_exponent(a)
    gate(f"expm min-products n={n}", E.to_numpy(), ref.expm(a), n)
    a = rnd(950 + n, n, n, 3.0)  # a larger norm: more squarings amplify whatever the solve left
    gate(f"expm min-products n={n}, norm ~ {3 * math.sqrt(n):.0f}", fm.expm(up(fm, a), fm.EXPM_MIN_PRODUCTS).to_numpy(), ref.expm(a), n)


def test_expm_neumann_solve_against_the_lu_solve():
    """The same call with the Neumann solve switched off (LU), and with its dominance bound lowered so that the device-side check
    fails and the pipeline repeats itself with the pivoting LU (fresh processes: the switches are read once)."""
    import os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, math, numpy as np
sys.path.insert(0, {root!r})
import sci_b200 as fm
fm.init(0)
a = np.asfortranarray(np.random.default_rng(4242).standard_normal((512, 512)) / math.sqrt(512))
np.save(sys.argv[1], fm.expm(fm.matrix(a), fm.EXPM_MIN_PRODUCTS).to_numpy())
"""
    outs = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, env in (("neumann", {}), ("lu", {"FLOMAT_EXPM_NEUMANN_MAX_N": "0"}), ("fallback", {"FLOMAT_EXPM_NEUMANN_BOUND": "0.001"}),
                         ("pivot", {"FLOMAT_EXPM_PIVOT": "1"})):
            f = os.path.join(d, tag + ".npy")
            subprocess.run([sys.executable, "-c", code, f], check=True, env={**os.environ, **env}, timeout=600)
            outs[tag] = np.load(f)
    gate("expm: Neumann solve vs LU solve", outs["neumann"], outs["lu"], 512)
    assert np.array_equal(outs["fallback"], outs["pivot"])  # the repeated pass IS the pivoting pipeline


def test_in_place_sum_family(fm):
    """plus! / minus! / flomat+! / flomat-! / flomat-scale! / constant*flomat+flomat(!) (flomat.rkt:3148-3178, :821-834, :1032, :805-819):
    in place, bit-exact (one rounding per element, alpha*x + y for a general alpha as daxpy does it), views included -- and the
    reference's `minus!` quirk: a matrix argument is subtracted INTO (`flomat-! B A` is B := B - A), A comes back unchanged."""
    a, b, c = rnd(61, 37, 21), rnd(62, 37, 21), rnd(63, 37, 21)
    A, B, C = up(fm, a), up(fm, b), up(fm, c)
    R = fm.plus_bang(A, B, 2.5, C)
    assert R is A and np.array_equal(A.to_numpy(), ((a + b) + 2.5) + c) and np.array_equal(B.to_numpy(), b)
    A = up(fm, a)
    assert np.array_equal(fm.minus_bang(A).to_numpy(), -a)
    A = up(fm, a)
    R = fm.minus_bang(A, 0.75)
    assert R is A and np.array_equal(A.to_numpy(), a - 0.75)
    A, B = up(fm, a), up(fm, b)
    R = fm.minus_bang(A, B)                       # the quirk
    assert R is A and np.array_equal(A.to_numpy(), a) and np.array_equal(B.to_numpy(), b - a)
    A, B = up(fm, a), up(fm, b)
    assert fm.flomat_minus_bang(A, B) is A and np.array_equal(A.to_numpy(), a - b) and np.array_equal(B.to_numpy(), b)
    A = up(fm, a)
    assert fm.flomat_scale_bang(-3.0, A) is A and np.array_equal(A.to_numpy(), -3.0 * a)
    A, B = up(fm, a), up(fm, b)
    assert fm.constant_times_flomat_plus_flomat_bang(0.3, A, B) is B
    got, want = B.to_numpy(), 0.3 * a + b
    assert np.all(np.abs(got - want) <= np.spacing(np.abs(want)))  # a general alpha: fused or unfused multiply-add, one ulp apart at most
    A, B = up(fm, a), up(fm, b)
    out = fm.constant_times_flomat_plus_flomat(-1.0, A, B)
    assert np.array_equal(out.to_numpy(), b - a) and np.array_equal(B.to_numpy(), b)
    out = fm.constant_times_flomat_plus_flomat(2.0, A, B)
    assert np.array_equal(out.to_numpy(), 2.0 * a + b) and np.array_equal(B.to_numpy(), b)  # (2x is exact: fused or not, the same)
    # on a view (lda > m): only the view changes
    big = rnd(64, 50, 30)
    Bg = up(fm, big)
    V = Bg.sub(5, 3, 37, 21)
    fm.plus_bang(V, up(fm, a))
    want = big.copy()
    want[5:42, 3:24] += a
    assert np.array_equal(Bg.to_numpy(), want)
    with pytest.raises(ValueError):
        fm.plus_bang(up(fm, a), up(fm, rnd(65, 3, 3)))


def test_named_forms_of_the_lu_norm_and_expm_helpers(fm, ref):
    """The reference's other names on the path: flomat-inverse(!) (:1770-1782), flomat-determinant(-from-plu) (:1837-1848), pivots-sign
    (:1487), flomat-norm and its named forms (:1328-1352), flomat-identity / -zeros / -ones (:548-571), flomat*! (:898), and expm's two
    helpers find-scaling-exponent / repeated-squaring (expm.rkt:209-215) -- the same answers as the front-door functions."""
    n = 96
    a, b = rnd(71, n, n), rnd(72, n, n)
    A = up(fm, a)
    gate_inverse("flomat-inverse", a, fm.flomat_inverse(A).to_numpy(), ref.inv(a), n)
    assert np.array_equal(A.to_numpy(), a)
    W = up(fm, a)
    assert fm.flomat_inverse_bang(W) is W and np.array_equal(W.to_numpy(), fm.inv(A).to_numpy())
    LU = up(fm, a)
    piv = fm.flomat_lu_bang(LU)
    piv = piv[0] if isinstance(piv, tuple) else piv
    d = fm.flomat_determinant_from_plu(LU, piv)
    assert d == fm.det(A) == fm.flomat_determinant(A) and math.isclose(d, ref.det(a), rel_tol=1e-10)
    assert fm.pivots_sign(piv) in (1, -1) and fm.pivots_sign(piv) == piv.sign()
    for kind, f in ((1, fm.flomat_norm1), ("inf", fm.flomat_norm_inf), ("frob", fm.flomat_norm_frob), ("max", fm.flomat_norm_max)):
        assert f(A) == fm.norm(A, kind) == fm.flomat_norm(A, kind)
    assert fm.flomat_norm(A) == fm.norm(A, "frob")
    assert np.array_equal(fm.flomat_identity(5).to_numpy(), np.eye(5)) and np.array_equal(fm.flomat_zeros(2, 3).to_numpy(), np.zeros((2, 3)))
    assert np.array_equal(fm.flomat_ones(3, 2).to_numpy(), np.ones((3, 2)))
    C = fm.zeros(n, n)
    assert fm.flomat_times_bang(A, up(fm, b), C) is C
    gate("flomat*!", C.to_numpy(), ref.times(a, b), n)
    assert fm.find_scaling_exponent(A) == ref.scaling_exponent(a)
    assert fm.find_scaling_exponent(fm.zeros(3, 3)) == -math.inf
    x = rnd(73, n, n, 0.02)
    X = up(fm, x)
    gate("repeated-squaring", fm.repeated_squaring(X, 3).to_numpy(), np.linalg.matrix_power(x, 8), n)
    assert fm.repeated_squaring(X, 0) is X and fm.repeated_squaring(X, -2) is X


# ---------------------------------------------------------------- expm
@pytest.mark.parametrize("mode", [0, 1])
def test_expm_closed_forms_and_quirks(fm, mode):
    E = fm.expm(up(fm, [[1, 0], [0, 2]]), mode).to_numpy()
    assert normwise_rel_err(E, np.diag([math.e, math.e ** 2])) <= tol(2)
    t = 0.7
    E = fm.expm(up(fm, [[0, -t], [t, 0]]), mode).to_numpy()
    assert normwise_rel_err(E, [[math.cos(t), -math.sin(t)], [math.sin(t), math.cos(t)]]) <= tol(2)
    N = np.array([[0, 1, 2], [0, 0, 3], [0, 0, 0]], dtype=float)
    assert normwise_rel_err(fm.expm(up(fm, N), mode).to_numpy(), np.eye(3) + N + N @ N / 2) <= tol(3)
    # s <= 0 is reproduced by default, fixed with EXPM_CLAMP_S
    q = np.diag([0.1, 0.05])
    E, s = fm.expm(up(fm, q), mode, return_s=True)
    assert s == -2.0
    assert normwise_rel_err(E.to_numpy(), np.diag([math.exp(0.4), math.exp(0.2)])) <= 1e-14
    E = fm.expm(up(fm, q), mode | fm.EXPM_CLAMP_S).to_numpy()
    assert normwise_rel_err(E, np.diag([math.exp(0.1), math.exp(0.05)])) <= 1e-14
    with pytest.raises(fm.FlomatCudaError):  # F5-iii: the reference would spin forever
        fm.expm(up(fm, [[np.inf, 0], [0, 1]]), mode)


@pytest.mark.parametrize("n", [8, 33, 64, 130])
@pytest.mark.parametrize("mode", [0, 1])
def test_expm_fixtures(fm, fixtures, n, mode):
    a = np.asfortranarray(fixtures[f"expm_in_{n}"])
    E, s = fm.expm(up(fm, a), mode, return_s=True)
    assert s == float(fixtures[f"oracle_expm_s_{n}"])
    assert normwise_rel_err(E.to_numpy(), fixtures[f"oracle_expm_out_{n}"]) <= tol(n)


@pytest.mark.parametrize("n", [256, 512])
def test_expm_parity_config_sizes(fm, ref, n):
    """BASELINE configs 2 (n=512) and 5 (n=256 items): randn/sqrt(n) inputs of SURVEY 8d."""
    a = rnd(2001, n, n, 1.0 / math.sqrt(n))
    want = ref.expm(a)
    for mode in (0, 1):
        E, s = fm.expm(up(fm, a), mode, return_s=True)
        assert s == ref.scaling_exponent(a)
        assert normwise_rel_err(E.to_numpy(), want) <= tol(n), mode


def test_expm_batched(fm, ref):
    n, batch = 64, 9
    rng = np.random.default_rng(5001)
    items = [np.asfortranarray(rng.standard_normal((n, n)) / 8.0 * (0.3 if z == 4 else 1.0 + 0.5 * (z % 3))) for z in range(batch)]
    stacked = np.concatenate([x.ravel(order="F") for x in items])
    lib = fm.load()
    src = fm.Flomat.from_numpy(stacked.reshape(-1, 1))
    dst = fm.Flomat.alloc(n * n * batch, 1)
    s_out = np.zeros(batch)
    fm.expm_batched(src.a, n, batch, dst.a, 0, s_out)
    got = dst.to_numpy().ravel()
    assert len(set(s_out.tolist())) > 1  # items with different squaring counts in one batch
    for z in range(batch):
        assert s_out[z] == ref.scaling_exponent(items[z])
        Ez = got[z * n * n:(z + 1) * n * n].reshape(n, n, order="F")
        assert normwise_rel_err(Ez, ref.expm(items[z])) <= tol(n), z


@pytest.mark.parametrize("n,batch", [(33, 5), (129, 3), (256, 6)])
@pytest.mark.parametrize("mode", [0, 1])
def test_expm_batched_odd_sizes_and_modes(fm, ref, n, batch, mode):
    """Odd n: the workspaces carry a padding row (ld = n + 1) that the flattened polynomial-term and scaling kernels walk over;
    n = 256 is the item size of BASELINE config 5.  Both product orders, uniform and mixed squaring counts."""
    rng = np.random.default_rng(7000 + n)
    items = [np.asfortranarray(rng.standard_normal((n, n)) / math.sqrt(n) * (1.0 if z % 2 == 0 or n == 256 else 3.0)) for z in range(batch)]
    stacked = np.concatenate([x.ravel(order="F") for x in items])
    src = fm.Flomat.from_numpy(stacked.reshape(-1, 1))
    dst = fm.Flomat.alloc(n * n * batch, 1)
    s_out = np.zeros(batch)
    fm.expm_batched(src.a, n, batch, dst.a, mode, s_out)
    got = dst.to_numpy().ravel()
    for z in range(batch):
        assert s_out[z] == ref.scaling_exponent(items[z])
        Ez = got[z * n * n:(z + 1) * n * n].reshape(n, n, order="F")
        assert normwise_rel_err(Ez, ref.expm(items[z])) <= tol(n), z


def test_expm_semigroup_property_large(fm):
    """exp(A)^2 == exp(2A) at n=1024 (size-independent property; no CPU expm needed)."""
    n = 1024
    a = rnd(6003, n, n, 1.0 / math.sqrt(n))
    E1 = fm.expm(up(fm, a))
    E2 = fm.expm(up(fm, 2.0 * a))
    sq = fm.times(E1, E1)
    assert fm.norm(fm.minus(sq, E2), 1) / fm.norm(E2, 1) <= tol(n)


def test_expm_batched_config5_shard(fm, ref):
    """BASELINE config 5, one GPU's shard at full size: 512 items of 256 x 256 (A_i = N(0,1)/16) in one call.  Sampled items are
    checked against the oracle's expm of the same bytes, every item against exp(A_i) exp(-A_i) == I summed over the batch."""
    n, batch = 256, 512
    rng = np.random.default_rng(5001)
    host = rng.standard_normal(n * n * batch) / 16.0
    src = fm.Flomat.from_numpy(host.reshape(-1, 1))
    neg = fm.dot_times(src, -1.0)
    dst, dstn = fm.Flomat.alloc(n * n * batch, 1), fm.Flomat.alloc(n * n * batch, 1)
    s_out = np.zeros(batch)
    fm.expm_batched(src.a, n, batch, dst.a, fm.EXPM_MIN_PRODUCTS, s_out)
    fm.expm_batched(neg.a, n, batch, dstn.a, fm.EXPM_MIN_PRODUCTS)
    got = dst.to_numpy().ravel()
    for z in (0, 1, 255, 511):
        item = host[z * n * n:(z + 1) * n * n].reshape(n, n, order="F")
        assert s_out[z] == ref.scaling_exponent(item)
        Ez = got[z * n * n:(z + 1) * n * n].reshape(n, n, order="F")
        assert normwise_rel_err(Ez, ref.expm(np.asfortranarray(item))) <= tol(n), z
    # all 512 items: P_z = exp(A_z) exp(-A_z) by one batched product, then || [P_0 .. P_511] - [I .. I] ||_1 over the n x (n*batch) strip
    lib = fm.load()
    prod = fm.Flomat.alloc(n, n * batch)
    rc = lib.fm_dgemm_batched(n, n, n, 1.0, dst.a, n, n * n, dstn.a, n, n * n, 0.0, prod.a, n, n * n, -1.0, batch)  # diag: P_z - I
    assert rc == 0, lib.fm_last_error()
    e1 = fm.norm(fm.Flomat(n, n * batch, dst.a, n, dst._buf), 1)
    assert fm.norm(prod, 1) <= tol(n) * e1 * e1


def test_times_full_size_freivalds_and_linearity(fm):
    """BASELINE config 4 at its full size on one GPU (n = 16384, 3 x 2 GiB, no CPU product possible): inputs from LAPACK's
    generator on the device; Freivalds' check C x == A (B x) for random x and linearity in B, both normwise within 64 n eps."""
    n = 16384
    A, B = fm.randn(n, n, iseed=np.array([41, 1, 2, 3], dtype=np.int32)), fm.randn(n, n, iseed=np.array([42, 1, 2, 3], dtype=np.int32))
    C = fm.times(A, B)
    scale = fm.norm(A, 1) * fm.norm(B, 1)
    for k in range(3):
        x = fm.randn(n, 1, iseed=np.array([43 + k, 5, 6, 7], dtype=np.int32))
        lhs = fm.flomat_times_vector(C, x)
        rhs = fm.flomat_times_vector(A, fm.flomat_times_vector(B, x))
        assert fm.norm(fm.minus(lhs, rhs), 1) / (scale * fm.norm(x, 1)) <= tol(n)
    # (A)(2B) == 2 (AB) exactly (scaling by 2 commutes with every rounding), a bit-level check of the whole 16384^2 result
    C2 = fm.times(A, fm.dot_times(B, 2.0))
    assert fm.norm(fm.minus(C2, fm.dot_times(C, 2.0)), "max") == 0.0


def test_expm_full_size_properties(fm):
    """The metric's expm size (n = 8192, A = N(0,1)/sqrt(n), s = 8) on device-generated input: exp(A) exp(-A) == I and
    exp(A)^2 == exp(2A), normwise within 64 n eps; both product orders of the polynomial agree."""
    n = 8192
    A = fm.dot_times(fm.randn(n, n, iseed=np.array([61, 3, 5, 7], dtype=np.int32)), 1.0 / math.sqrt(n))
    E, s = fm.expm(A, fm.EXPM_MIN_PRODUCTS, return_s=True)
    assert s == 8.0
    Em = fm.expm(fm.dot_times(A, -1.0), fm.EXPM_MIN_PRODUCTS)
    R = fm.minus(fm.times(E, Em), fm.eye(n))
    assert fm.norm(R, 1) <= tol(n) * fm.norm(E, 1) * fm.norm(Em, 1)
    del Em, R
    E2 = fm.expm(fm.dot_times(A, 2.0), fm.EXPM_MIN_PRODUCTS)
    assert fm.norm(fm.minus(fm.times(E, E), E2), 1) / fm.norm(E2, 1) <= tol(n)
    del E2
    Eref = fm.expm(A, fm.EXPM_REFERENCE_ORDER)
    assert fm.norm(fm.minus(E, Eref), 1) / fm.norm(Eref, 1) <= tol(n)


# ---------------------------------------------------------------- multi-GPU (needs >= 2 B200; skipped on a 1-GPU box)
def test_colshard_fused_two_gpus(fm, tmp_path):
    """Column-sharded times with the all-gather fused into the GEMM epilogue (peer stores over NVLink), 2 ranks:
    every rank's assembled C must equal the producer's copy bit for bit and match an independent product."""
    import os
    import socket
    import subprocess
    import sys
    import textwrap

    if fm.load().fm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {root!r})
        import torch, torch.distributed as dist
        import sci_b200 as fm
        from sci_b200.multigpu import ColShardTimes
        rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local); fm.init(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        for n in (1024, 8192):
            cs = ColShardTimes(n, 2, rank)
            for _ in range(3):
                cs.c_tensor().fill_(-7.0); torch.cuda.synchronize(); dist.barrier()
                cs.step_fused(); torch.cuda.synchronize(); dist.barrier()
                assert cs.verify(), ("fused", n)
            cs.c_tensor().fill_(-7.0); cs.step_nccl(); torch.cuda.synchronize(); dist.barrier()
            assert cs.verify(), ("nccl", n)
            cs.close()
        dist.destroy_process_group(); print("ok", rank)
    """))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o
