"""The callers on either side of the dense hot path: transpose / mrdivide / triangle extraction /
pivots->flomat / flomat-plu, dgemv / ddot / dnrm2 / idamax / dswap / row and column sums, the rest of the pointwise
family and LAPACK's dlarnv generator.

CPU part: the oracle restatements against the reference's own rackunit vectors (tests/golden/reference_tests.json) and,
for dlarnv, against the real LAPACK routine inside the OpenBLAS build the oracle binds (bit-exact uniforms and seeds).
GPU part (-m gpu): the CUDA path through the C ABI against the oracle.  Data movement, sqr, sqrt and the uniform
generator are bit-exact; libm-backed maps are compared within 4 ulp (CUDA's documented bound is 2 ulp for these, glibc's
is 1); reductions within n * eps * sum|terms|."""
import ctypes

import numpy as np
import pytest

from oracle.flomat_oracle import Flomat as Oracle, find_openblas, fortran

EPS = 2.0 ** -52


def rnd(seed, m, n, scale=1.0):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)) * scale)


# ------------------------------------------------------------------------------------------ CPU: oracle pinning
def _lapack_dlarnv(idist, iseed, n):
    lib = ctypes.CDLL(find_openblas())
    x = np.zeros(n)
    s = np.array(iseed, dtype=np.int32)
    lib.scipy_dlarnv_(ctypes.byref(ctypes.c_int(idist)), s.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ctypes.c_int(n)),
                      x.ctypes.data_as(ctypes.c_void_p))
    return x, s


@pytest.mark.skipif(find_openblas() is None, reason="no LP64 OpenBLAS with LAPACK in this image")
@pytest.mark.parametrize("idist", [1, 2, 3])
def test_oracle_dlarnv_is_lapack(idist):
    for n in (1, 63, 64, 65, 128, 129, 1000):
        for seed in ([1, 2, 3, 5], [4095, 4095, 4095, 4095], [0, 0, 0, 1], [123, 3001, 77, 2049]):
            xr, sr = _lapack_dlarnv(idist, seed, n)
            x, s2 = Oracle.dlarnv(idist, seed, n)
            assert np.array_equal(s2, sr)
            if idist < 3:
                assert np.array_equal(x, xr)  # exact 48-bit integers scaled by 2^-48
            else:
                assert np.abs(x - xr).max() <= 1e-14


def test_oracle_neighbours_golden(golden, providers):
    for prov in providers:
        F = Oracle(prov)
        for t in golden["gemv"]:
            y = F.times_vector(fortran(t["A"]), fortran(t["X"]), transpose_a=t["transpose_a"])
            assert np.array_equal(y, fortran(t["Y"])), t["cite"]
        for t in golden["column_dot"]:
            assert F.dot(np.array(t["x"], float), np.array(t["y"], float)) == t["dot"], t["cite"]
        for t in golden["column_norm"]:
            assert abs(F.column_norm(np.array(t["x"], float)) - np.sqrt(t["norm_squared"])) <= 1e-15
        a = rnd(11, 7, 7)
        p, l, u = F.plu(a)
        assert np.abs(p @ l @ u - a).max() <= 1e-13
        b = rnd(12, 3, 7)
        assert np.abs(F.mrdivide(b, a) @ a - b).max() <= 1e-11
        assert np.array_equal(F.transpose(rnd(13, 3, 5)), rnd(13, 3, 5).T)


# ------------------------------------------------------------------------------------------ GPU: parity through the C ABI
gpu = pytest.mark.gpu


def up(fm, x):
    return fm.matrix(np.asfortranarray(x, dtype=np.float64))


def ulps(x, ref):
    x, ref = np.asarray(x), np.asarray(ref)
    spacing = np.spacing(np.abs(ref))
    d = np.abs(x - ref) / np.where(spacing > 0, spacing, 1.0)
    return float(np.max(np.where(np.isfinite(ref), d, 0.0))) if ref.size else 0.0


@gpu
@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (33, 65), (64, 64), (1000, 37), (513, 1026)])
def test_transpose_and_triangles_bit_exact(fm, ref, m, n):
    a = rnd(1, m, n)
    A = up(fm, a)
    assert np.array_equal(fm.transpose(A).to_numpy(), ref.transpose(a))
    assert np.array_equal(fm.flomat_extract_upper(A).to_numpy(), ref.extract_upper(a))
    assert np.array_equal(fm.flomat_extract_lower(A).to_numpy(), ref.extract_lower(a))
    assert np.array_equal(fm.flomat_extract_lower_ones(A).to_numpy(), ref.extract_lower(a, ones=True))
    if m > 4 and n > 3:  # a view (lda > m)
        v = A.sub(1, 1, m - 2, n - 2)
        assert np.array_equal(fm.transpose(v).to_numpy(), a[1:m - 1, 1:n - 1].T)


@gpu
@pytest.mark.parametrize("n", [1, 2, 7, 130, 300])
def test_plu_and_pivot_matrix(fm, ref, n):
    a = rnd(20 + n, n, n)
    P, L, U = fm.flomat_plu(up(fm, a))
    lu, ipiv, _ = ref.lu(a)
    Pg = P.to_numpy()
    assert np.array_equal(Pg, ref.pivots_to_matrix(ipiv)), "same pivots -> the same permutation matrix, exactly"
    assert np.array_equal(np.sort(Pg.sum(axis=0)), np.ones(n)) and np.array_equal(np.sort(Pg.sum(axis=1)), np.ones(n))
    rec = Pg @ L.to_numpy() @ U.to_numpy()
    assert np.abs(rec - a).sum(axis=0).max() / np.abs(a).sum(axis=0).max() <= 64 * n * EPS


@gpu
def test_golden_gemv_dot_norm(fm, golden):
    for t in golden["gemv"]:
        y = fm.flomat_times_vector(up(fm, t["A"]), up(fm, t["X"]), transpose_a=t["transpose_a"])
        assert np.array_equal(y.to_numpy(), fortran(t["Y"])), t["cite"]
    for t in golden["column_dot"]:
        assert fm.dot(up(fm, np.array(t["x"], float)), up(fm, np.array(t["y"], float))) == t["dot"], t["cite"]
    for t in golden["column_norm"]:
        assert abs(fm.flcolumn_norm(up(fm, np.array(t["x"], float))) - np.sqrt(t["norm_squared"])) <= 1e-15


@gpu
@pytest.mark.parametrize("m,n", [(1, 1), (3, 5), (257, 129), (2000, 33), (40, 3000)])
def test_gemv_sums_dot_parity(fm, ref, m, n):
    a, x, xt, y0 = rnd(2, m, n), rnd(3, n, 1), rnd(4, m, 1), rnd(5, m, 1)
    A = up(fm, a)
    bound = lambda terms: 4 * max(m, n) * EPS * terms  # noqa: E731
    got = fm.flomat_times_vector(A, up(fm, x)).to_numpy()
    assert np.all(np.abs(got - ref.times_vector(a, x)) <= bound(np.abs(a) @ np.abs(x)) + 1e-300)
    got = fm.flomat_times_vector(A, up(fm, xt), transpose_a=True).to_numpy()
    assert np.all(np.abs(got - ref.times_vector(a, xt, transpose_a=True)) <= bound(np.abs(a.T) @ np.abs(xt)) + 1e-300)
    y = up(fm, y0)
    got = fm.flomat_times_vector(A, up(fm, x), y, alpha=-2.0, beta=0.5).to_numpy()
    want = ref.times_vector(a, x, y0, alpha=-2.0, beta=0.5)
    assert np.all(np.abs(got - want) <= bound(2 * np.abs(a) @ np.abs(x) + np.abs(y0)) + 1e-300)
    assert np.all(np.abs(fm.colsums(A).to_numpy() - ref.colsums(a)) <= bound(np.abs(a).sum(axis=0, keepdims=True)))
    assert np.all(np.abs(fm.rowsums(A).to_numpy() - ref.rowsums(a)) <= bound(np.abs(a).sum(axis=1, keepdims=True)))
    assert abs(fm.dot(up(fm, xt), up(fm, y0)) - ref.dot(xt, y0)) <= 4 * m * EPS * float(np.abs(xt * y0).sum()) + 1e-300
    assert abs(fm.flcolumn_norm(up(fm, xt)) - ref.column_norm(xt)) <= 4 * EPS * ref.column_norm(xt) * max(1, np.log2(m))
    assert fm.flomat_max_abs_index(A) == ref.max_abs_index(a)
    assert fm.flomat_max_abs_value(A) == a[ref.max_abs_index(a)]


@gpu
def test_iamax_ties_views_and_nrm2_range(fm, ref):
    a = np.asfortranarray(np.array([[1.0, -7.0, 2.0], [7.0, 3.0, -7.0]]))  # three entries of magnitude 7: the first wins
    assert fm.flomat_max_abs_index(up(fm, a)) == ref.max_abs_index(a) == (1, 0)
    big = rnd(6, 9, 8)
    v = up(fm, big).sub(2, 1, 5, 6)
    assert v.lda > v.m and fm.flomat_max_abs_index(v) == ref.max_abs_index(np.asfortranarray(big[2:7, 1:7]))
    huge = np.full((5, 1), 1e200)  # the plain sum of squares overflows; dnrm2's scaled form does not
    assert abs(fm.flcolumn_norm(up(fm, huge)) - 1e200 * np.sqrt(5)) <= 1e185
    assert fm.flcolumn_norm(up(fm, np.zeros((4, 1)))) == 0.0


@gpu
def test_swap_rows_columns_and_mrdivide(fm, ref):
    a = rnd(7, 6, 9)
    A = up(fm, a)
    fm.flomat_swap_rows_bang(A, 1, 4)
    a[[1, 4], :] = a[[4, 1], :]
    assert np.array_equal(A.to_numpy(), a)
    fm.flomat_swap_columns_bang(A, 0, 8)
    a[:, [0, 8]] = a[:, [8, 0]]
    assert np.array_equal(A.to_numpy(), a)
    for n, k in [(5, 3), (200, 17)]:
        m_, b_ = rnd(8, n, n) + n * np.eye(n), rnd(9, k, n)
        got = fm.mrdivide(up(fm, b_), up(fm, m_)).to_numpy()
        want = ref.mrdivide(b_, m_)
        assert np.abs(got - want).max() / np.abs(want).max() <= 64 * n * EPS


@gpu
@pytest.mark.parametrize("m,n", [(1, 1), (7, 5), (256, 130), (1001, 3)])
def test_pointwise_family(fm, ref, m, n):
    a = rnd(30, m, n)
    pos = np.abs(a) + 0.25
    A, Pm = up(fm, a), up(fm, pos)
    assert np.array_equal(fm.dot_sqr(A).to_numpy(), ref.pointwise("sqr", a))        # one IEEE multiply
    assert np.array_equal(fm.dot_sqrt(Pm).to_numpy(), ref.pointwise("sqrt", pos))   # correctly rounded on both sides
    for name, f, arg, h in [("sin", fm.dot_sin, A, a), ("cos", fm.dot_cos, A, a), ("tan", fm.dot_tan, A, a),
                            ("log", fm.dot_log, Pm, pos), ("exp", fm.dot_exp, A, a)]:
        assert ulps(f(arg).to_numpy(), ref.pointwise(name, h)) <= 4, name
    assert ulps(fm.dot_expt(Pm, A).to_numpy(), ref.pointwise("expt", pos, a)) <= 4
    assert ulps(fm.dot_expt(Pm, 2.5).to_numpy(), ref.pointwise("expt", pos, 2.5)) <= 4
    assert ulps(fm.dot_expt(2.0, A).to_numpy(), ref.pointwise("expt", np.full_like(a, 2.0), a)) <= 4
    C = up(fm, pos.copy())
    fm.dot_sqrt(C, C)  # the `!` form: in place
    assert np.array_equal(C.to_numpy(), ref.pointwise("sqrt", pos))


@gpu
@pytest.mark.parametrize("idist", [1, 2, 3])
def test_larnv_matches_lapack_stream(fm, ref, idist):
    lib = fm.load()
    for n, seed in [(1, [1, 2, 3, 5]), (130, [123, 3001, 77, 2049]), (100003, [0, 0, 0, 1])]:
        want, seed_after = ref.dlarnv(idist, seed, n)
        iseed = np.array(seed, dtype=np.int32)
        X = fm.Flomat.alloc(n, 1)
        fm._lib.check(lib.fm_larnv(idist, iseed.ctypes.data, n, X.a), "fm_larnv")
        got = X.to_numpy().reshape(-1)
        assert np.array_equal(iseed, seed_after), "the seed advances exactly like DLARNV's"
        if idist < 3:
            assert np.array_equal(got, want)
        else:
            assert np.abs(got - want).max() <= 1e-13
    # two calls continue the stream (what `the-iseed` relies on, flomat.rkt:2818)
    iseed = np.array([1, 2, 3, 5], dtype=np.int32)
    a = fm.rand(10, 7, iseed=iseed).to_numpy().reshape(-1, order="F")
    b = fm.rand(5, 5, iseed=iseed).to_numpy().reshape(-1, order="F")
    want, _ = ref.dlarnv(1, [1, 2, 3, 5], 95)
    assert np.array_equal(np.concatenate([a, b]), want)
    z = fm.randn(300, 300, iseed=np.array([7, 7, 7, 7], dtype=np.int32)).to_numpy()
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02


@gpu
def test_compat_level12_symbols_host_pointers(fm, ref):
    """cblas_dgemv / ddot / dnrm2 / idamax / dswap / dlarnv_ with HOST buffers: what an unmodified ffi binding would pass."""
    lib = fm.load()
    a, x = rnd(40, 6, 4), rnd(41, 4, 1)
    y = np.zeros((6, 1), order="F")
    lib.cblas_dgemv(102, 111, 6, 4, 1.0, a.ctypes.data, 6, x.ctypes.data, 1, 0.0, y.ctypes.data, 1)
    assert np.abs(y - a @ x).max() <= 1e-14
    u, v = rnd(42, 50, 1), rnd(43, 50, 1)
    assert abs(lib.cblas_ddot(50, u.ctypes.data, 1, v.ctypes.data, 1) - float((u.T @ v)[0, 0])) <= 1e-13
    one = np.ones(1)
    assert abs(lib.cblas_ddot(4, a.ctypes.data, 6, one.ctypes.data, 0) - a[0, :].sum()) <= 1e-14  # `unsafe-sum` (flomat.rkt:2074)
    assert abs(lib.cblas_dnrm2(50, u.ctypes.data, 1) - np.linalg.norm(u)) <= 1e-14
    assert lib.cblas_idamax(50, u.ctypes.data, 1) == int(np.argmax(np.abs(u)))
    r1, r2 = a[1, :].copy(), a[3, :].copy()
    lib.cblas_dswap(4, a.ctypes.data + 8, 6, a.ctypes.data + 24, 6)  # rows 1 and 3, stride lda (flomat.rkt:1088)
    assert np.array_equal(a[1, :], r2) and np.array_equal(a[3, :], r1)
    iseed = np.array([1, 2, 3, 5], dtype=np.int32)
    out = np.zeros(77)
    lib.dlarnv_(ctypes.byref(ctypes.c_int(1)), iseed.ctypes.data, ctypes.byref(ctypes.c_int(77)), out.ctypes.data)
    want, seed_after = ref.dlarnv(1, [1, 2, 3, 5], 77)
    assert np.array_equal(out, want) and np.array_equal(iseed, seed_after)


# ------------------------------------------------------------------------------------------ Cholesky (dpotrf_, flomat.rkt:1794-1826)
def _spd(seed, n):
    m = rnd(seed, n, n)
    return np.asfortranarray(m @ m.T + n * np.eye(n))


def test_oracle_cholesky_matches_numpy():
    for n in (1, 5, 40):
        a = _spd(70 + n, n)
        l, info = Oracle.cholesky(a)
        assert info == 0 and np.abs(l - np.linalg.cholesky(a)).max() <= 1e-13
        u, info = Oracle.cholesky(a, upper=True)
        assert info == 0 and np.abs(u - np.linalg.cholesky(a).T).max() <= 1e-13
    bad = np.asfortranarray(np.array([[4.0, 2.0, 1.0], [2.0, 1.0, 3.0], [1.0, 3.0, -1.0]]))  # second leading minor is singular
    assert Oracle.cholesky(bad)[1] == 2


@gpu
@pytest.mark.parametrize("n", [1, 2, 17, 128, 129, 300, 1000])
@pytest.mark.parametrize("upper", [False, True])
def test_cholesky_parity(fm, ref, n, upper):
    a = _spd(80 + n, n)
    want, _ = ref.cholesky(a, upper=upper)
    got = fm.flomat_cholesky(up(fm, a), upper).to_numpy()
    assert np.abs(got - want).sum(axis=0).max() / np.abs(want).sum(axis=0).max() <= 64 * n * EPS
    rec = got.T @ got if upper else got @ got.T
    assert np.abs(rec - a).sum(axis=0).max() / np.abs(a).sum(axis=0).max() <= 64 * n * EPS
    # in-place form: the other triangle of A is left exactly as it was (LAPACK's contract), info = 0
    A = up(fm, a)
    assert fm.flomat_cholesky_bang(A, upper) == 0
    after = A.to_numpy()
    other = np.tril(after, -1) if upper else np.triu(after, 1)
    assert np.array_equal(other, np.tril(a, -1) if upper else np.triu(a, 1))


@gpu
def test_cholesky_not_positive_definite_and_compat(fm):
    bad = np.asfortranarray(np.array([[4.0, 2.0, 1.0], [2.0, 1.0, 3.0], [1.0, 3.0, -1.0]]))
    assert fm.flomat_cholesky_bang(up(fm, bad)) == 2
    n = 200
    a = _spd(90, n)
    big = a.copy(order="F")
    big[150, 150] = -1e6  # breaks positive definiteness inside the second 128-block
    assert fm.flomat_cholesky_bang(up(fm, big)) == 151
    lib = fm.load()
    h = a.copy(order="F")
    info = ctypes.c_int(-7)
    lib.dpotrf_(b"L", ctypes.byref(ctypes.c_int(n)), h.ctypes.data, ctypes.byref(ctypes.c_int(n)), ctypes.byref(info))  # host buffer
    assert info.value == 0
    assert np.abs(np.tril(h) - np.linalg.cholesky(a)).max() <= 1e-10 and np.array_equal(np.triu(h, 1), np.triu(a, 1))


@gpu
def test_cholesky_large_property(fm):
    n = 4096
    iseed = np.array([5, 6, 7, 9], dtype=np.int32)
    M = fm.randn(n, n, iseed=iseed)
    A = fm.plus(fm.times(M, fm.transpose(M)), fm.dot_times(fm.eye(n), float(n)))
    L = fm.flomat_cholesky(A)
    R = fm.minus(fm.times(L, fm.transpose(L)), A)
    assert fm.norm(R, 1) / fm.norm(A, 1) <= 64 * n * EPS
