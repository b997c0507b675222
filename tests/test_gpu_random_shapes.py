"""GPU: randomized shapes (seeded) through the LU / solve / GEMM paths: non-multiples of the 128-column panels, of the
16-column leaves and of the 8-column DMMA tiles, tall and wide matrices, views with odd offsets, many right-hand sides.
Pivot sequences must equal the oracle's, factors and products stay inside 64 n eps."""
import numpy as np
import pytest

from oracle.flomat_oracle import normwise_rel_err, tol

pytestmark = pytest.mark.gpu


def up(fm, x):
    return fm.matrix(np.asfortranarray(x, dtype=np.float64))


def plu_matrix(lu, ipiv, m, n):
    k = min(m, n)
    l = np.tril(lu[:, :k], -1) + np.eye(m, k)
    u = np.triu(lu[:k, :])
    perm = np.arange(m)
    for i, p in enumerate(ipiv):
        perm[[i, p - 1]] = perm[[p - 1, i]]
    return l @ u, perm


def test_lu_random_shapes(fm, ref):
    rng = np.random.default_rng(2024)
    shapes = [(int(rng.integers(1, 620)), int(rng.integers(1, 620))) for _ in range(36)]
    shapes += [(129, 129), (257, 255), (144, 16), (16, 144), (136, 136), (1, 300), (300, 1), (513, 140), (140, 513)]
    for m, n in shapes:
        a = np.asfortranarray(rng.standard_normal((m, n)))
        lu = up(fm, a)
        piv, info = fm.flomat_lu_bang(lu)
        _, want_piv, _ = ref.lu(a)
        got, ip = lu.to_numpy(), piv.to_numpy()
        assert info == 0, (m, n)
        assert np.array_equal(ip, want_piv[: min(m, n)]), (m, n)
        rec, perm = plu_matrix(got, ip, m, n)
        assert normwise_rel_err(rec, a[perm, :]) <= tol(max(m, n)), (m, n)


def test_lu_first_column_ties_pick_first_row(fm):
    rng = np.random.default_rng(7)
    for n in (5, 130, 300):
        a = np.asfortranarray(rng.standard_normal((n, n)))
        a[:, 0] = rng.choice([-1.0, 1.0], size=n)  # every entry of column 0 has the same magnitude: IDAMAX takes row 0
        lu = up(fm, a)
        piv, _ = fm.flomat_lu_bang(lu)
        assert piv.to_numpy()[0] == 1


def test_solve_random_shapes(fm, ref):
    rng = np.random.default_rng(99)
    for _ in range(14):
        n, nrhs = int(rng.integers(1, 700)), int(rng.integers(1, 420))
        a = np.asfortranarray(rng.standard_normal((n, n)) + np.sqrt(n) * np.eye(n))
        b = np.asfortranarray(rng.standard_normal((n, nrhs)))
        x = fm.mldivide(up(fm, a), up(fm, b)).to_numpy()
        r = np.abs(a @ x - b).sum(axis=0).max() / (np.abs(a).sum(axis=0).max() * np.abs(x).sum(axis=0).max())
        assert r <= tol(n), (n, nrhs)
        assert normwise_rel_err(x, ref.mldivide(a, b)) <= tol(n) * 8, (n, nrhs)


def test_gemm_random_shapes_and_views(fm, ref):
    rng = np.random.default_rng(5)
    for _ in range(24):
        m, n, k = (int(rng.integers(1, 500)) for _ in range(3))
        a, b = np.asfortranarray(rng.standard_normal((m, k))), np.asfortranarray(rng.standard_normal((k, n)))
        assert normwise_rel_err(fm.times(up(fm, a), up(fm, b)).to_numpy(), a @ b) <= tol(max(m, n, k)), (m, n, k)
    big_a, big_b = np.asfortranarray(rng.standard_normal((300, 280))), np.asfortranarray(rng.standard_normal((290, 310)))
    A, B = up(fm, big_a), up(fm, big_b)
    for (i, j, mm, kk, jj, nn) in [(1, 1, 150, 200, 3, 170), (2, 4, 128, 128, 2, 128), (7, 0, 33, 257, 1, 9)]:
        va, vb = A.sub(i, j, mm, kk), B.sub(j, jj, kk, nn)  # odd offsets: 8-byte aligned views, lda > m
        want = big_a[i:i + mm, j:j + kk] @ big_b[j:j + kk, jj:jj + nn]
        assert normwise_rel_err(fm.times(va, vb).to_numpy(), want) <= tol(max(mm, nn, kk))


def test_gemm_large_transposed_and_unaligned_operands(fm, ref):
    """Products big enough to be worth it re-lay transposed / 8-byte-aligned / odd-ld operands into aligned workspaces and run
    the TMA kernel: same results as the generic path and the oracle."""
    rng = np.random.default_rng(11)
    m, n, k = 301, 283, 277
    for ta, tb in [(True, False), (False, True), (True, True)]:
        a = np.asfortranarray(rng.standard_normal((k, m) if ta else (m, k)))
        b = np.asfortranarray(rng.standard_normal((n, k) if tb else (k, n)))
        c = np.asfortranarray(rng.standard_normal((m, n)))
        C = up(fm, c)
        fm.flomat_times(up(fm, a), up(fm, b), C, alpha=1.5, beta=-0.25, transpose_a=ta, transpose_b=tb)
        want = ref.gemm(a, b, c.copy(order="F"), alpha=1.5, beta=-0.25, ta=ta, tb=tb)
        assert normwise_rel_err(C.to_numpy(), want) <= tol(max(m, n, k)), (ta, tb)
    big_a, big_b = np.asfortranarray(rng.standard_normal((601, 580))), np.asfortranarray(rng.standard_normal((590, 611)))
    A, B = up(fm, big_a), up(fm, big_b)  # odd leading dimensions and odd offsets
    va, vb = A.sub(1, 3, 500, 400), B.sub(3, 1, 400, 450)
    want = big_a[1:501, 3:403] @ big_b[3:403, 1:451]
    assert normwise_rel_err(fm.times(va, vb).to_numpy(), want) <= tol(500)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nrhs", [(129, 1), (257, 65), (640, 8), (1000, 200), (2050, 64), (3001, 130), (4100, 17)])
def test_solve_few_right_hand_sides_wavefront(fm, ref, n, nrhs):
    """dgetrs with nrhs <= 256 and n > 128 runs the single-launch wavefront triangular solves (one CTA per 128-row block,
    x_k handed from CTA to CTA through release/acquire flags): partial last blocks, partial column chunks, idle warps."""
    rng = np.random.default_rng(n + nrhs)
    a = np.asfortranarray(rng.standard_normal((n, n)))
    b = np.asfortranarray(rng.standard_normal((n, nrhs)))
    x = fm.mldivide(up(fm, a), up(fm, b)).to_numpy()
    r = np.abs(a @ x - b).sum(axis=0).max() / (np.abs(a).sum(axis=0).max() * np.abs(x).sum(axis=0).max())
    assert r <= tol(n), (n, nrhs)
    assert normwise_rel_err(x, ref.mldivide(a, b)) <= tol(n) * 50, (n, nrhs)


@pytest.mark.gpu
def test_getrs_wavefront_unaligned_view_and_repeat(fm, ref):
    """B as an 8-byte-aligned view with an odd leading dimension (scalar load/store path), solved twice with the same factors
    (the flag epoch must separate the launches)."""
    lib = fm.load()
    n, nrhs = 700, 50
    rng = np.random.default_rng(4)
    a = np.asfortranarray(rng.standard_normal((n, n)))
    lu = up(fm, a)
    piv, info = fm.flomat_lu_bang(lu)
    assert info == 0
    for rep in range(3):
        b = np.asfortranarray(rng.standard_normal((n + 3, nrhs + 2)))
        big = up(fm, b)
        view = big.sub(1, 1, n, nrhs)  # odd row offset: 8-byte aligned, ld = n + 3 (odd)
        assert lib.fm_getrs(n, nrhs, lu.a, lu.lda, piv.ptr, view.a, view.lda) == 0
        got = big.to_numpy()
        want = np.linalg.solve(a, b[1:n + 1, 1:nrhs + 1])
        assert normwise_rel_err(got[1:n + 1, 1:nrhs + 1], want) <= tol(n) * 50
        untouched = np.ones_like(b, dtype=bool)
        untouched[1:n + 1, 1:nrhs + 1] = False
        assert np.array_equal(got[untouched], b[untouched])  # nothing outside the view was written


@pytest.mark.gpu
def test_wavefront_solve_is_repeatable(fm):
    """The wavefront hands x_k from CTA to CTA through flags in global memory: 40 solves with the same factors must give the
    same bits every time and a backward error inside the bound (a rare ordering bug would show up as a sporadic difference)."""
    import torch

    lib = fm.load()
    n, nrhs = 1500, 130
    g = torch.Generator(device="cuda"); g.manual_seed(77)
    a = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(n * nrhs, dtype=torch.float64, device="cuda", generator=g)
    A, LU = fm.Flomat.alloc(n, n), fm.Flomat.alloc(n, n)
    lib.fm_copy2d(A.a, n, a.data_ptr(), n, n, n)
    lib.fm_copy2d(LU.a, n, a.data_ptr(), n, n, n)
    piv = fm.Pivots(n)
    assert lib.fm_getrf(n, n, LU.a, n, piv.ptr, piv.info_ptr) == 0
    B, X = fm.Flomat.alloc(n, nrhs), fm.Flomat.alloc(n, nrhs)
    lib.fm_copy2d(B.a, n, b.data_ptr(), n, n, nrhs)
    first = None
    for _ in range(40):
        lib.fm_copy2d(X.a, n, b.data_ptr(), n, n, nrhs)
        assert lib.fm_getrs(n, nrhs, LU.a, n, piv.ptr, X.a, n) == 0
        x = X.to_numpy()
        if first is None:
            first = x
            res = fm.norm(fm.minus(fm.times(A, X), B), 1) / (fm.norm(A, 1) * fm.norm(X, 1))
            assert res <= tol(n)
        else:
            assert np.array_equal(first, x)
