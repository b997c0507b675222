"""CPU: pins the oracle (both providers) against every golden vector the reference's own tests hold for the hot path
(tests/golden/reference_tests.json, transcribed from flomat.rkt:3317-3660), against closed-form exponentials, and
against the committed oracle fixtures."""
import math

import numpy as np
import pytest

from oracle.flomat_oracle import PADE8, Flomat, fortran, normwise_rel_err, tol

EPS = 1e-13  # flomat's `epsilon` (flomat.rkt:160): absolute per-entry tolerance of flomat=


def F_(p):
    return Flomat(p)


def close(x, y, tol_=EPS):
    return np.max(np.abs(np.asarray(x) - np.asarray(y))) <= tol_


def test_gemm_golden(golden, providers):
    for p in providers:
        for case in golden["gemm"]:
            got = F_(p).times(fortran(case["A"]), fortran(case["B"]))
            assert np.array_equal(got, np.array(case["AB"], dtype=float)), (p.name, case["cite"])


def test_expt_golden(golden, providers):
    A = fortran(golden["expt"]["A"])
    for p in providers:
        f = F_(p)

        def expt(a, n):  # flomat-expt (flomat.rkt:918-927)
            if n == 0:
                return f.eye(a.shape[0])
            if n == 1:
                return f.copy(a)
            if n == 2:
                return f.gemm(a, a)
            if n % 2 == 0:
                h = expt(a, n // 2)
                return f.gemm(h, h)
            return f.gemm(a, expt(a, n - 1))

        for k, want in golden["expt"]["powers"].items():
            assert np.array_equal(expt(A, int(k)), np.array(want, dtype=float)), (p.name, k)


def test_plus_minus_scale_golden(golden, providers):
    g = golden["plus_minus"]
    for p in providers:
        f = F_(p)
        A, B = fortran(g["A"]), fortran(g["B"])
        assert np.array_equal(f.plus(A, B), np.array(g["A+B"], dtype=float))
        assert np.array_equal(f.minus(A, B), np.array(g["A-B"], dtype=float))
        assert np.array_equal(f.minus(A), np.array(g["-A"], dtype=float))
        s = golden["scale"]
        assert np.array_equal(f.scale(s["s"], fortran(s["A"])), np.array(s["sA"], dtype=float))


def test_solve_inverse_golden(golden, providers):
    for p in providers:
        f = F_(p)
        M, b = fortran(golden["solve"]["M"]), fortran(golden["solve"]["b"])
        assert close(f.gemm(M, f.mldivide(M, b)), b)
        M = fortran(golden["inverse"]["M"])
        assert close(f.gemm(M, f.inv(M)), np.eye(2))
        assert close(f.gemm(f.inv(M), M), np.eye(2))


def test_det_golden(golden, providers):
    for p in providers:
        f = F_(p)
        for case in golden["det"]:
            d = f.det(fortran(case["M"]))
            if case["tol"] == 0.0:
                assert d == case["det"], (p.name, case["cite"], d)
            else:
                assert abs(d - case["det"]) < case["tol"], (p.name, case["cite"], d)


def test_plu_golden(golden, providers):
    M = fortran(golden["plu"]["M"])
    for p in providers:
        lu, ipiv, info = F_(p).lu(M)
        assert info == 0
        n = M.shape[0]
        L = np.tril(lu, -1) + np.eye(n)
        U = np.triu(lu)
        P = np.eye(n)
        for i in range(n - 1, -1, -1):  # pivots->flomat (flomat.rkt:1471-1485)
            ip = ipiv[i] - 1
            if ip != i:
                P[[i, ip], :] = P[[ip, i], :]
        assert close(P @ L @ U, M)


def test_pade_coefficients(golden):
    want = [float.fromhex(h) for h in golden["pade8_coefficients_hex"]["c"]]
    assert PADE8 == want
    # exp-pade (expm.rkt:155-157) agrees with exp on a scalar, as the reference's comment at :159 says
    num = sum(c * 1.0 ** i for i, c in enumerate(PADE8))
    den = sum(c * (-1.0) ** i for i, c in enumerate(PADE8))
    assert abs(num / den - math.e) < 1e-15 * 4


def test_expm_closed_forms(providers):
    for p in providers:
        f = F_(p)
        f.gemm_calls = 0
        E = f.expm(fortran([[1, 0], [0, 2]]))           # manual-flomat.scrbl:823
        assert f.gemm_calls == 11 + 2                     # s = 2  (11 + s dgemms)
        assert normwise_rel_err(E, np.diag([math.e, math.e ** 2])) < tol(2)
        # rotation generator: exp([[0,-t],[t,0]]) = [[cos t, -sin t],[sin t, cos t]]
        t = 0.7
        E = f.expm(fortran([[0, -t], [t, 0]]))
        assert normwise_rel_err(E, [[math.cos(t), -math.sin(t)], [math.sin(t), math.cos(t)]]) < tol(2)
        # nilpotent: exp(N) = I + N + N^2/2
        N = fortran([[0, 1, 2], [0, 0, 3], [0, 0, 0]])
        assert normwise_rel_err(f.expm(N), np.eye(3) + N + N @ N / 2) < tol(3)


def test_expm_reference_quirks(providers):
    """s <= 0 scales A up and never squares; the zero matrix gives NaN."""
    for p in providers:
        f = F_(p)
        A = fortran(np.diag([0.1, 0.05]))
        assert f.scaling_exponent(A) == -2.0
        E = f.expm(A)
        assert normwise_rel_err(E, np.diag([math.exp(0.4), math.exp(0.2)])) < 1e-14
        with np.errstate(all="ignore"):
            Z = f.expm(fortran(np.zeros((2, 2))))
        assert np.isnan(Z).all()


def test_providers_agree_and_match_fixtures(fixtures, providers):
    for n in (8, 33, 64, 130):
        a = np.asfortranarray(fixtures[f"expm_in_{n}"])
        m = np.asfortranarray(fixtures[f"lu_in_{n}"])
        b = np.asfortranarray(fixtures[f"rhs_in_{n}"])
        for p in providers:
            f = F_(p)
            assert f.scaling_exponent(a) == float(fixtures[f"oracle_expm_s_{n}"])
            assert normwise_rel_err(f.expm(a), fixtures[f"oracle_expm_out_{n}"]) < tol(n)
            assert normwise_rel_err(f.mldivide(m, b), fixtures[f"oracle_solve_out_{n}"]) < tol(n) * 10
            assert normwise_rel_err(f.inv(m), fixtures[f"oracle_inv_out_{n}"]) < tol(n) * 10
            d = float(fixtures[f"oracle_det_out_{n}"])
            assert abs(f.det(m) - d) <= 1e-10 * abs(d)
            lu, ipiv, info = f.lu(m)
            assert np.array_equal(ipiv, fixtures[f"oracle_ipiv_{n}"])


def test_expm_matches_scipy_pade13(providers):
    import scipy.linalg

    rng = np.random.default_rng(7)
    for n in (16, 96):
        a = np.asfortranarray(rng.standard_normal((n, n)) / np.sqrt(n))
        want = scipy.linalg.expm(a)
        for p in providers:
            assert normwise_rel_err(F_(p).expm(a), want) < tol(n)


def test_norms(providers):
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.standard_normal((37, 23)))
    for p in providers:
        f = F_(p)
        assert abs(f.norm(a, 1) - np.abs(a).sum(axis=0).max()) < 1e-12
        assert abs(f.norm(a, "inf") - np.abs(a).sum(axis=1).max()) < 1e-12
        assert abs(f.norm(a, "frob") - np.linalg.norm(a)) < 1e-12
        assert f.norm(a, "max") == np.abs(a).max()


def test_dimension_checks(providers):
    f = F_(providers[0])
    with pytest.raises(ValueError):
        f.times(fortran(np.ones((2, 3))), fortran(np.ones((2, 3))))
    with pytest.raises(ValueError):
        f.plus(fortran(np.ones((2, 3))), fortran(np.ones((3, 2))))


def test_pointwise_and_scalar_paths(providers):
    rng = np.random.default_rng(11)
    a = np.asfortranarray(rng.standard_normal((5, 4)))
    b = np.asfortranarray(rng.standard_normal((5, 4)))
    for p in providers:
        f = F_(p)
        assert np.array_equal(f.dot_times(a, b), a * b)
        assert np.array_equal(f.dot_plus(a, 2.5), a + 2.5)
        assert np.array_equal(f.dot_minus(3.0, b), 3.0 - b)
        assert np.array_equal(f.dot_divide(a, b), a / b)
        assert np.array_equal(f.dot_divide(a), 1.0 / a)
        assert np.array_equal(f.times(a, 2.0), a * 2.0)
