"""GPU: the unmodified reference front end, call for call, over the COMPAT symbols of libflomat_cuda with pageable host buffers
(sci_b200/racket_front.py; SURVEY 8f row N1 -- `raco test` itself cannot run, there is no racket in the image).

The same call sequences run in the oracle over OpenBLAS; results must agree within 64 n eps (bit-exact where the reference's own
operations are single IEEE roundings), and the NUMBER of foreign calls per symbol must be the reference's (11 + s dgemms for expm,
one daxpy per column for plus/minus, ...)."""
import math

import numpy as np
import pytest

from oracle.flomat_oracle import tol
from parity import gate, gate_inverse, gate_solve

pytestmark = pytest.mark.gpu


def rnd(seed, m, n, scale=1.0):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)) * scale)


@pytest.fixture(scope="module")
def front(fm):
    from sci_b200.racket_front import RacketFrontEnd

    return RacketFrontEnd()


def test_reference_goldens_through_compat_front(front, golden):
    """The reference's own rackunit vectors for the path (flomat.rkt:3335-3617) through the zero-change front end."""
    from oracle.flomat_oracle import fortran

    for case in golden["gemm"]:
        assert np.array_equal(front.times(fortran(case["A"]), fortran(case["B"])), np.array(case["AB"], dtype=float)), case["cite"]
    g = golden["plus_minus"]
    A, B = fortran(g["A"]), fortran(g["B"])
    assert np.array_equal(front.plus(A, B), np.array(g["A+B"], dtype=float))
    assert np.array_equal(front.minus(A, B), np.array(g["A-B"], dtype=float))
    s = golden["scale"]
    assert np.array_equal(front.flomat_scale(s["s"], fortran(s["A"])), np.array(s["sA"], dtype=float))
    M, b = fortran(golden["solve"]["M"]), fortran(golden["solve"]["b"])
    assert np.max(np.abs(M @ front.mldivide(M, b) - b)) <= 1e-13
    M = fortran(golden["inverse"]["M"])
    assert np.max(np.abs(M @ front.inv(M) - np.eye(2))) <= 1e-13
    for case in golden["det"]:
        d = front.det(fortran(case["M"]))
        assert (d == case["det"]) if case["tol"] == 0.0 else (abs(d - case["det"]) < case["tol"]), case["cite"]


@pytest.mark.parametrize("n", [64, 512, 2048])
def test_front_end_sequences_match_oracle(front, ref, n):
    a, b = rnd(100 + n, n, n), rnd(200 + n, n, n)
    rhs = rnd(300 + n, n, 5)
    front.calls.clear()
    # elementwise through daxpy / dscal / dcopy: one rounding per element, identical bytes
    assert np.array_equal(front.plus(a, b), ref.plus(a, b))
    assert np.array_equal(front.minus(a, b), ref.minus(a, b))
    assert np.array_equal(front.flomat_scale(-0.75, a), ref.scale(-0.75, a))
    assert np.array_equal(front.make_flomat(n, 3, 2.5), np.full((n, 3), 2.5))
    assert front.calls["cblas_daxpy"] == 2 * n and front.calls["cblas_dscal"] == 1 and front.calls["cblas_dcopy"] == 4
    gate(f"front times n={n}", front.times(a, b), ref.times(a, b), n)
    for kind in (1, "inf", "frob", "max"):
        got, want = front.norm(a, kind), ref.norm(a, kind)
        assert abs(got - want) <= 4 * n * 2.0 ** -52 * abs(want), kind
    gate_solve(f"front mldivide n={n}", a, rhs, front.mldivide(a, rhs), ref.mldivide(a, rhs), n)
    gate_inverse(f"front inv n={n}", a, front.inv(a), ref.inv(a), n)
    if n <= 512:
        d, want = front.det(a), ref.det(a)
        assert abs(d - want) <= 64 * n * 2.0 ** -52 * abs(want) or (math.isinf(want) and d == want)


@pytest.mark.parametrize("n", [64, 512, 2048])
def test_front_end_expm_matches_oracle_and_call_counts(front, ref, n):
    x = rnd(2001, n, n, 1.0 / math.sqrt(n))
    s = ref.scaling_exponent(x)
    front.calls.clear()
    got = front.expm(x)
    gate(f"front expm n={n}", got, ref.expm(x), n)
    assert front.find_scaling_exponent(x) == s
    # what expm.rkt executes: 11 + s products (two of them trivial), 11 plus/minus of n daxpys each, one dgesv with n right-hand sides
    assert front.calls["cblas_dgemm"] == 11 + int(s)
    assert front.calls["cblas_daxpy"] == 11 * n
    assert front.calls["dgesv_"] == 1 and front.calls["dlange_"] == 2


def test_front_end_quirks(front):
    """SURVEY F5 through the zero-change path: s <= 0 scales A up and never squares; the zero matrix gives NaN."""
    q = np.asfortranarray(np.diag([0.1, 0.05]))
    e = front.expm(q)
    assert abs(e[0, 0] - math.exp(0.4)) <= 1e-14 and abs(e[1, 1] - math.exp(0.2)) <= 1e-14
    assert np.isnan(front.expm(np.zeros((3, 3), order="F"))).all()
