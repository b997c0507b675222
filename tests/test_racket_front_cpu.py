"""CPU: sci_b200/racket_front.py (the unmodified flomat.rkt / expm.rkt call sequences, foreign call by foreign call) driven over a HOST
BLAS/LAPACK that exports the same eight symbols.  With OpenBLAS behind it the class IS the reference (same library class, same calls,
same arguments), so it must agree with the oracle's independent restatement of those sequences and make exactly the foreign calls the
reference makes (11 + s cblas_dgemm, 11 n cblas_daxpy, one dgesv_, one dlange_ per expm).  The same class over
libflomat_cuda's compat symbols is what tests/test_gpu_compat_front.py and bench.py's `expm_compat` run on the GPU."""
import ctypes
import math

import numpy as np
import pytest

from oracle.flomat_oracle import Flomat, OpenBLASProvider, normwise_rel_err, tol
from sci_b200.racket_front import RacketFrontEnd


class HostLibrary:
    """The eight symbols of flomat.rkt's FFI (:483, :751, :855, :1014, :1317, :1505, :1754, :2112) from the oracle's OpenBLAS handle,
    callable the way the front end calls them: raw addresses for the arrays, byref(c_int) for LAPACK's scalars."""

    def __init__(self, provider):
        def wrap(fn):
            types = fn.argtypes

            def call(*args):
                conv = []
                for a, t in zip(args, types):
                    if isinstance(a, int) and isinstance(t, type) and issubclass(t, ctypes._Pointer):
                        a = ctypes.cast(a, t)
                    elif isinstance(a, bytes):
                        a = ctypes.cast(ctypes.c_char_p(a), t)
                    conv.append(a)
                return fn(*conv)

            return call

        for name, attr in (("cblas_dgemm", "_dgemm"), ("cblas_daxpy", "_daxpy"), ("cblas_dcopy", "_dcopy"), ("cblas_dscal", "_dscal"),
                           ("dlange_", "_dlange"), ("dgetrf_", "_dgetrf"), ("dgesv_", "_dgesv"), ("dgetri_", "_dgetri")):
            setattr(self, name, wrap(getattr(provider, attr)))


@pytest.fixture(scope="module")
def both():
    try:
        prov = OpenBLASProvider()
    except (FileNotFoundError, OSError, AttributeError) as e:  # pragma: no cover
        pytest.skip(f"no LP64 OpenBLAS in this image: {e}")
    return RacketFrontEnd(HostLibrary(prov)), Flomat(prov)


def rnd(seed, m, n, scale=1.0):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)) * scale)


@pytest.mark.parametrize("n,scale", [(3, 1.0), (16, 0.25), (33, 1.0), (64, 1.0 / 8.0), (50, 3.0)])
def test_expm_call_sequence_and_result(both, n, scale):
    front, ref = both
    a = rnd(100 + n, n, n, scale)
    front.calls.clear()
    got = front.expm(a)
    want = ref.expm(a)
    assert normwise_rel_err(got, want) <= tol(n)
    s = ref.scaling_exponent(a)
    squarings = int(s) if s > 0 else 0
    assert front.calls["cblas_dgemm"] == 11 + squarings   # expm.rkt:194-199, 209-210: 1 + 5 + 4 + 1, then s squarings
    assert front.calls["cblas_daxpy"] == 11 * n             # 9 `plus` inside the two mpoly + num + den, one daxpy per column (:805-813)
    assert front.calls["dgesv_"] == 1 and front.calls["dlange_"] == 1
    assert front.calls["dgetrf_"] == 0 and front.calls["dgetri_"] == 0
    assert front.find_scaling_exponent(a) == s


def test_the_other_hot_path_functions(both):
    front, ref = both
    n = 40
    a, b, c = rnd(1, n, n), rnd(2, n, n), rnd(3, n, 5)
    assert np.array_equal(front.plus(a, b), ref.plus(a, b)) and np.array_equal(front.minus(a, b), ref.minus(a, b))
    assert np.array_equal(front.plus(a, b, a), a + b + a)
    assert normwise_rel_err(front.times(a, b), ref.times(a, b)) <= tol(n)
    assert normwise_rel_err(front.mldivide(a, c), ref.mldivide(a, c)) <= 100 * tol(n)
    assert normwise_rel_err(front.inv(a), ref.inv(a)) <= 100 * tol(n)
    assert math.isclose(front.det(a), ref.det(a), rel_tol=1e-10)
    for kind in (1, "inf", "frob", "max"):
        assert math.isclose(front.norm(a, kind), ref.norm(a, kind), rel_tol=1e-13)
    assert np.array_equal(front.flomat_scale(-2.5, a), -2.5 * a)
    assert np.array_equal(front.make_flomat(4, 3, 7.5), np.full((4, 3), 7.5)) and np.array_equal(front.eye(5), np.eye(5))
    x = front.dot_divide(a, 4.0)
    assert np.array_equal(x, a / 4.0) and x.flags.f_contiguous


def test_quirks_of_the_reference_survive(both):
    front, _ = both
    assert np.isnan(front.expm(np.zeros((3, 3), order="F"))).all()       # A / 2^-inf (expm.rkt:205) = 0/0
    q = np.asfortranarray(np.diag([0.1, 0.05]))
    assert front.find_scaling_exponent(q) == -2.0                         # s <= 0: scaled UP, never squared
    assert normwise_rel_err(front.expm(q), np.diag([math.exp(0.4), math.exp(0.2)])) <= 1e-14
    with pytest.raises(FloatingPointError):
        front.expm(np.asfortranarray([[np.inf, 0.0], [0.0, 1.0]]))
