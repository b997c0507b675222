"""GPU: the paths with asynchronous hand-overs -- split-K products chained as programmatic dependents (dgemm.cu), the Neumann-product
and LU solves inside expm, the look-ahead LU (panel cluster beside the trailing update), the host-operand LU pipeline -- are
deterministic by construction (fixed tile and summation order).  Running each several times on the same input and comparing BIT FOR
BIT with the first result turns an ordering bug (a kernel that starts before its input is complete) into a failure.  The long form
of this check is tools/repeat_check.py (profiles/r02_repeat_check.txt)."""
import math
from ctypes import byref, c_int

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rnd(seed, m, n, scale=1.0):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)) * scale)


def same_every_time(fn, reps):
    first = fn()
    return sum(0 if np.array_equal(fn(), first) else 1 for _ in range(reps - 1))


@pytest.mark.parametrize("n", [256, 512, 768])
def test_split_k_chain_is_repeatable(fm, n):
    A, B = fm.matrix(rnd(1, n, n)), fm.matrix(rnd(2, n, n))

    def chain():  # products feeding products, an elementwise kernel in between
        C = fm.times(A, B)
        D = fm.times(C, B)
        return fm.times(fm.plus(D, C), A).to_numpy()

    assert same_every_time(chain, 60) == 0


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n", [200, 512, 1024])
def test_expm_is_repeatable(fm, n, mode):
    X = fm.matrix(rnd(3, n, n, 1.0 / math.sqrt(n)))
    assert same_every_time(lambda: fm.expm(X, mode).to_numpy(), 30) == 0


@pytest.mark.parametrize("n", [1000, 2304])
def test_look_ahead_lu_is_repeatable(fm, n):
    A = fm.matrix(rnd(4, n, n))

    def plu():
        P, L, U = fm.flomat_plu(A)
        return np.concatenate([L.to_numpy().ravel(), U.to_numpy().ravel(), P.to_numpy().ravel()])

    assert same_every_time(plu, 10) == 0


def test_host_lu_pipeline_is_repeatable(fm, monkeypatch):
    monkeypatch.setenv("FLOMAT_CUDA_LU_PIPELINE_MIN_N", "2048")
    lib = fm.load()
    n, nrhs = 2560, 3
    a, b = rnd(5, n, n), rnd(6, n, nrhs)

    def solve():
        A, B = a.copy(order="F"), b.copy(order="F")
        ipiv = np.zeros(n, dtype=np.int32)
        info = c_int(-1)
        lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
        assert info.value == 0
        return np.concatenate([A.ravel(), B.ravel(), ipiv.astype(np.float64)])

    assert same_every_time(solve, 8) == 0
