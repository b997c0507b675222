"""The UNMODIFIED reference front end, restated call for call over the compat tier of libflomat_cuda.

`racket` is not installed in this image, so `raco test flomat/flomat.rkt` cannot run.  This module is the closest executable form of
it: it performs exactly the foreign calls that flomat.rkt / expm.rkt make when `cblas-lib` / `lapack-lib` (flomat.rkt:92-135) resolve
to libflomat_cuda.so and NOTHING else in the Racket sources changes --

  * matrices are GC'd `malloc` buffers, i.e. pageable HOST memory (alloc-flomat, flomat.rkt:437-441): numpy Fortran-order arrays here;
  * every arithmetic step that the reference sends through its FFI goes through the same symbol with the same arguments:
    cblas_dgemm (:888), cblas_daxpy one column at a time (:805-813), cblas_dcopy (:497-500, and :544 with incX = 0), cblas_dscal
    (:1020-1030), dlange_ (:1340), dgetrf_ (:1527), dgetri_ (:1778), dgesv_ (:2131, :2140);
  * what the reference does in Racket itself stays on the host, as it would: `memset` for zero matrices (:543), the `ptr-set!` loop
    of flomat-eye (:2642), and the per-element loops of the pointwise macros `.* ./ .+ .-` (:3021-3146).  numpy stands in for the
    Racket interpreter there (one IEEE operation per element, same as the loops).

It exists to TEST and TIME the zero-change drop-in (tests/test_gpu_compat_front.py, bench.py `other_configs.expm_compat`); the fast
path for a maintained binding layer is sci_b200.flomat (device-resident tier).  `calls` counts the foreign calls by symbol.
"""
from __future__ import annotations

import math
from collections import Counter
from ctypes import byref, c_int

import numpy as np

from . import _lib

COL_MAJOR, NO_TRANS, TRANS = 102, 111, 112  # flomat.rkt:846-853


def _pade8():
    """ncoef (expm.rkt:136-147): c_i = (16-i)! 8! / (16! i! (8-i)!), exact rational -> flonum."""
    from fractions import Fraction
    from math import factorial as f

    return [float(Fraction(f(16 - i) * f(8), f(16) * f(i) * f(8 - i))) for i in range(9)]


N8 = _pade8()


class RacketFrontEnd:
    def __init__(self, lib=None):
        """lib: the object whose attributes are the eight foreign functions (default: libflomat_cuda.so through ctypes -- the product).
        The CPU tests pass a host BLAS/LAPACK with the same symbols, which turns this class into the reference itself."""
        self.L = lib if lib is not None else _lib.load()
        self.calls: Counter = Counter()

    # ---- allocation and copies (flomat.rkt:437-545) ------------------------------------------------------------------
    def make_flomat(self, m: int, n: int, x: float = 0.0) -> np.ndarray:
        """make-flomat (:538-545): memset for 0.0, otherwise ONE cblas_dcopy with incX = 0 from a one-element buffer."""
        a = np.empty((m, n), dtype=np.float64, order="F")
        if x == 0.0:
            a.fill(0.0)  # (memset a 0 ...) in Racket
        elif a.size:
            one = np.array([float(x)])
            self.calls["cblas_dcopy"] += 1
            self.L.cblas_dcopy(m * n, one.ctypes.data, 0, a.ctypes.data, 1)
        return a

    def zeros(self, n: int) -> np.ndarray:
        return self.make_flomat(n, n, 0.0)

    def eye(self, n: int) -> np.ndarray:
        """flomat-eye (:2642): zeros, then a Racket loop of ptr-set! on the diagonal."""
        a = self.make_flomat(n, n, 0.0)
        a[np.arange(n), np.arange(n)] = 1.0
        return a

    def copy_flomat(self, a: np.ndarray) -> np.ndarray:
        """copy-flomat (:516-527): lda = m -> one cblas_dcopy of m*n elements (unsafe-vector-copy! :497-500)."""
        b = np.empty(a.shape, dtype=np.float64, order="F")
        if a.size:
            self.calls["cblas_dcopy"] += 1
            self.L.cblas_dcopy(a.size, a.ctypes.data, 1, b.ctypes.data, 1)
        return b

    # ---- times (:3202 -> flomat* :905 -> :877 -> cblas_dgemm :888) ----------------------------------------------------------
    def flomat_times(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        m, k = a.shape
        k2, n = b.shape
        if k != k2:
            raise ValueError("check-product-dimensions (flomat.rkt:340)")
        c = self.make_flomat(m, n, 0.0)  # flomat* allocates a ZERO matrix and multiplies with alpha = beta = 1 (:909-912)
        self.calls["cblas_dgemm"] += 1
        self.L.cblas_dgemm(COL_MAJOR, NO_TRANS, NO_TRANS, m, n, k, 1.0, a.ctypes.data, m, b.ctypes.data, k, 1.0, c.ctypes.data, m)
        return c

    def times(self, a, *bs):
        for b in bs:  # left fold (:3202-3212)
            if np.isscalar(a) and np.isscalar(b):
                a = a * b
            elif np.isscalar(a) or np.isscalar(b):
                a = self.dot_times(a, b)
            else:
                a = self.flomat_times(a, b)
        return a

    # ---- plus / minus (:3158, :3180 -> flomat+ :826, flomat- :836 -> :815 -> :805: n daxpy calls) ---------------------------
    def _axpy_columns(self, alpha: float, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """constant*flomat+flomat! (:805-813): B := alpha A + B, one cblas_daxpy per column."""
        m, n = a.shape
        pa, pb = a.ctypes.data, b.ctypes.data
        daxpy = self.L.cblas_daxpy
        for j in range(n):
            daxpy(m, alpha, pa + 8 * j * m, 1, pb + 8 * j * m, 1)
        self.calls["cblas_daxpy"] += n
        return b

    def plus(self, a, *bs):
        for b in bs:
            if np.isscalar(a) or np.isscalar(b):
                a = self.dot_plus(a, b)
            else:
                if a.shape != b.shape:
                    raise ValueError("check-same-dimensions (flomat.rkt:326)")
                a = self._axpy_columns(1.0, a, self.copy_flomat(b))
        return a

    def minus(self, a, *bs):
        if not bs:
            return self.dot_minus(a)
        for b in bs:
            if np.isscalar(a) or np.isscalar(b):
                a = self.dot_minus(a, b)
            else:
                if a.shape != b.shape:
                    raise ValueError("check-same-dimensions (flomat.rkt:326)")
                a = self._axpy_columns(-1.0, b, self.copy_flomat(a))
        return a

    def flomat_scale(self, s: float, a: np.ndarray) -> np.ndarray:
        """flomat-scale (:1036) -> constant*matrix! (:1020-1030): copy, then one cblas_dscal over the dense buffer."""
        c = self.copy_flomat(a)
        if c.size:
            self.calls["cblas_dscal"] += 1
            self.L.cblas_dscal(c.size, float(s), c.ctypes.data, 1)
        return c

    # ---- pointwise macros (:3021-3146): Racket loops in the reference, host numpy here (they never cross the FFI) ----------
    @staticmethod
    def _pointwise(op, a, b):
        with np.errstate(all="ignore"):
            return np.asfortranarray(op(a, b))

    def dot_times(self, a, b):
        return self._pointwise(np.multiply, a, b)

    def dot_plus(self, a, b):
        return self._pointwise(np.add, a, b)

    def dot_minus(self, a, b=None):
        return np.asfortranarray(-a) if b is None else self._pointwise(np.subtract, a, b)

    def dot_divide(self, a, b=None):
        return self._pointwise(np.divide, 1.0, a) if b is None else self._pointwise(np.divide, a, b)

    # ---- norm (:3273 -> :1328 -> dlange_ :1340) ------------------------------------------------------------------------------
    def norm(self, a: np.ndarray, kind=1) -> float:
        code = {1: b"1", "inf": b"I", "frob": b"F", "max": b"M"}[kind]
        m, n = a.shape
        work = self.make_flomat(max(1, m) if kind == "inf" else 1, 1)  # (make-flomat lwork 1), :1338
        self.calls["dlange_"] += 1
        return float(self.L.dlange_(code, byref(c_int(m)), byref(c_int(n)), a.ctypes.data, byref(c_int(m)), work.ctypes.data))

    # ---- LU family ---------------------------------------------------------------------------------------------------------------
    def flomat_lu_bang(self, a: np.ndarray):
        """flomat-lu! (:1525-1527): dgetrf_ in place; ipiv is a fresh u32vector of m entries."""
        m, n = a.shape
        ipiv, info = np.zeros(max(m, 1), dtype=np.int32), c_int(0)
        self.calls["dgetrf_"] += 1
        self.L.dgetrf_(byref(c_int(m)), byref(c_int(n)), a.ctypes.data, byref(c_int(m)), ipiv.ctypes.data, byref(info))
        return ipiv, info.value

    def mldivide(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """mldivide (:3289) -> flomat-solve-many (:2145-2152): copy A, copy B, dgesv_, info ignored."""
        lu, x = self.copy_flomat(a), self.copy_flomat(b)
        n, nrhs = lu.shape[0], x.shape[1]
        ipiv, info = np.zeros(max(n, 1), dtype=np.int32), c_int(0)
        self.calls["dgesv_"] += 1
        self.L.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), lu.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, x.ctypes.data,
                      byref(c_int(n)), byref(info))
        self.last_info = info.value
        return x

    def inv(self, a: np.ndarray) -> np.ndarray:
        """inv (:3297) -> flomat-inverse! (:1770-1779) on a copy: work = another copy, dgetrf_, dgetri_ with lwork = m*m."""
        lu = self.copy_flomat(a)
        work = self.copy_flomat(lu)
        n = lu.shape[0]
        ipiv, _ = self.flomat_lu_bang(lu)
        info = c_int(0)
        self.calls["dgetri_"] += 1
        self.L.dgetri_(byref(c_int(n)), lu.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, work.ctypes.data, byref(c_int(n * n)), byref(info))
        self.last_info = info.value
        return lu

    def det(self, a: np.ndarray) -> float:
        """det (:3276) -> :1844 -> :1837: dgetrf_ on a copy, then pivots-sign (:1487) times the left-to-right diagonal product in Racket."""
        lu = self.copy_flomat(a)
        ipiv, _ = self.flomat_lu_bang(lu)
        sign = 1
        for i in range(lu.shape[0]):
            if ipiv[i] - 1 != i:
                sign = -sign
        prod = 1.0
        for i in range(lu.shape[0]):
            prod = prod * float(lu[i, i])
        return sign * prod

    # ---- expm (expm.rkt:167-215) ---------------------------------------------------------------------------------------------
    def mpoly(self, cs, a: np.ndarray) -> np.ndarray:
        n = a.shape[0]
        ident, zero = self.eye(n), self.zeros(n)

        def loop(cs):
            if not cs:
                return zero
            return self.plus(self.dot_times(cs[0], ident), self.times(a, loop(cs[1:])))

        return loop(list(cs))

    def find_scaling_exponent(self, a: np.ndarray) -> float:
        nrm = self.norm(a, 1)
        if nrm == 0.0:
            return -math.inf
        if math.isnan(nrm) or math.isinf(nrm):
            return nrm
        return 1.0 + math.ceil(math.log(nrm) / math.log(2.0))  # (+ 1 (ceiling (log N 2))), expm.rkt:214

    def expm_pade_no_scaling(self, x: np.ndarray) -> np.ndarray:
        aa = self.times(x, x)
        even = self.mpoly(N8[0::2], aa)
        odd = self.times(x, self.mpoly(N8[1::2], aa))
        num = self.plus(even, odd)
        den = self.minus(even, odd)
        return self.mldivide(den, num)

    def expm(self, a: np.ndarray) -> np.ndarray:
        s = self.find_scaling_exponent(a)
        if math.isnan(s) or s == math.inf:
            raise FloatingPointError("the reference does not terminate on an inf/NaN norm (expm.rkt:209)")
        two_s = 0.0 if s == -math.inf else 2.0 ** s
        r = self.expm_pade_no_scaling(self.dot_divide(a, two_s))
        while s > 0:  # repeated-squaring (expm.rkt:209-210)
            r = self.times(r, r)
            s -= 1
        return r
