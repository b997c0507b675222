"""CPU oracle for flomat's dense FP64 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under sci_b200/ imports it; the product path fails loudly without its CUDA library.

Two layers:

1. *Providers* of the eight native symbols the reference binds on this path (flomat/flomat.rkt:139-140
   `define-cblas` / `define-lapack`), called with the same argument shapes as the Racket `_fun` types:
     - ``CProvider``        -> oracle/flomat_oracle.c (netlib reference algorithms restated in plain C)
     - ``OpenBLASProvider`` -> the LP64 OpenBLAS shipped inside SciPy's wheel (`scipy_`-prefixed symbols).
       The reference's arithmetic is "whatever system CBLAS/LAPACK ffi-lib finds" (flomat.rkt:92-135,
       un-pinned), and OpenBLAS is what Debian/Ubuntu/Arch/Fedora resolve `libblas.so.3` / `libcblas.so.3`
       to in practice, so this provider *is* the reference's CPU path and is what bench.py times.

2. *Call sequences*: the Racket-level functions of the hot path restated operation by operation
   (``Flomat`` below), each citing the reference lines it follows.  Matrices are numpy float64 arrays in
   Fortran (column-major) order, exactly the reference's `(m n a lda)` layout with lda = m.

Parity pinning: tests/test_oracle_golden.py checks both providers against every golden vector the
reference's rackunit suite holds for this path (tests/golden/reference_tests.json, transcribed from
flomat.rkt:3335-3660 with line citations).  expm.rkt has no tests at all: expm parity is UNPINNED by the
reference beyond the manual's doc examples (manual-flomat.scrbl:823-826) and the closed-form cases in
tests/golden (diagonal, nilpotent and rotation generators whose exponentials are known exactly).
"""
from __future__ import annotations

import ctypes
import glob
import math
import os
from ctypes import POINTER, byref, c_char, c_double, c_int, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DP = POINTER(c_double)
_IP = POINTER(c_int)

COL_MAJOR, NO_TRANS, TRANS = 102, 111, 112  # flomat.rkt:846-853


def _dp(x: np.ndarray):
    return x.ctypes.data_as(_DP)


def fortran(x) -> np.ndarray:
    """float64, column-major, owned copy -- what `copy-flomat` (flomat.rkt:516) returns."""
    return np.array(x, dtype=np.float64, order="F", copy=True, ndmin=2)


# --------------------------------------------------------------------------------------------------
# Providers
# --------------------------------------------------------------------------------------------------
class _Provider:
    name = "?"

    # every provider exposes: dgemm daxpy dcopy dscal dlange dgetrf dgesv dgetri (raw, in place)
    def threads(self) -> int:
        return 1


class CProvider(_Provider):
    """oracle/flomat_oracle.c, built by `make -C oracle` (or __graft_entry__.build())."""

    name = "c-restatement"

    def __init__(self, threads: int | None = None):
        path = os.path.join(_HERE, "_build", "libflomat_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` or __graft_entry__.build()")
        L = ctypes.CDLL(path)
        L.ora_dgemm.argtypes = [c_int, c_int, c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]
        L.ora_daxpy.argtypes = [c_int, c_double, _DP, c_int, _DP, c_int]
        L.ora_dcopy.argtypes = [c_int, _DP, c_int, _DP, c_int]
        L.ora_dscal.argtypes = [c_int, c_double, _DP, c_int]
        L.ora_dlange.argtypes = [POINTER(c_char), _IP, _IP, _DP, _IP, _DP]
        L.ora_dlange.restype = c_double
        L.ora_dgetrf.argtypes = [_IP, _IP, _DP, _IP, _IP, _IP]
        L.ora_dgesv.argtypes = [_IP, _IP, _DP, _IP, _IP, _DP, _IP, _IP]
        L.ora_dgetri.argtypes = [_IP, _DP, _IP, _IP, _DP, _IP, _IP]
        L.ora_ewise.argtypes = [c_int, c_int, c_int, c_void_p, c_int, c_double, c_void_p, c_int, c_double, _DP, c_int]
        L.ora_set_threads.argtypes = [c_int]
        self._t = threads or min(os.cpu_count() or 1, 16)
        L.ora_set_threads(self._t)
        self.L = L

    def threads(self):
        return self._t

    def dgemm(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        self.L.ora_dgemm(ta, tb, m, n, k, alpha, _dp(a), lda, _dp(b), ldb, beta, _dp(c), ldc)

    def daxpy(self, n, alpha, x, incx, y, incy):
        self.L.ora_daxpy(n, alpha, x, incx, y, incy)

    def dcopy(self, n, x, incx, y, incy):
        self.L.ora_dcopy(n, x, incx, y, incy)

    def dscal(self, n, alpha, x, incx):
        self.L.ora_dscal(n, alpha, x, incx)

    def dlange(self, norm: bytes, m, n, a, lda, work):
        return self.L.ora_dlange(ctypes.c_char_p(norm), byref(c_int(m)), byref(c_int(n)), _dp(a), byref(c_int(lda)),
                                 _dp(work) if work is not None else None)

    def dgetrf(self, m, n, a, lda, ipiv):
        info = c_int(0)
        self.L.ora_dgetrf(byref(c_int(m)), byref(c_int(n)), _dp(a), byref(c_int(lda)), ipiv.ctypes.data_as(_IP), byref(info))
        return info.value

    def dgesv(self, n, nrhs, a, lda, ipiv, b, ldb):
        info = c_int(0)
        self.L.ora_dgesv(byref(c_int(n)), byref(c_int(nrhs)), _dp(a), byref(c_int(lda)), ipiv.ctypes.data_as(_IP),
                         _dp(b), byref(c_int(ldb)), byref(info))
        return info.value

    def dgetri(self, n, a, lda, ipiv, work, lwork):
        info = c_int(0)
        self.L.ora_dgetri(byref(c_int(n)), _dp(a), byref(c_int(lda)), ipiv.ctypes.data_as(_IP),
                          _dp(work) if work is not None else None, byref(c_int(lwork)), byref(info))
        return info.value

    def ewise(self, op, m, n, a, lda, sa, b, ldb, sb, c, ldc):
        pa = a.ctypes.data if a is not None else None
        pb = b.ctypes.data if b is not None else None
        self.L.ora_ewise(op, m, n, pa, lda, sa, pb, ldb, sb, _dp(c), ldc)


def find_openblas() -> str | None:
    import scipy  # noqa: F401  (only to locate the wheel's bundled library)

    root = os.path.join(os.path.dirname(scipy.__file__), os.pardir, "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas-*.so")))
    return hits[0] if hits else None


class OpenBLASProvider(_Provider):
    """LP64 OpenBLAS from SciPy's wheel: same symbols the reference binds, with a `scipy_` prefix."""

    name = "openblas"

    def __init__(self, threads: int | None = None):
        path = find_openblas()
        if path is None:
            raise FileNotFoundError("scipy.libs/libscipy_openblas-*.so not found")
        L = ctypes.CDLL(path)
        self.path = path
        g = lambda s: getattr(L, "scipy_" + s)  # noqa: E731
        self._dgemm = g("cblas_dgemm")
        self._dgemm.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]
        self._daxpy = g("cblas_daxpy"); self._daxpy.argtypes = [c_int, c_double, _DP, c_int, _DP, c_int]
        self._dcopy = g("cblas_dcopy"); self._dcopy.argtypes = [c_int, _DP, c_int, _DP, c_int]
        self._dscal = g("cblas_dscal"); self._dscal.argtypes = [c_int, c_double, _DP, c_int]
        self._dlange = g("dlange_"); self._dlange.argtypes = [POINTER(c_char), _IP, _IP, _DP, _IP, _DP]
        self._dlange.restype = c_double
        self._dgetrf = g("dgetrf_"); self._dgetrf.argtypes = [_IP, _IP, _DP, _IP, _IP, _IP]
        self._dgesv = g("dgesv_"); self._dgesv.argtypes = [_IP, _IP, _DP, _IP, _IP, _DP, _IP, _IP]
        self._dgetri = g("dgetri_"); self._dgetri.argtypes = [_IP, _DP, _IP, _IP, _DP, _IP, _IP]
        self._set = g("openblas_set_num_threads"); self._set.argtypes = [c_int]
        self._get = g("openblas_get_num_threads"); self._get.restype = c_int
        if threads:
            self._set(threads)

    def threads(self):
        return int(self._get())

    def dgemm(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        self._dgemm(COL_MAJOR, ta, tb, m, n, k, alpha, _dp(a), lda, _dp(b), ldb, beta, _dp(c), ldc)

    def daxpy(self, n, alpha, x, incx, y, incy):
        self._daxpy(n, alpha, x, incx, y, incy)

    def dcopy(self, n, x, incx, y, incy):
        self._dcopy(n, x, incx, y, incy)

    def dscal(self, n, alpha, x, incx):
        self._dscal(n, alpha, x, incx)

    def dlange(self, norm: bytes, m, n, a, lda, work):
        return self._dlange(ctypes.c_char_p(norm), byref(c_int(m)), byref(c_int(n)), _dp(a), byref(c_int(lda)),
                            _dp(work) if work is not None else None)

    def dgetrf(self, m, n, a, lda, ipiv):
        info = c_int(0)
        self._dgetrf(byref(c_int(m)), byref(c_int(n)), _dp(a), byref(c_int(lda)), ipiv.ctypes.data_as(_IP), byref(info))
        return info.value

    def dgesv(self, n, nrhs, a, lda, ipiv, b, ldb):
        info = c_int(0)
        self._dgesv(byref(c_int(n)), byref(c_int(nrhs)), _dp(a), byref(c_int(lda)), ipiv.ctypes.data_as(_IP), _dp(b),
                    byref(c_int(ldb)), byref(info))
        return info.value

    def dgetri(self, n, a, lda, ipiv, work, lwork):
        info = c_int(0)
        self._dgetri(byref(c_int(n)), _dp(a), byref(c_int(lda)), ipiv.ctypes.data_as(_IP),
                     _dp(work) if work is not None else None, byref(c_int(lwork)), byref(info))
        return info.value

    def ewise(self, op, m, n, a, lda, sa, b, ldb, sb, c, ldc):
        # The pointwise macros are Racket loops, not BLAS (flomat.rkt:3021-3146): numpy does the same
        # single IEEE operation per element.
        x = a if a is not None else sa
        y = b if b is not None else sb
        c[...] = (np.add, np.subtract, np.multiply, np.divide)[op](x, y)


def default_provider(prefer: str = "openblas") -> _Provider:
    if prefer == "openblas":
        try:
            return OpenBLASProvider()
        except (FileNotFoundError, OSError, ImportError, AttributeError):
            pass
    return CProvider()


# --------------------------------------------------------------------------------------------------
# Pade-8 coefficients, expm.rkt:134-147: c_i = (16-i)! 8! / (16! i! (8-i)!), exact rational -> double
# --------------------------------------------------------------------------------------------------
def pade8_coefs() -> list[float]:
    from fractions import Fraction
    from math import factorial as f

    p = 8
    return [float(Fraction(f(2 * p - i) * f(p), f(2 * p) * f(i) * f(p - i))) for i in range(p + 1)]


PADE8 = pade8_coefs()


# --------------------------------------------------------------------------------------------------
# The reference's call sequences
# --------------------------------------------------------------------------------------------------
class Flomat:
    """Racket-level hot-path functions restated over a provider.  Inputs/outputs: (m,n) float64 F-order."""

    def __init__(self, provider: _Provider | None = None):
        self.p = provider or default_provider()
        self.gemm_calls = 0

    # -- alloc/copy: flomat.rkt:437, :516, :538 ----------------------------------------------------
    @staticmethod
    def make(m, n, x=0.0):
        return np.full((m, n), float(x), dtype=np.float64, order="F")  # make-flomat :538

    @staticmethod
    def copy(a):
        return np.array(a, dtype=np.float64, order="F", copy=True)  # copy-flomat :516

    @staticmethod
    def eye(m, n=None):
        return np.asfortranarray(np.eye(m, n if n is not None else m))  # flomat-eye :2642

    # -- times: flomat.rkt:3202 -> flomat* :905 -> :877 -> cblas_dgemm :888 (alpha=beta=1, C zeroed) --
    def gemm(self, a, b, c=None, alpha=1.0, beta=1.0, ta=False, tb=False):
        m, ka = (a.shape[1], a.shape[0]) if ta else a.shape
        kb, n = (b.shape[1], b.shape[0]) if tb else b.shape
        if ka != kb:
            raise ValueError("check-product-dimensions (flomat.rkt:340)")
        if c is None:
            c = self.make(m, n)
        self.p.dgemm(TRANS if ta else NO_TRANS, TRANS if tb else NO_TRANS, m, n, ka, float(alpha), a, a.shape[0],
                     b, b.shape[0], float(beta), c, c.shape[0])
        self.gemm_calls += 1
        return c

    def times(self, a, *bs):
        for b in bs:  # left fold, flomat.rkt:3202-3212
            if np.isscalar(a) and np.isscalar(b):
                a = a * b
            elif np.isscalar(a) or np.isscalar(b):
                a = self.dot_times(a, b)
            else:
                a = self.gemm(a, b)
        return a

    # -- plus / minus: flomat.rkt:3158, :3180 -> flomat+ :826, flomat- :836 -> :815/:805 (n daxpy) ---
    def _axpy_cols(self, alpha, a, b):
        """constant*flomat+flomat! (flomat.rkt:805): B := alpha*A + B, one daxpy per column."""
        m, n = a.shape
        for j in range(n):
            self.p.daxpy(m, float(alpha), _dp(a[:, j]), 1, _dp(b[:, j]), 1)
        return b

    def plus(self, a, *bs):
        for b in bs:
            if np.isscalar(a) or np.isscalar(b):
                a = self.dot_plus(a, b)
            else:
                if a.shape != b.shape:
                    raise ValueError("check-same-dimensions (flomat.rkt:326)")
                a = self._axpy_cols(1.0, a, self.copy(b))  # flomat+ = copy B then += 1*A
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
                a = self._axpy_cols(-1.0, b, self.copy(a))  # flomat- = copy A then += -1*B
        return a

    def scale(self, s, a):
        """flomat-scale (flomat.rkt:1036) -> dscal per matrix (lda = m) :1026."""
        c = self.copy(a)
        self.p.dscal(c.size, float(s), _dp(c), 1)
        return c

    # -- pointwise macros: flomat.rkt:3021-3146 --------------------------------------------------------
    def _ew(self, op, a, b):
        sa = sb = 0.0
        A = B = None
        if np.isscalar(a):
            sa = float(a)
        else:
            A = a
        if np.isscalar(b):
            sb = float(b)
        else:
            B = b
        ref = A if A is not None else B
        if A is not None and B is not None and A.shape != B.shape:
            raise ValueError("check-same-dimensions (flomat.rkt:326)")
        m, n = ref.shape
        c = np.empty((m, n), dtype=np.float64, order="F")
        self.p.ewise(op, m, n, A, A.shape[0] if A is not None else 1, sa, B, B.shape[0] if B is not None else 1, sb, c, m)
        return c

    def dot_plus(self, a, b):
        return self._ew(0, a, b)

    def dot_minus(self, a, b=None):
        if b is None:  # unary .- : (f aij) with f = - (flomat.rkt:3123-3133) -> negate
            return -self.copy(a)
        return self._ew(1, a, b)

    def dot_times(self, a, b):
        return self._ew(2, a, b)

    def dot_divide(self, a, b=None):
        if b is None:  # unary ./ : reciprocal
            return self._ew(3, 1.0, a)
        return self._ew(3, a, b)

    # -- norm: flomat.rkt:3273 -> :1328 -> dlange_ :1340 -------------------------------------------------
    def norm(self, a, kind="frob"):
        code = {1: b"1", "1": b"1", "inf": b"I", "frob": b"F", "max": b"M"}[kind]
        m, n = a.shape
        work = np.zeros(max(1, m) if kind == "inf" else 1)
        return float(self.p.dlange(code, m, n, a, a.shape[0], work))

    # -- LU family -----------------------------------------------------------------------------------
    def lu(self, a):
        """flomat-lu! on a copy (flomat.rkt:1525): returns (LU, ipiv 1-based int32, info)."""
        lu = self.copy(a)
        m, n = lu.shape
        ipiv = np.zeros(max(1, min(m, n)), dtype=np.int32)
        info = self.p.dgetrf(m, n, lu, m, ipiv)
        return lu, ipiv, info

    def mldivide(self, a, b):
        """mldivide (flomat.rkt:3289) -> flomat-solve-many :2145: copy A, copy B, dgesv, info ignored."""
        lu, x = self.copy(a), self.copy(b)
        n, nrhs = lu.shape[0], x.shape[1]
        ipiv = np.zeros(max(1, n), dtype=np.int32)
        self.last_info = self.p.dgesv(n, nrhs, lu, n, ipiv, x, x.shape[0])
        return x

    def inv(self, a):
        """inv (flomat.rkt:3297) -> flomat-inverse! :1770: work = copy of A, dgetrf, dgetri lwork=m*m."""
        lu = self.copy(a)
        n = lu.shape[0]
        work = self.copy(a)
        ipiv = np.zeros(max(1, n), dtype=np.int32)
        self.p.dgetrf(n, n, lu, n, ipiv)
        self.last_info = self.p.dgetri(n, lu, n, ipiv, work, n * n)
        return lu

    def det(self, a):
        """det (flomat.rkt:3276) -> :1844 -> :1837: pivots-sign :1487 times the left-to-right diagonal product."""
        lu, ipiv, _ = self.lu(a)
        sign = 1
        for i in range(lu.shape[0]):
            if ipiv[i] - 1 != i:
                sign = -sign
        prod = 1.0
        for i in range(lu.shape[0]):
            prod = prod * float(lu[i, i])
        return sign * prod

    # -- neighbours of the hot path: restated from the Racket loops / BLAS call sites ----------
    @staticmethod
    def transpose(a):
        """flomat-transpose (flomat.rkt:1360-1369): AT[j, i] = A[i, j]."""
        return fortran(np.asarray(a).T.copy())

    def mrdivide(self, b, a):
        """mrdivide (flomat.rkt:3293-3295): (transpose (mldivide (transpose A) (transpose B)))."""
        return self.transpose(self.mldivide(self.transpose(a), self.transpose(b)))

    @staticmethod
    def extract_upper(a):
        """flomat-extract-upper (flomat.rkt:1540-1551): k x k, entries with i <= j."""
        k = min(a.shape)
        u = np.zeros((k, k), order="F")
        for j in range(k):
            for i in range(min(j + 1, k)):
                u[i, j] = a[i, j]
        return u

    @staticmethod
    def extract_lower(a, ones=False):
        """flomat-extract-lower (flomat.rkt:1566-1573) / flomat-extract-lower/ones (:1553-1564)."""
        l = fortran(np.array(a, dtype=np.float64))
        m, n = l.shape
        for j in range(n):
            for i in range(min(j, m)):
                l[i, j] = 0.0
        if ones:
            for j in range(min(m, n)):
                l[j, j] = 1.0
        return l

    @staticmethod
    def pivots_to_matrix(ipiv):
        """pivots->flomat (flomat.rkt:1471-1484): identity, then for i = k-1 .. 0 swap rows i and ipiv[i] (0-based there)."""
        k = len(ipiv)
        p = np.zeros((k, k), order="F")
        for i in range(k):
            p[i, i] = 1.0
        for i in range(k - 1, -1, -1):
            i2 = int(ipiv[i]) - 1
            if i != i2:
                p[[i, i2], :] = p[[i2, i], :]
        return p

    def plu(self, a):
        """flomat-plu (flomat.rkt:1529-1536)."""
        lu, ipiv, _ = self.lu(a)
        return self.pivots_to_matrix(ipiv[: min(a.shape)]), self.extract_lower(lu, ones=True), self.extract_upper(lu)

    @staticmethod
    def times_vector(a, x, y=None, alpha=1.0, beta=1.0, transpose_a=False):
        """flomat*vector (flomat.rkt:970-978) -> cblas_dgemv (:949): Y := alpha op(A) X + beta Y, Y = zeros when omitted."""
        op = a.T if transpose_a else a
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        acc = np.zeros(op.shape[0])
        for j in range(op.shape[1]):  # reference DGEMV (no-transpose form): column sweeps
            acc += op[:, j] * x[j]
        y0 = np.zeros(op.shape[0]) if y is None else np.asarray(y, dtype=np.float64).reshape(-1)
        return fortran((alpha * acc + beta * y0).reshape(-1, 1))

    @staticmethod
    def dot(x, y):
        """dot / flcolumn-dot (flomat.rkt:3216, :1926) -> cblas_ddot: sequential sum of products."""
        acc = 0.0
        for u, v in zip(np.asarray(x).reshape(-1), np.asarray(y).reshape(-1)):
            acc += float(u) * float(v)
        return acc

    @staticmethod
    def column_norm(v):
        """flcolumn-norm (flomat.rkt:1941) -> cblas_dnrm2 (scaled sum of squares)."""
        x = np.abs(np.asarray(v, dtype=np.float64).reshape(-1))
        s = x.max() if x.size else 0.0
        return float(s * math.sqrt(((x / s) ** 2).sum())) if s > 0 else 0.0

    @staticmethod
    def max_abs_index(a):
        """flomat-max-abs-index (flomat.rkt:1163-1178): idamax over the column-major buffer, first maximum wins."""
        flat = np.abs(np.asarray(a, dtype=np.float64)).reshape(-1, order="F")
        idx = int(np.argmax(flat))
        return idx % a.shape[0], idx // a.shape[0]

    def colsums(self, a):
        """flomat-column-sums (flomat.rkt:2094): 1 x n, each entry `unsafe-sum` = ddot against a 1.0 with stride 0 (:2066)."""
        return fortran(np.array([[self.dot(a[:, j], np.ones(a.shape[0])) for j in range(a.shape[1])]]))

    def rowsums(self, a):
        """flomat-row-sums (flomat.rkt:2101): m x 1."""
        return fortran(np.array([[self.dot(a[i, :], np.ones(a.shape[1]))] for i in range(a.shape[0])]))

    @staticmethod
    def pointwise(fn, a, b=None):
        """define-pointwise-unaries / -binaries (flomat.rkt:2990-3076): f applied per element with the C library's libm
        (Racket's flsin, flcos, ... call it), `sqr` = x*x, `expt` = pow."""
        f = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "sqr": lambda x: x * x, "sqrt": math.sqrt, "log": math.log,
             "exp": math.exp, "expt": math.pow}[fn]
        a = np.asarray(a, dtype=np.float64)
        out = np.empty(a.shape, order="F")
        if b is None:
            for idx in np.ndindex(a.shape):
                out[idx] = f(float(a[idx]))
        else:
            bb = np.broadcast_to(np.asarray(b, dtype=np.float64), a.shape)
            for idx in np.ndindex(a.shape):
                out[idx] = f(float(a[idx]), float(bb[idx]))
        return out

    @staticmethod
    def cholesky(a, upper=False):
        """flomat-cholesky (flomat.rkt:1821-1826) -> dpotrf_: unblocked DPOTF2 restated (lower form; upper = its transpose),
        then the requested triangle is extracted.  Returns (factor, info)."""
        n = a.shape[0]
        l = np.tril(np.array(a.T if upper else a, dtype=np.float64))
        info = 0
        for j in range(n):
            d = l[j, j]
            if not d > 0.0:
                info = j + 1
                break
            l[j, j] = math.sqrt(d)
            l[j + 1:, j] /= l[j, j]
            for c in range(j + 1, n):
                l[c:, c] -= l[c:, j] * l[c, j]
        return (fortran(l.T.copy()) if upper else fortran(l)), info

    # -- dlarnv (flomat.rkt:2798-2852): LAPACK's DLARUV multiplicative congruential generator restated in exact integers ----
    LCG_A = 33952834046453  # dlaruv.f: MM(1,:) = (494, 322, 2508, 2549) in base 4096; row i of the table is a^i mod 2^48

    @classmethod
    def dlarnv(cls, idist, iseed, n):
        """DLARNV: returns (x, new_iseed).  x_k = (s0 a^k mod 2^48) 2^-48 (DLARUV), idist 2: 2u-1, idist 3:
        sqrt(-2 ln u_{2k-1}) cos(2 pi u_{2k}); the seed advances by one multiplication per uniform drawn."""
        s = (int(iseed[0]) << 36) | (int(iseed[1]) << 24) | (int(iseed[2]) << 12) | int(iseed[3])
        mask = (1 << 48) - 1
        x = np.empty(n)
        twopi = 6.28318530717958647692528676655900576839
        for k in range(n):
            s = (s * cls.LCG_A) & mask
            u = s * 2.0 ** -48
            if idist == 1:
                x[k] = u
            elif idist == 2:
                x[k] = 2.0 * u - 1.0
            else:
                s = (s * cls.LCG_A) & mask
                u2 = s * 2.0 ** -48
                x[k] = math.sqrt(-2.0 * math.log(u)) * math.cos(twopi * u2)
        return x, np.array([(s >> 36) & 4095, (s >> 24) & 4095, (s >> 12) & 4095, s & 4095], dtype=np.int32)

    # -- expm: expm.rkt:167-215 ----------------------------------------------------------------------------
    def mpoly(self, cs, a):
        """mpoly (expm.rkt:167-177): Horner seeded with the zero matrix; (plus (.* c I) (times A rest))."""
        n = a.shape[0]
        ident, zero = self.eye(n), self.make(n, n)

        def loop(cs):
            if not cs:
                return zero
            return self.plus(self.dot_times(cs[0], ident), self.times(a, loop(cs[1:])))

        return loop(list(cs))

    def scaling_exponent(self, a):
        """find-scaling-exponent (expm.rkt:212-215): s = 1 + ceiling(log N / log 2), a flonum; may be <= 0."""
        nrm = self.norm(a, 1)
        if nrm == 0.0:
            return -math.inf
        if math.isnan(nrm) or math.isinf(nrm):
            return nrm
        return 1.0 + math.ceil(math.log(nrm) / math.log(2.0))

    def expm_pade_no_scaling(self, x):
        """expm-pade/no-scaling (expm.rkt:193-199)."""
        c = PADE8
        aa = self.times(x, x)
        even = self.mpoly(c[0::2], aa)
        odd = self.times(x, self.mpoly(c[1::2], aa))
        num = self.plus(even, odd)
        den = self.minus(even, odd)
        return self.mldivide(den, num)

    def expm(self, a):
        """expm (expm.rkt:203-210), quirks included: s<=0 scales A *up* and never squares."""
        s = self.scaling_exponent(a)
        if math.isinf(s) or math.isnan(s):
            # reference: (expt 2 -inf)=0 -> A/0 = NaN matrix (zero input); inf/NaN norm never terminates.
            with np.errstate(all="ignore"):
                two_s = 0.0 if s < 0 else math.inf
                x = self.dot_divide(a, two_s)
            if s > 0 or math.isnan(s):
                raise FloatingPointError("reference does not terminate for inf/NaN norm (expm.rkt:209)")
            return self.expm_pade_no_scaling(x)
        x = self.dot_divide(a, 2.0 ** s)
        r = self.expm_pade_no_scaling(x)
        k = s
        while k > 0:  # repeated-squaring (expm.rkt:209-210)
            r = self.times(r, r)
            k -= 1
        return r


def normwise_rel_err(x, ref) -> float:
    """||x - ref||_1 / ||ref||_1 (SURVEY 8d parity gate)."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).sum(axis=0).max() if ref.size else 0.0
    num = np.abs(x - ref).sum(axis=0).max() if ref.size else 0.0
    return float(num / den) if den > 0 else float(num)


def tol(n: int) -> float:
    """north_star tolerance: 64 * n * eps."""
    return 64.0 * n * 2.0 ** -52
