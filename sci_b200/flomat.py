"""Host-side mirror of flomat's high-level API for the dense FP64 hot path, over the libflomat_cuda C ABI.

Racket is not available in this image, so this module plays the role of the (replaced) binding layer plus the thin
Racket wrappers above it: same function names, argument meaning and error behaviour as the reference, with the
matrix buffer living in HBM.  File:line citations are into /root/reference/flomat/.

    struct flomat (m n a lda)        flomat.rkt:219      -> class Flomat (a = device address)
    alloc/copy/make/zeros/ones/eye   :437 :516 :538 :2913 :2642
    shared-submatrix!  / sub!        :295                 -> Flomat.sub (a view: same buffer, lda kept)
    times / plus / minus             :3202 :3158 :3180
    .+ .- .* ./                      :3021-3146           -> dot_plus dot_minus dot_times dot_div
    norm det inv mldivide power      :3273 :3276 :3297 :3289 :3220
    expm                             expm.rkt:203

Explicit host sync points: `matrix()` / `Flomat.from_numpy` upload, `Flomat.to_numpy()` / `ref` download.
Nothing here computes on the CPU and there is no fallback: without the CUDA library every call raises.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int32

import math

import numpy as np

from . import _lib
from ._lib import FlomatCudaError, check

NO_TRANS, TRANS = 111, 112
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_EXPT = 0, 1, 2, 3, 4
MAP_SIN, MAP_COS, MAP_TAN, MAP_SQR, MAP_SQRT, MAP_LOG, MAP_EXP = 5, 6, 7, 8, 9, 10, 11


def _L():
    lib = _lib.load()
    return lib


class _Buffer:
    """Owns one device allocation (the reference leaves this to the GC: flomat.rkt:437-441)."""

    def __init__(self, nbytes: int):
        self.ptr = _L().fm_alloc_bytes(max(int(nbytes), 16))
        if not self.ptr:
            raise FlomatCudaError(f"fm_alloc_bytes({nbytes}) failed: {_lib.last_error()}")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                _L().fm_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _host_layout_ok(h, m: int, n: int) -> bool:
    """A host array the async transfers can use in place: float64, either column-major with the matrix's shape or a flat
    contiguous buffer of m*n elements holding the column-major bytes."""
    if not isinstance(h, np.ndarray) or h.dtype != np.float64:
        return False
    if h.ndim == 2 and h.shape == (m, n) and h.flags.f_contiguous:
        return True
    return h.ndim == 1 and h.size == m * n and h.flags.c_contiguous


class Flomat:
    """`(struct flomat (m n a lda))`: m rows, n columns, device address `a`, leading dimension lda (column-major)."""

    __slots__ = ("m", "n", "a", "lda", "_buf")

    def __init__(self, m: int, n: int, a: int, lda: int, buf):
        self.m, self.n, self.a, self.lda, self._buf = int(m), int(n), int(a), int(lda), buf

    # ---- construction -----------------------------------------------------------------------------
    @staticmethod
    def alloc(m: int, n: int) -> "Flomat":
        """alloc-flomat (flomat.rkt:437): uninitialised m x n, lda = m."""
        if m < 0 or n < 0:
            raise ValueError("alloc-flomat: expected non-negative dimensions")
        buf = _Buffer(m * n * 8)
        return Flomat(m, n, buf.ptr, max(m, 1), buf)

    @staticmethod
    def from_numpy(x) -> "Flomat":
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(-1, 1)
        if x.ndim != 2:
            raise ValueError("matrix: expected a 2-d array")
        h = np.asfortranarray(x)
        out = Flomat.alloc(*h.shape)
        if h.size:
            check(_L().fm_upload(out.a, out.lda, h.ctypes.data, max(h.shape[0], 1), h.shape[0], h.shape[1]), "fm_upload")
            check(_L().fm_sync(), "fm_sync")  # h may be a temporary
        return out

    def to_numpy(self) -> np.ndarray:
        h = np.empty((self.m, self.n), dtype=np.float64, order="F")
        if h.size:
            check(_L().fm_download(h.ctypes.data, max(self.m, 1), self.a, self.lda, self.m, self.n), "fm_download")
        return h

    # ---- asynchronous transfers (library transfer stream): for pinned host arrays that outlive the transfer ----------
    def upload_async(self, h: np.ndarray) -> "Flomat":
        """Start copying the column-major host array `h` (m x n float64, pinned for a real overlap) into this matrix.  The copy
        may overlap kernels queued before the call; everything queued after it sees the data.  `h` must stay alive and
        unchanged until transfer_sync() / sync()."""
        if not _host_layout_ok(h, self.m, self.n):
            raise ValueError("upload_async: expected a column-major float64 array of the matrix's shape (or its flat column-major bytes)")
        if h.size:
            check(_L().fm_upload_async(self.a, self.lda, h.ctypes.data, max(self.m, 1), self.m, self.n), "fm_upload_async")
        return self

    def download_async(self, h: np.ndarray) -> None:
        """Start copying this matrix into the column-major host array `h` once the work queued so far is done; runs beside
        whatever is queued afterwards.  `h` is valid after transfer_sync() / sync(); keep this matrix alive until then."""
        if not _host_layout_ok(h, self.m, self.n) or not h.flags.writeable:
            raise ValueError("download_async: expected a writable column-major float64 array of the matrix's shape (or a flat buffer of m*n)")
        if h.size:
            check(_L().fm_download_async(h.ctypes.data, max(self.m, 1), self.a, self.lda, self.m, self.n), "fm_download_async")

    # ---- views (shared-submatrix!, flomat.rkt:295-300): pointer arithmetic only -----------------------
    def sub(self, i: int, j: int, m: int, n: int) -> "Flomat":
        if not (0 <= i and 0 <= j and m >= 0 and n >= 0 and i + m <= self.m and j + n <= self.n):
            raise IndexError("shared-submatrix!: index out of range")
        return Flomat(m, n, self.a + 8 * (i + j * self.lda), self.lda, self._buf)

    # ---- element access (unsafe-ref / unsafe-set!, flomat.rkt:986-1002): one host sync each ------------
    def ref(self, i: int, j: int) -> float:
        if not (0 <= i < self.m and 0 <= j < self.n):
            raise IndexError("flomat-ref: index out of range")
        out = c_double()
        check(_L().fm_get(self.a, i + j * self.lda, byref(out)), "fm_get")
        return out.value

    def set(self, i: int, j: int, x: float) -> None:
        if not (0 <= i < self.m and 0 <= j < self.n):
            raise IndexError("flomat-set!: index out of range")
        check(_L().fm_set(self.a, i + j * self.lda, float(x)), "fm_set")

    @property
    def shape(self):
        return (self.m, self.n)

    def __repr__(self):
        return f"Flomat({self.m}x{self.n}, lda={self.lda}, a=0x{self.a:x})"


# --------------------------------------------------------------------------------------------------
# checks (flomat.rkt:322-402) -- error text follows the reference's
# --------------------------------------------------------------------------------------------------
def _check_same(a: Flomat, b: Flomat, who: str):
    if a.m != b.m or a.n != b.n:
        raise ValueError(f"{who}: expected two matrices of the same size, got {a.m}x{a.n} and {b.m}x{b.n}")


def _check_square(a: Flomat, who: str):
    if a.m != a.n:
        raise ValueError(f"{who}: expected square matrix, got {a.m}x{a.n}")


def _is_num(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


# --------------------------------------------------------------------------------------------------
# constructors
# --------------------------------------------------------------------------------------------------
def matrix(x) -> Flomat:
    """`(matrix obj)` (flomat.rkt:2880): upload a nested list / numpy array (row-major literal, stored column-major)."""
    return x if isinstance(x, Flomat) else Flomat.from_numpy(x)


def make_flomat(m: int, n: int, x: float = 0.0) -> Flomat:
    """make-flomat (flomat.rkt:538)."""
    out = Flomat.alloc(m, n)
    check(_L().fm_fill(out.a, out.lda, m, n, float(x)), "fm_fill")
    return out


def zeros(m: int, n: int | None = None) -> Flomat:
    return make_flomat(m, m if n is None else n, 0.0)


def ones(m: int, n: int | None = None) -> Flomat:
    return make_flomat(m, m if n is None else n, 1.0)


def eye(m: int, n: int | None = None, k: int = 0) -> Flomat:
    """eye / flomat-eye (flomat.rkt:3303, :2642)."""
    n = m if n is None else n
    out = Flomat.alloc(m, n)
    check(_L().fm_eye(out.a, out.lda, m, n, int(k)), "fm_eye")
    return out


def copy_flomat(a: Flomat) -> Flomat:
    """copy-flomat (flomat.rkt:516): dense copy (lda = m) of a matrix or view."""
    out = Flomat.alloc(a.m, a.n)
    check(_L().fm_copy2d(out.a, out.lda, a.a, a.lda, a.m, a.n), "fm_copy2d")
    return out


def nrows(a: Flomat) -> int:
    return a.m


def ncols(a: Flomat) -> int:
    return a.n


# --------------------------------------------------------------------------------------------------
# pointwise (.+ .- .* ./ and their ! forms, flomat.rkt:3021-3146)
# --------------------------------------------------------------------------------------------------
def _ewise(op: int, a, b, c: Flomat | None, who: str) -> Flomat:
    if isinstance(a, Flomat) and isinstance(b, Flomat):
        _check_same(a, b, who)
        ref = a
    elif isinstance(a, Flomat) and _is_num(b):
        ref = a
    elif _is_num(a) and isinstance(b, Flomat):
        ref = b
    else:
        raise TypeError(f"{who}: wrong input types")
    if c is None:
        c = Flomat.alloc(ref.m, ref.n)
    else:
        _check_same(ref, c, who)
    pa, lda, sa = (a.a, a.lda, 0.0) if isinstance(a, Flomat) else (None, 1, float(a))
    pb, ldb, sb = (b.a, b.lda, 0.0) if isinstance(b, Flomat) else (None, 1, float(b))
    check(_L().fm_ewise(op, ref.m, ref.n, pa, lda, sa, pb, ldb, sb, c.a, c.lda), who)
    return c


def dot_plus(a, b, c: Flomat | None = None) -> Flomat:
    """`.+` / `.+!` (pass c=a for the ! form)."""
    return _ewise(OP_ADD, a, b, c, ".+")


def dot_times(a, b, c: Flomat | None = None) -> Flomat:
    """`.*`."""
    return _ewise(OP_MUL, a, b, c, ".*")


def dot_minus(a, b=None, c: Flomat | None = None) -> Flomat:
    """`.-` binary, or unary negation (flomat.rkt:3123-3133)."""
    if b is None:
        return _ewise(OP_MUL, a, -1.0, c, ".-")  # (- aij): exact negation, signed zeros included
    return _ewise(OP_SUB, a, b, c, ".-")


def dot_div(a, b=None, c: Flomat | None = None) -> Flomat:
    """`./` binary, or unary reciprocal."""
    if b is None:
        return _ewise(OP_DIV, 1.0, a, c, "./")
    return _ewise(OP_DIV, a, b, c, "./")


# --------------------------------------------------------------------------------------------------
# plus / minus (flomat.rkt:3148-3190 -> flomat+ :826, flomat- :836 -> constant*flomat+flomat :805-819)
# --------------------------------------------------------------------------------------------------
def _axpy_into_copy(alpha: float, a: Flomat, b: Flomat, who: str) -> Flomat:
    """constant*flomat+flomat (flomat.rkt:815): copy B, then B' := alpha*A + B' -- fused here into ONE pass that
    reads A and B once and writes the result once (24 B/element instead of the reference's 40)."""
    _check_same(a, b, who)
    op = OP_ADD if alpha == 1.0 else OP_SUB
    # alpha = +1: B + A ; alpha = -1: B - A   (one IEEE add/sub per element, identical to daxpy with alpha = +-1)
    return _ewise(op, b, a, None, who)


def flomat_plus(a: Flomat, b: Flomat) -> Flomat:
    """flomat+ (flomat.rkt:826): A + B = copy B then += 1.0*A."""
    return _axpy_into_copy(1.0, a, b, "flomat+")


def flomat_minus(a: Flomat, b: Flomat | None = None) -> Flomat:
    """flomat- (flomat.rkt:836): A - B = copy A then += -1.0*B; unary: scale by -1 (dscal)."""
    if b is None:
        return flomat_scale(-1.0, a)
    return _axpy_into_copy(-1.0, b, a, "flomat-")


def flomat_plus_bang(a: Flomat, b: Flomat) -> Flomat:
    """flomat+! (flomat.rkt:821): B := A + B in place (n daxpys in the reference, one kernel here)."""
    _check_same(a, b, "flomat+!")
    check(_L().fm_axpy2d(a.m, a.n, 1.0, a.a, a.lda, b.a, b.lda), "fm_axpy2d")
    return b


def constant_times_flomat_plus_flomat_bang(alpha: float, a: Flomat, b: Flomat) -> Flomat:
    """constant*flomat+flomat! (flomat.rkt:805-813): B := alpha*A + B in place (n cblas_daxpy in the reference, one kernel here
    with daxpy's arithmetic per element)."""
    _check_same(a, b, "constant*flomat+flomat!")
    check(_L().fm_axpy2d(a.m, a.n, float(alpha), a.a, a.lda, b.a, b.lda), "fm_axpy2d")
    return b


def constant_times_flomat_plus_flomat(alpha: float, a: Flomat, b: Flomat) -> Flomat:
    """constant*flomat+flomat (flomat.rkt:815-819): alpha*A + B on a copy of B."""
    if alpha in (1.0, -1.0):
        return _axpy_into_copy(float(alpha), a, b, "constant*flomat+flomat")
    return constant_times_flomat_plus_flomat_bang(alpha, a, copy_flomat(b))


def flomat_minus_bang(a: Flomat, b: Flomat) -> Flomat:
    """flomat-! (flomat.rkt:831-834): A := A - B in place, as B' := -1.0*B + A with the roles the reference gives them."""
    _check_same(a, b, "flomat-!")
    return constant_times_flomat_plus_flomat_bang(-1.0, b, a)


def flomat_scale_bang(s: float, a: Flomat) -> Flomat:
    """flomat-scale! (flomat.rkt:1032) -> constant*matrix! (:1020-1030): A := s*A in place (cblas_dscal)."""
    check(_L().fm_scal2d(a.m, a.n, float(s), a.a, a.lda), "fm_scal2d")
    return a


def plus_bang(a: Flomat, *bs):
    """`plus!` (flomat.rkt:3148-3156): every matrix B is added into A (`flomat+! B A`), every number through `.+!`; returns A."""
    if not isinstance(a, Flomat):
        raise TypeError("plus!: expected a flomat as first argument")  # check-flomat
    _check_all_same([x for x in (a, *bs) if isinstance(x, Flomat)], "plus!")
    for b in bs:
        if isinstance(b, Flomat):
            flomat_plus_bang(b, a)
        else:
            dot_plus(a, b, a)
    return a


def minus_bang(a: Flomat, *bs):
    """`minus!` (flomat.rkt:3168-3178).  No further argument: A := -A (`.-!`).  A number: A := A - x.  A MATRIX argument B is handled by
    `(flomat-! B A)`, which is B := B - A (flomat.rkt:831): the reference subtracts into the OPERAND and returns A unchanged -- reproduced
    here as it stands (a caller that wants A := A - B calls flomat_minus_bang(A, B))."""
    if not isinstance(a, Flomat):
        raise TypeError("minus!: expected a flomat as first argument")
    _check_all_same([x for x in (a, *bs) if isinstance(x, Flomat)], "minus!")
    if not bs:
        return dot_minus(a, None, a)
    for b in bs:
        if isinstance(b, Flomat):
            flomat_minus_bang(b, a)
        else:
            dot_minus(a, b, a)
    return a


def flomat_scale(s: float, a: Flomat) -> Flomat:
    """flomat-scale (flomat.rkt:1036) = copy + dscal; fused into one s*a_ij pass."""
    return _ewise(OP_MUL, float(s), a, None, "flomat-scale")


def plus(a, *bs):
    """`plus` (flomat.rkt:3158): left fold over matrices and numbers."""
    _check_all_same([x for x in (a, *bs) if isinstance(x, Flomat)], "plus")
    for b in bs:
        if isinstance(a, Flomat) and isinstance(b, Flomat):
            a = flomat_plus(a, b)
        else:
            a = dot_plus(a, b) if (isinstance(a, Flomat) or isinstance(b, Flomat)) else a + b
    return a


def minus(a, *bs):
    """`minus` (flomat.rkt:3180): unary negation or left-folded difference."""
    _check_all_same([x for x in (a, *bs) if isinstance(x, Flomat)], "minus")
    if not bs:
        return dot_minus(a)
    for b in bs:
        if isinstance(a, Flomat) and isinstance(b, Flomat):
            a = flomat_minus(a, b)
        else:
            a = dot_minus(a, b) if (isinstance(a, Flomat) or isinstance(b, Flomat)) else a - b
    return a


def _check_all_same(ms, who):
    """check-all-matrices-same-size (flomat.rkt:330)."""
    for x in ms[1:]:
        _check_same(ms[0], x, who)


# --------------------------------------------------------------------------------------------------
# times (flomat.rkt:3202 -> flomat* :905 -> constant*matrix*matrix+constant*matrix! :877 -> cblas_dgemm :888)
# --------------------------------------------------------------------------------------------------
def flomat_times(a: Flomat, b: Flomat, c: Flomat | None = None, alpha: float = 1.0, beta: float = 1.0,
                 transpose_a: bool = False, transpose_b: bool = False) -> Flomat:
    """flomat* (flomat.rkt:905): C := alpha*op(A)*op(B) + beta*C.  Without C the reference allocates a ZERO matrix and
    calls dgemm with beta = 1 (:909, :538); that equals beta = 0 on an uninitialised buffer, which is what runs here
    (saves the memset and the read of C)."""
    ma, na = (a.n, a.m) if transpose_a else (a.m, a.n)
    mb, nb = (b.n, b.m) if transpose_b else (b.m, b.n)
    if na != mb:
        raise ValueError(f"flomat*: expected two matrices with compatible dimensions, got {ma}x{na} and {mb}x{nb}")
    if c is None:
        c = Flomat.alloc(ma, nb)
        beta = 0.0
    elif c.m != ma or c.n != nb:
        raise ValueError("flomat*!: result matrix has the wrong dimensions")
    check(_L().fm_dgemm(TRANS if transpose_a else NO_TRANS, TRANS if transpose_b else NO_TRANS, ma, nb, na, float(alpha),
                        a.a, a.lda, b.a, b.lda, float(beta), c.a, c.lda), "fm_dgemm")
    return c


def times_host(a: np.ndarray, b: np.ndarray, c: np.ndarray | None = None, alpha: float = 1.0, beta: float = 1.0,
               wait: bool = True) -> np.ndarray:
    """flomat* (flomat.rkt:905) on HOST matrices, the call the unmodified reference makes (cblas_dgemm with GC'd host memory,
    :888): C := alpha*A*B + beta*C through fm_dgemm_host, which moves B and C in column blocks so that upload, product and
    download overlap.  a, b, c are column-major float64 arrays (pinned memory for a real overlap).  Without c a new array is
    returned and beta is 0 (the reference's zero matrix with beta = 1).  wait=False returns at once; c is valid after
    transfer_sync()."""
    for x in (a, b) + ((c,) if c is not None else ()):
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.ndim == 2 and x.flags.f_contiguous):
            raise ValueError("times_host: expected column-major float64 matrices")
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"flomat*: expected two matrices with compatible dimensions, got {a.shape[0]}x{a.shape[1]} and {b.shape[0]}x{b.shape[1]}")
    m, k, n = a.shape[0], a.shape[1], b.shape[1]
    if c is None:
        c, beta = np.empty((m, n), order="F"), 0.0
    elif c.shape != (m, n):
        raise ValueError("flomat*!: result matrix has the wrong dimensions")
    check(_L().fm_dgemm_host(m, n, k, float(alpha), a.ctypes.data, max(1, m), b.ctypes.data, max(1, k), float(beta), c.ctypes.data,
                             max(1, m), 0 if wait else 1), "fm_dgemm_host")
    return c


def times(a, *bs):
    """`times` (flomat.rkt:3202): left fold; matrix*matrix -> flomat*, anything with a number -> `.*`."""
    for b in bs:
        if isinstance(a, Flomat) and isinstance(b, Flomat):
            a = flomat_times(a, b)
        elif isinstance(a, Flomat) or isinstance(b, Flomat):
            a = dot_times(a, b)
        else:
            a = a * b
    return a


def power(a: Flomat, n: int) -> Flomat:
    """`power` / flomat-expt (flomat.rkt:918-927): same recursion, hence the same product order."""
    _check_square(a, "matrix-expt")
    if n < 0:
        raise ValueError("flomat-expt: expected a natural number")
    if n == 0:
        return eye(a.m)
    if n == 1:
        return copy_flomat(a)
    if n == 2:
        return flomat_times(a, a)
    if n % 2 == 0:
        h = power(a, n // 2)
        return flomat_times(h, h)
    return flomat_times(a, power(a, n - 1))


# --------------------------------------------------------------------------------------------------
# norm / LU family
# --------------------------------------------------------------------------------------------------
def norm(a: Flomat, kind="frob") -> float:
    """`norm` (flomat.rkt:3273 -> flomat-norm :1328 -> dlange_): kind in 1, 'inf', 'frob', 'max'."""
    code = {1: b"1", "1": b"1", "inf": b"I", "frob": b"F", "max": b"M"}.get(kind)
    if code is None:
        raise ValueError("flomat-norm: unknown norm type")
    out = c_double()
    check(_L().fm_norm(code, a.m, a.n, a.a, a.lda, byref(out)), "fm_norm")
    return out.value


class Pivots:
    """`(struct pivots (ps))` (flomat.rkt:1457): int32, 1-based, device resident; plus the LAPACK `info` word."""

    def __init__(self, n: int):
        self.n = int(n)
        off = 4 * max(self.n, 1)
        off += (16 - off % 16) % 16
        self._buf = _Buffer(off + 16)
        self.ptr = self._buf.ptr
        self.info_ptr = self.ptr + off

    def to_numpy(self) -> np.ndarray:
        h = np.zeros(self.n, dtype=np.int32)
        if self.n:
            check(_L().fm_download_i32(h.ctypes.data, self.ptr, self.n), "fm_download_i32")
        return h

    def info(self) -> int:
        h = np.zeros(1, dtype=np.int32)
        check(_L().fm_download_i32(h.ctypes.data, self.info_ptr, 1), "fm_download_i32")
        return int(h[0])

    def sign(self) -> int:
        """pivots-sign (flomat.rkt:1487)."""
        p = self.to_numpy()
        return -1 if int(np.count_nonzero(p - 1 != np.arange(self.n))) % 2 else 1


def _pivots(n: int) -> Pivots:
    return Pivots(n)


def flomat_lu_bang(a: Flomat):
    """flomat-lu! (flomat.rkt:1525): in-place dgetrf; returns (pivots, info)."""
    piv = _pivots(min(a.m, a.n))
    check(_L().fm_getrf(a.m, a.n, a.a, a.lda, piv.ptr, piv.info_ptr), "fm_getrf")
    return piv, piv.info()


def flomat_lu_nopivot_bang(a: Flomat):
    """LU without interchanges, verified (fm_getrf_nopiv): returns (pivots = 1..n, info, violated).  `violated` is True when partial
    pivoting would have interchanged rows; the factors are then NOT dgetrf's and flomat_lu_bang must be used on the original matrix.
    No reference counterpart: it is the fast path of the solve inside `expm`, whose denominator is column diagonally dominant."""
    _check_square(a, "flomat-lu (no pivoting)")
    piv = Pivots(a.m + 4)  # (the violation word sits behind the pivots)
    viol_ptr = piv.ptr + 4 * a.m
    check(_L().fm_getrf_nopiv(a.m, a.a, a.lda, piv.ptr, piv.info_ptr, viol_ptr), "fm_getrf_nopiv")
    h = np.zeros(1, dtype=np.int32)
    check(_L().fm_download_i32(h.ctypes.data, viol_ptr, 1), "fm_download_i32")
    piv.n = a.m
    return piv, piv.info(), bool(h[0])


def mldivide(a: Flomat, b: Flomat) -> Flomat:
    """`mldivide` (flomat.rkt:3289) -> flomat-solve-many :2145: copies A and B, dgesv, info ignored; returns X."""
    _check_square(a, "mldivide")
    if b.m != a.m:
        raise ValueError("mldivide: A and B must have the same number of rows")
    lu, x = copy_flomat(a), copy_flomat(b)
    piv = _pivots(a.m)
    check(_L().fm_gesv(a.m, x.n, lu.a, lu.lda, piv.ptr, x.a, x.lda, piv.info_ptr), "fm_gesv")
    return x


def result_flcolumn(b) -> Flomat:
    """result-flcolumn (flomat.rkt:1896): a vector or a list becomes an m x 1 matrix, a matrix passes through unchanged."""
    if isinstance(b, Flomat):
        return b
    return Flomat.from_numpy(np.asarray(b, dtype=np.float64).reshape(-1, 1))


def flomat_augment(*cols: Flomat) -> Flomat:
    """flomat-augment (flomat.rkt:1223): [C1 C2 ...] side by side, a new matrix (one device copy per piece)."""
    if not cols:
        raise ValueError("flomat-augment: expected at least one matrix")
    m = cols[0].m
    if any(c.m != m for c in cols):
        raise ValueError("flomat-augment: all arguments must have same number of rows")
    out = Flomat.alloc(m, sum(c.n for c in cols))
    j = 0
    for c in cols:
        v = out.sub(0, j, m, c.n)
        check(_L().fm_copy2d(v.a, v.lda, c.a, c.lda, m, c.n), "fm_copy2d")
        j += c.n
    return out


def flomat_solve_many_bang(a: Flomat, b: Flomat) -> Flomat:
    """flomat-solve-many! (flomat.rkt:2135): dgesv in place, A and B overwritten, info ignored; returns B."""
    _check_square(a, "flomat-solve-many!")
    if b.m != a.m:
        raise ValueError("flomat-solve-many!: A and B must have the same number of rows")
    piv = _pivots(a.m)
    check(_L().fm_gesv(a.m, b.n, a.a, a.lda, piv.ptr, b.a, b.lda, piv.info_ptr), "fm_gesv")
    return b


def flomat_solve_many(a: Flomat, bs_or_b) -> Flomat:
    """flomat-solve-many (flomat.rkt:2145): B is a matrix (copied) or a LIST of columns (augmented); A is copied."""
    if isinstance(bs_or_b, (list, tuple)):
        b = flomat_augment(*[result_flcolumn(c) for c in bs_or_b])
    else:
        b = copy_flomat(bs_or_b)
    return flomat_solve_many_bang(copy_flomat(a), b)


flomat_left_divide = flomat_solve_many  # flomat.rkt:2159


def flomat_solve(a: Flomat, b) -> Flomat:
    """flomat-solve (flomat.rkt:2126): one right-hand side column; copies A and b, dgesv with nrhs = 1, info ignored."""
    return flomat_solve_many_bang(copy_flomat(a), copy_flomat(result_flcolumn(b)))


def inv(a: Flomat) -> Flomat:
    """`inv` (flomat.rkt:3297) -> flomat-inverse! :1770: copy, dgetrf, dgetri."""
    _check_square(a, "flomat-inverse!")
    lu = copy_flomat(a)
    piv = _pivots(a.m)
    check(_L().fm_getrf(a.m, a.n, lu.a, lu.lda, piv.ptr, piv.info_ptr), "fm_getrf")
    check(_L().fm_getri(a.m, lu.a, lu.lda, piv.ptr, None), "fm_getri")
    return lu


def det(a: Flomat) -> float:
    """`det` (flomat.rkt:3276) -> :1844 -> :1837: copy, dgetrf, sign(ipiv) * left-to-right diagonal product."""
    _check_square(a, "matrix-determinant")
    lu = copy_flomat(a)
    piv = _pivots(a.m)
    check(_L().fm_getrf(a.m, a.n, lu.a, lu.lda, piv.ptr, piv.info_ptr), "fm_getrf")
    out = c_double()
    check(_L().fm_det(a.m, lu.a, lu.lda, piv.ptr, byref(out)), "fm_det")
    return out.value


def flomat_inverse_bang(a: Flomat) -> Flomat:
    """flomat-inverse! (flomat.rkt:1770-1779): A := A^-1 in place (dgetrf_, dgetri_ with the copy as workspace in the reference; the
    device path needs no workspace).  `info` is dropped, as the reference drops it."""
    _check_square(a, "flomat-inverse!")
    piv = _pivots(a.m)
    check(_L().fm_getrf(a.m, a.n, a.a, a.lda, piv.ptr, piv.info_ptr), "fm_getrf")
    check(_L().fm_getri(a.m, a.a, a.lda, piv.ptr, None), "fm_getri")
    return a


def flomat_inverse(a: Flomat) -> Flomat:
    """flomat-inverse (flomat.rkt:1781): (flomat-inverse! (copy-flomat A))."""
    return flomat_inverse_bang(copy_flomat(a))


def pivots_sign(piv: Pivots) -> int:
    """pivots-sign (flomat.rkt:1487): (-1)^(number of rows i with ipiv[i] != i)."""
    return piv.sign()


def flomat_determinant_from_plu(lu: Flomat, piv: Pivots) -> float:
    """flomat-determinant-from-plu (flomat.rkt:1837-1842): pivots-sign times the left-to-right product of the diagonal of L\\U."""
    _check_square(lu, "flomat-determinant-from-plu")
    out = c_double()
    check(_L().fm_det(lu.m, lu.a, lu.lda, piv.ptr, byref(out)), "fm_det")
    return out.value


flomat_determinant = det      # flomat.rkt:1844
flomat_zeros = zeros          # flomat.rkt:548
flomat_ones = ones            # flomat.rkt:551
flomat_eye = eye              # flomat.rkt:2642
flomat_expt = power           # flomat.rkt:918


def flomat_identity(m: int) -> Flomat:
    """flomat-identity (flomat.rkt:571)."""
    return eye(m)


def flomat_norm(a: Flomat, kind="frob") -> float:
    """flomat-norm (flomat.rkt:1328; default 'frob) and its four named forms (:1343-1352)."""
    return norm(a, kind)


def flomat_norm1(a: Flomat) -> float:
    return norm(a, 1)


def flomat_norm_inf(a: Flomat) -> float:
    return norm(a, "inf")


def flomat_norm_frob(a: Flomat) -> float:
    return norm(a, "frob")


def flomat_norm_max(a: Flomat) -> float:
    return norm(a, "max")


def flomat_times_bang(a: Flomat, b: Flomat, c: Flomat, alpha: float = 1.0, beta: float = 1.0, transpose_a: bool = False,
                      transpose_b: bool = False) -> Flomat:
    """flomat*! (flomat.rkt:898-903): C := alpha*op(A)*op(B) + beta*C into the caller's C."""
    return flomat_times(a, b, c, alpha, beta, transpose_a, transpose_b)


def find_scaling_exponent(a: Flomat) -> float:
    """find-scaling-exponent (expm.rkt:212-215): s = 1 + ceiling(log(||A||_1) / log 2), as a float (it may be <= 0, -inf for the zero matrix)."""
    nrm = norm(a, 1)
    if nrm == 0.0:
        return -math.inf
    if math.isnan(nrm) or math.isinf(nrm):
        return nrm
    return 1.0 + math.ceil(math.log(nrm) / math.log(2.0))


def repeated_squaring(a: Flomat, s) -> Flomat:
    """repeated-squaring (expm.rkt:209-210): A^(2^s) by s squarings; s <= 0 returns A itself."""
    while s > 0:
        a = flomat_times(a, a)
        s -= 1
    return a


# --------------------------------------------------------------------------------------------------
# neighbours of the hot path kept on the device
# --------------------------------------------------------------------------------------------------
def transpose(a: Flomat) -> Flomat:
    """`transpose` / flomat-transpose (flomat.rkt:1360): a fresh n x m matrix."""
    out = Flomat.alloc(a.n, a.m)
    check(_L().fm_transpose(a.m, a.n, a.a, a.lda, out.a, out.lda), "fm_transpose")
    return out


def mrdivide(b: Flomat, a: Flomat) -> Flomat:
    """`mrdivide` (flomat.rkt:3293): B/A = (A'\\B')'."""
    return transpose(mldivide(transpose(a), transpose(b)))


def flomat_extract_upper(a: Flomat) -> Flomat:
    """flomat-extract-upper (flomat.rkt:1540): the k x k upper triangle (k = min(m, n)), diagonal included."""
    k = min(a.m, a.n)
    out = Flomat.alloc(k, k)
    check(_L().fm_tri(0, 0, k, k, a.a, a.lda, out.a, out.lda), "fm_tri")
    return out


def flomat_extract_lower(a: Flomat) -> Flomat:
    """flomat-extract-lower (flomat.rkt:1566): copy with everything above the diagonal zeroed."""
    out = Flomat.alloc(a.m, a.n)
    check(_L().fm_tri(1, 0, a.m, a.n, a.a, a.lda, out.a, out.lda), "fm_tri")
    return out


def flomat_extract_lower_ones(a: Flomat) -> Flomat:
    """flomat-extract-lower/ones (flomat.rkt:1553): strictly lower part with ones on the diagonal."""
    out = Flomat.alloc(a.m, a.n)
    check(_L().fm_tri(1, 1, a.m, a.n, a.a, a.lda, out.a, out.lda), "fm_tri")
    return out


def pivots_to_flomat(ps: "Pivots") -> Flomat:
    """pivots->flomat (flomat.rkt:1471): the permutation matrix P with A = P L U."""
    out = Flomat.alloc(ps.n, ps.n)
    check(_L().fm_pivots_to_matrix(ps.n, ps.ptr, out.a, out.lda), "fm_pivots_to_matrix")
    return out


def flomat_plu(a: Flomat):
    """flomat-plu (flomat.rkt:1529): copy, dgetrf, then P, L (unit lower), U."""
    b = copy_flomat(a)
    ps, _info = flomat_lu_bang(b)
    return pivots_to_flomat(ps), flomat_extract_lower_ones(b), flomat_extract_upper(b)


def flomat_times_vector(a: Flomat, x: Flomat, y: Flomat | None = None, alpha: float = 1.0, beta: float = 1.0,
                        transpose_a: bool = False) -> Flomat:
    """flomat*vector (flomat.rkt:970): Y := alpha op(A) X + beta Y via dgemv; without Y a zero column is allocated."""
    rows, cols = (a.n, a.m) if transpose_a else (a.m, a.n)
    if x.n != 1 or x.m != cols:
        raise ValueError(f"constant*matrix*vector+constant*vector!: expected a column of {cols} rows, got {x.m}x{x.n}")
    if y is None:
        y, beta = make_flomat(rows, 1, 0.0), 1.0
    elif y.n != 1 or y.m != rows:
        raise ValueError(f"constant*matrix*vector+constant*vector!: expected a column of {rows} rows, got {y.m}x{y.n}")
    check(_L().fm_gemv(TRANS if transpose_a else NO_TRANS, a.m, a.n, float(alpha), a.a, a.lda, x.a, 1, float(beta), y.a, 1), "fm_gemv")
    return y


def dot(x: Flomat, y: Flomat) -> float:
    """`dot` / flcolumn-dot (flomat.rkt:3216, :1926): ddot of two m x 1 columns."""
    if x.n != 1 or y.n != 1 or x.m != y.m:
        raise ValueError(f"column-dot: expected two mx1 matrices with same number of rows, got {x.m}x{x.n} and {y.m}x{y.n}")
    out = c_double()
    check(_L().fm_dot(x.m, x.a, 1, y.a, 1, byref(out)), "fm_dot")
    return out.value


def flcolumn_norm(v: Flomat) -> float:
    """flcolumn-norm (flomat.rkt:1941): dnrm2 of the first column."""
    out = c_double()
    check(_L().fm_nrm2(v.m, v.a, 1, byref(out)), "fm_nrm2")
    return out.value


def flomat_max_abs_index(a: Flomat):
    """flomat-max-abs-index (flomat.rkt:1163): (i, j) of the first entry of largest magnitude (column-major order)."""
    idx = ctypes.c_int()
    if a.lda == a.m or a.n == 1:
        check(_L().fm_iamax(a.m * a.n, a.a, 1, byref(idx)), "fm_iamax")
        return idx.value % max(a.m, 1), idx.value // max(a.m, 1)
    best = None
    for j in range(a.n):  # a view: one idamax per column, as the reference does (:1173-1178)
        check(_L().fm_iamax(a.m, a.a + 8 * j * a.lda, 1, byref(idx)), "fm_iamax")
        v = abs(a.ref(idx.value, j))
        if best is None or v > best[0]:
            best = (v, idx.value, j)
    return best[1], best[2]


def flomat_max_abs_value(a: Flomat) -> float:
    i, j = flomat_max_abs_index(a)
    return a.ref(i, j)


def flomat_swap_rows_bang(a: Flomat, i1: int, i2: int) -> Flomat:
    """flomat-swap-rows! (flomat.rkt:1081): dswap with stride lda."""
    if not (0 <= i1 < a.m and 0 <= i2 < a.m):
        raise IndexError("flomat-swap-rows!: row index out of range")
    if i1 != i2:
        check(_L().fm_swap(a.n, a.a + 8 * i1, a.lda, a.a + 8 * i2, a.lda), "fm_swap")
    return a


def flomat_swap_columns_bang(a: Flomat, j1: int, j2: int) -> Flomat:
    """flomat-swap-columns! (flomat.rkt:1096)."""
    if not (0 <= j1 < a.n and 0 <= j2 < a.n):
        raise IndexError("flomat-swap-columns!: column index out of range")
    if j1 != j2:
        check(_L().fm_swap(a.m, a.a + 8 * j1 * a.lda, 1, a.a + 8 * j2 * a.lda, 1), "fm_swap")
    return a


def colsums(a: Flomat) -> Flomat:
    """`colsums` / flomat-column-sums (flomat.rkt:2094): 1 x n row of column sums (n ddots against 1.0 in the reference)."""
    out = Flomat.alloc(1, a.n)
    check(_L().fm_colsums(a.m, a.n, a.a, a.lda, out.a), "fm_colsums")
    return out


def rowsums(a: Flomat) -> Flomat:
    """`rowsums` / flomat-row-sums (flomat.rkt:2101): m x 1 column of row sums."""
    out = Flomat.alloc(a.m, 1)
    check(_L().fm_rowsums(a.m, a.n, a.a, a.lda, out.a), "fm_rowsums")
    return out


def flomat_cholesky_bang(a: Flomat, upper: bool = False) -> int:
    """flomat-cholesky! (flomat.rkt:1815): dpotrf in place, returns LAPACK's info (0, or the order of the first leading minor
    that is not positive definite).  Only the selected triangle of A is overwritten."""
    _check_square(a, "flomat-cholesky!")
    info = Pivots(0)
    check(_L().fm_potrf(b"U" if upper else b"L", a.m, a.a, a.lda, info.info_ptr), "fm_potrf")
    return info.info()


def flomat_cholesky(a: Flomat, upper: bool = False) -> Flomat:
    """flomat-cholesky (flomat.rkt:1821): copy, factor, extract the requested triangle."""
    b = copy_flomat(a)
    flomat_cholesky_bang(b, upper)
    return flomat_extract_upper(b) if upper else flomat_extract_lower(b)


def _map(fn: int, a: Flomat, c: Flomat | None, who: str) -> Flomat:
    if c is None:
        c = Flomat.alloc(a.m, a.n)
    else:
        _check_same(a, c, who)
    check(_L().fm_map(fn, a.m, a.n, a.a, a.lda, c.a, c.lda), who)
    return c


def dot_sin(a: Flomat, c: Flomat | None = None) -> Flomat:
    """`.sin` (`define-pointwise-unaries`, flomat.rkt:3019); pass c=a for `.sin!`."""
    return _map(MAP_SIN, a, c, ".sin")


def dot_cos(a: Flomat, c: Flomat | None = None) -> Flomat:
    return _map(MAP_COS, a, c, ".cos")


def dot_tan(a: Flomat, c: Flomat | None = None) -> Flomat:
    return _map(MAP_TAN, a, c, ".tan")


def dot_sqr(a: Flomat, c: Flomat | None = None) -> Flomat:
    return _map(MAP_SQR, a, c, ".sqr")


def dot_sqrt(a: Flomat, c: Flomat | None = None) -> Flomat:
    return _map(MAP_SQRT, a, c, ".sqrt")


def dot_log(a: Flomat, c: Flomat | None = None) -> Flomat:
    return _map(MAP_LOG, a, c, ".log")


def dot_exp(a: Flomat, c: Flomat | None = None) -> Flomat:
    return _map(MAP_EXP, a, c, ".exp")


def dot_expt(a, b, c: Flomat | None = None) -> Flomat:
    """`.expt` (`define-pointwise-binaries`, flomat.rkt:3076): matrix^matrix, matrix^number or number^matrix."""
    return _ewise(OP_EXPT, a, b, c, ".expt")


the_iseed = np.array([1, 2, 3, 5], dtype=np.int32)  # the package-wide LAPACK seed (the reference keeps one `the-iseed`)


def _larnv(idist: int, m: int, n: int, iseed) -> Flomat:
    seed = the_iseed if iseed is None else iseed
    if seed.dtype != np.int32 or seed.shape != (4,):
        raise ValueError("iseed: expected an int32 array of 4 entries")
    out = Flomat.alloc(m, n)
    check(_L().fm_larnv(idist, seed.ctypes.data, m * n, out.a), "fm_larnv")
    return out


def rand(m: int, n: int | None = None, iseed=None) -> Flomat:
    """`rand` (flomat.rkt:2814): uniform (0,1) from LAPACK's dlarnv generator, generated on the device."""
    return _larnv(1, m, m if n is None else n, iseed)


def randn(m: int, n: int | None = None, iseed=None) -> Flomat:
    """`randn` (flomat.rkt:2835): standard normal (dlarnv idist = 3)."""
    return _larnv(3, m, m if n is None else n, iseed)


# --------------------------------------------------------------------------------------------------
# expm (expm.rkt:203)
# --------------------------------------------------------------------------------------------------
EXPM_REFERENCE_ORDER = 0   # the reference's Horner order (7+s products)
EXPM_MIN_PRODUCTS = 1      # 5+s products
EXPM_CLAMP_S = 2           # do not reproduce the reference's s<=0 quirk


def expm(a: Flomat, mode: int = EXPM_REFERENCE_ORDER, return_s: bool = False):
    """`expm` (expm.rkt:203-207): Pade [8/8] scaling and squaring, fully on device."""
    _check_square(a, "expm")
    out = Flomat.alloc(a.m, a.n)
    s = c_double()
    info = c_int32()
    check(_L().fm_expm(a.m, a.a, a.lda, out.a, out.lda, int(mode), byref(s), byref(info)), "fm_expm")
    return (out, s.value) if return_s else out


def expm_to_host(a: Flomat, out: np.ndarray, mode: int = EXPM_REFERENCE_ORDER, return_s: bool = False):
    """`expm` (expm.rkt:203) with the result delivered into the column-major float64 host array `out`: the last squaring runs in
    column blocks whose downloads overlap the blocks still to be computed.  A pinned `out` is valid after transfer_sync()."""
    _check_square(a, "expm")
    if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (a.m, a.n) and out.flags.f_contiguous):
        raise ValueError("expm_to_host: expected a column-major float64 array of the matrix's shape")
    s = c_double()
    info = c_int32()
    check(_L().fm_expm(a.m, a.a, a.lda, out.ctypes.data, max(1, a.m), int(mode), byref(s), byref(info)), "fm_expm")
    return (out, s.value) if return_s else out


def expm_batched(a_ptr: int, n: int, batch: int, out_ptr: int, mode: int = EXPM_REFERENCE_ORDER, s_out: np.ndarray | None = None):
    """Batch of `batch` dense n x n matrices stored back to back (stride n*n) at device address a_ptr."""
    sp = s_out.ctypes.data if s_out is not None else None
    check(_L().fm_expm_batched(n, a_ptr, n, n * n, out_ptr, n, n * n, batch, int(mode), sp), "fm_expm_batched")


def sync() -> None:
    check(_L().fm_sync(), "fm_sync")


def transfer_sync() -> None:
    """Wait for the asynchronous uploads / downloads issued so far (Flomat.upload_async / download_async)."""
    check(_L().fm_transfer_sync(), "fm_transfer_sync")


# --------------------------------------------------------------------------------------------------
# all devices of the box from ONE process and ONE calling thread (include/flomat_cuda.h, tier 3b): the form `flomat*` can use
# --------------------------------------------------------------------------------------------------
def mg_init(ndev: int = 0) -> int:
    """Start the device group (one worker thread per device, peer access between all pairs); returns the number of devices."""
    check(_L().fm_mg_init(int(ndev)), "fm_mg_init")
    return _L().fm_mg_device_count()


def mg_sync() -> None:
    check(_L().fm_mg_sync(), "fm_mg_sync")


def times_mg(a: Flomat, b: Flomat, c: Flomat | None = None, alpha: float = 1.0, beta: float = 1.0) -> Flomat:
    """flomat* (flomat.rkt:905) on device-resident matrices, sharded over every device of the group by column blocks of B and C:
    the other devices fetch A and their block of B over NVLink and write their tiles of C straight into this device's C."""
    if a.n != b.m:
        raise ValueError(f"flomat*: expected two matrices with compatible dimensions, got {a.m}x{a.n} and {b.m}x{b.n}")
    if c is None:
        c, beta = Flomat.alloc(a.m, b.n), 0.0
    elif c.m != a.m or c.n != b.n:
        raise ValueError("flomat*!: result matrix has the wrong dimensions")
    check(_L().fm_mg_dgemm(NO_TRANS, NO_TRANS, a.m, b.n, a.n, float(alpha), a.a, a.lda, b.a, b.lda, float(beta), c.a, c.lda), "fm_mg_dgemm")
    return c


def times_host_mg(a: np.ndarray, b: np.ndarray, c: np.ndarray | None = None, alpha: float = 1.0, beta: float = 1.0) -> np.ndarray:
    """flomat* on HOST matrices (what the unmodified reference hands to cblas_dgemm, flomat.rkt:888) over every device of the group:
    each device uploads its slice of A and its column block of B over its own PCIe link, A is all-gathered over NVLink."""
    for x in (a, b) + ((c,) if c is not None else ()):
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.ndim == 2 and x.flags.f_contiguous):
            raise ValueError("times_host_mg: expected column-major float64 matrices")
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"flomat*: expected two matrices with compatible dimensions, got {a.shape[0]}x{a.shape[1]} and {b.shape[0]}x{b.shape[1]}")
    m, k, n = a.shape[0], a.shape[1], b.shape[1]
    if c is None:
        c, beta = np.empty((m, n), order="F"), 0.0
    elif c.shape != (m, n):
        raise ValueError("flomat*!: result matrix has the wrong dimensions")
    check(_L().fm_mg_dgemm(NO_TRANS, NO_TRANS, m, n, k, float(alpha), a.ctypes.data, max(1, m), b.ctypes.data, max(1, k), float(beta),
                           c.ctypes.data, max(1, m)), "fm_mg_dgemm")
    return c


def expm_mg(a, mode: int = EXPM_REFERENCE_ORDER, return_s: bool = False):
    """`expm` (expm.rkt:203) of ONE matrix over every device of the group: the products sharded by column blocks with the all-gather
    fused into the GEMM epilogue, the LU replicated, the solves sharded by right-hand sides.  `a`: a Flomat on this device (returns a
    Flomat) or a column-major float64 host array (returns a host array)."""
    s = c_double()
    info = c_int32()
    if isinstance(a, Flomat):
        _check_square(a, "expm")
        out = Flomat.alloc(a.m, a.n)
        check(_L().fm_mg_expm(a.m, a.a, a.lda, out.a, out.lda, int(mode), byref(s), byref(info)), "fm_mg_expm")
    else:
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == 2 and a.shape[0] == a.shape[1] and a.flags.f_contiguous):
            raise ValueError("expm_mg: expected a Flomat or a square column-major float64 array")
        n = a.shape[0]
        out = np.empty((n, n), dtype=np.float64, order="F")
        check(_L().fm_mg_expm(n, a.ctypes.data, max(1, n), out.ctypes.data, max(1, n), int(mode), byref(s), byref(info)), "fm_mg_expm")
    return (out, s.value) if return_s else out


def expm_batched_mg(a_ptr: int, n: int, batch: int, out_ptr: int, mode: int = EXPM_REFERENCE_ORDER, s_out: np.ndarray | None = None):
    """`batch` dense n x n matrices stored back to back at a_ptr (host or this device's memory), one contiguous block per device."""
    sp = s_out.ctypes.data if s_out is not None else None
    check(_L().fm_mg_expm_batched(n, a_ptr, n, n * n, out_ptr, n, n * n, batch, int(mode), sp), "fm_mg_expm_batched")
