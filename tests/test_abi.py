"""CPU: the C-ABI library builds/loads without a GPU and exports every symbol include/flomat_cuda.h declares;
host-side argument checks of the Python mirror raise before any device work; no product module imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))


def header_functions():
    text = open(os.path.join(ROOT, "include", "flomat_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|double|uint64_t|char)\s*\*?\s*\*?\s*(\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from sci_b200 import _lib

    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 45, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/flomat_cuda.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_compat_symbols_have_reference_names():
    lib = ctypes.CDLL(os.path.join(ROOT, "sci_b200", "lib", "libflomat_cuda.so"))
    for n in ("cblas_dgemm", "cblas_daxpy", "cblas_dcopy", "cblas_dscal", "dlange_", "dgetrf_", "dgesv_", "dgetri_"):
        getattr(lib, n)  # the names flomat.rkt binds (flomat.rkt:855, 751, 483, 1014, 1317, 1505, 2112, 1754)


def test_no_gpu_means_loud_failure_not_fallback():
    import sci_b200
    from sci_b200 import _lib

    lib = _lib.load()
    if lib.fm_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sci_b200.FlomatCudaError):
        sci_b200.zeros(4)
    assert b"no CPU fallback" in lib.fm_last_error()


def test_host_side_checks_raise_before_device_work():
    from sci_b200.flomat import Flomat, flomat_plus, flomat_times, inv, power

    a = Flomat(2, 3, 0, 2, None)
    b = Flomat(2, 3, 0, 2, None)
    with pytest.raises(ValueError):  # check-product-dimensions (flomat.rkt:340)
        flomat_times(a, b)
    with pytest.raises(ValueError):  # check-same-dimensions (flomat.rkt:326)
        flomat_plus(a, Flomat(3, 2, 0, 3, None))
    with pytest.raises(ValueError):  # check-square (flomat.rkt:388)
        inv(a)
    with pytest.raises(ValueError):
        power(a, 2)
    with pytest.raises(IndexError):
        a.sub(1, 1, 2, 3)
    v = Flomat(4, 4, 1024, 4, None).sub(1, 2, 2, 2)  # shared-submatrix!: pointer arithmetic only (flomat.rkt:295)
    assert (v.m, v.n, v.lda, v.a) == (2, 2, 4, 1024 + 8 * (1 + 2 * 4))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sci_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_dgemm_host_plan_covers_the_columns_in_full_waves():
    """fm_dgemm_host_plan (host arithmetic of the host-operand GEMM pipeline, api.cu): blocks are whole 128-column tile columns,
    cover [0, n) exactly with the remainder in the last block, and at the benchmark shapes their tiles fill >= 97 % of their waves."""
    from ctypes import byref, c_int

    from sci_b200 import _lib

    lib = _lib.load()
    for m, n, sms in [(8192, 8192, 148), (16384, 16384, 148), (4096, 4096, 148), (1024, 1024, 148), (700, 1300, 148), (129, 2055, 148),
                      (8192, 64, 148), (1, 1, 148), (8192, 8192, 0), (5000, 100000, 132)]:
        w, nb = c_int(-1), c_int(-1)
        assert lib.fm_dgemm_host_plan(m, n, sms, byref(w), byref(nb)) == 0
        w, nb = w.value, nb.value
        assert w > 0 and w % 128 == 0 and nb >= 1
        last = n - (nb - 1) * w
        assert 0 < last < 2 * w, (m, n, w, nb)                # blocks 0 .. nb-2 are w wide, the last takes what is left
        assert nb <= 12 or n > 12 * w, (m, n, w, nb)
        if n >= 1024:
            assert nb >= 2, (m, n, w, nb)                      # something to overlap
        if (m, n) in ((8192, 8192), (16384, 16384)):
            tiles = (w // 128) * ((m + 127) // 128)
            s = sms or 148
            waves = -(-tiles // s)
            assert tiles / (waves * s) >= 0.97, (m, n, w, tiles)
    assert lib.fm_dgemm_host_plan(0, 5, 148, byref(c_int()), byref(c_int())) != 0


def test_host_operand_entry_points_check_arguments_before_device_work():
    """times_host / expm_to_host (host mirror of flomat* and expm on host matrices): layout and dimension errors are raised on the
    host, with the reference's messages where it has one (check-product-dimensions, flomat.rkt:340)."""
    import numpy as np

    from sci_b200.flomat import Flomat, expm_to_host, times_host

    a, b = np.zeros((4, 3), order="F"), np.zeros((4, 5), order="F")
    with pytest.raises(ValueError, match="compatible dimensions"):
        times_host(a, b)
    with pytest.raises(ValueError, match="column-major"):
        times_host(np.zeros((4, 3)), np.zeros((3, 5)))          # C-ordered
    with pytest.raises(ValueError, match="column-major"):
        times_host(np.zeros((4, 3), order="F", dtype=np.float32), np.zeros((3, 5), order="F"))
    with pytest.raises(ValueError, match="wrong dimensions"):
        times_host(np.zeros((4, 3), order="F"), np.zeros((3, 5), order="F"), np.zeros((4, 4), order="F"))
    sq = Flomat(3, 3, 0, 3, None)
    with pytest.raises(ValueError):
        expm_to_host(Flomat(2, 3, 0, 2, None), np.zeros((2, 3), order="F"))   # check-square
    with pytest.raises(ValueError, match="column-major"):
        expm_to_host(sq, np.zeros((3, 4), order="F"))
    with pytest.raises(ValueError, match="column-major"):
        expm_to_host(sq, np.zeros((3, 3), dtype=np.float32, order="F"))


def test_dgemm_tile_plan_follows_the_wave_model():
    """fm_dgemm_tile_plan (dgemm.cu: gemm_tile_plan, host arithmetic): 64 x 128 tiles (two CTAs per SM) for few tiles, short inner dimensions
    and partial last waves of at most one tile per SM; 128 x 128 tiles otherwise; k slices only for single small products.  The shapes are
    the ones measured in profiles/r02_gemm_wave_shapes.json -- the plan must pick the faster configuration of each."""
    from ctypes import byref, c_int

    from sci_b200 import _lib

    lib = _lib.load()

    def plan(m, n, k, batch=1, sms=148):
        rows, ks = c_int(-1), c_int(-1)
        assert lib.fm_dgemm_tile_plan(m, n, k, batch, sms, byref(rows), byref(ks)) == 0
        return rows.value, ks.value

    faster_with_128 = [(8192, 8192, 8192), (16384, 16384, 16384), (4096, 4096, 4096), (8192, 2048, 8192), (3072, 3072, 3072), (2560, 2560, 2560),
                       (6144, 1536, 6144), (4608, 4608, 4608), (16384, 2048, 16384)]
    faster_with_64 = [(8192, 1024, 8192), (5120, 1024, 5120), (8192, 640, 8192), (2304, 2304, 2304), (4096, 2432, 4096), (4096, 3072, 4096),
                      (10624, 3072, 4096)]
    for shape in faster_with_128:
        assert plan(*shape) == (128, 1), shape
    for shape in faster_with_64:
        assert plan(*shape) == (64, 1), shape
    assert plan(8192, 8192, 128) == (64, 1)        # rank-128 update of the LU: short k
    assert plan(2048, 2048, 2048) == (64, 1)       # 256 tiles of 128 x 128 cannot fill 2 x 148 slots
    assert plan(512, 512, 512) == (64, 8)          # 32 tiles: k cut into slices of >= 64, at most 2 x 148 CTAs
    assert plan(256, 256, 256) == (64, 4)
    assert plan(1024, 1024, 1024) == (64, 2)
    assert plan(256, 256, 256, batch=512) == (64, 1)   # a batch fills the machine without slicing k
    assert plan(512, 512, 128)[1] == 1             # nothing to slice below k = 256
    assert lib.fm_dgemm_tile_plan(0, 1, 1, 1, 148, byref(c_int()), byref(c_int())) != 0


def test_lu_host_plan_chunks():
    """fm_lu_host_plan: the column chunks of the host-operand LU pipeline (api.cu getrf_host_pipeline): they cover [0, n), every inner bound
    is a multiple of the 128-column panel, the first two chunks are small (their upload overlaps nothing), no chunk exceeds n / 4
    rounded down to a panel (the last chunk's rows come down after everything else), and a short remainder rides with the last chunk."""
    from ctypes import byref, c_int

    from sci_b200 import _lib

    lib = _lib.load()
    for n in (2048, 6144, 8192, 8200, 9100, 16384, 20000, 32768, 1, 127, 1500):
        buf = (c_int * 128)()
        cnt = c_int(-1)
        assert lib.fm_lu_host_plan(n, buf, 128, byref(cnt)) == 0
        b = list(buf[: cnt.value + 1])
        assert b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:])), (n, b)
        assert all(x % 128 == 0 for x in b[:-1]), (n, b)
        widths = [y - x for x, y in zip(b, b[1:])]
        cap = max(1024, n // 4 // 128 * 128)
        assert all(w <= cap + 127 for w in widths), (n, widths)
        if n >= 4096:
            assert widths[0] == widths[1] == 1024, (n, widths)
        assert widths[-1] >= min(128, n), (n, widths)
    assert list((lambda buf, cnt: (lib.fm_lu_host_plan(8192, buf, 128, byref(cnt)), buf[: cnt.value + 1])[1])((c_int * 128)(), c_int())) == [0, 1024, 2048, 4096, 6144, 8192]
    assert lib.fm_lu_host_plan(8192, (c_int * 2)(), 2, byref(c_int())) != 0  # does not fit
