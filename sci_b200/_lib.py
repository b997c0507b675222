"""ctypes binding of libflomat_cuda.so -- the same C ABI a Racket `ffi-lib` would bind (include/flomat_cuda.h).

Fails loudly: if the shared library has not been built, or no B200 is present when a kernel is needed, the caller gets
an exception -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char, c_char_p, c_double, c_int, c_int32, c_longlong, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOMAT_CUDA_LIB", os.path.join(_HERE, "lib", "libflomat_cuda.so"))

_DP = c_void_p  # device or host double*; passed as raw addresses
_IP = c_void_p


class FlomatCudaError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
SIGNATURES = {
    # tier 1: compat symbols (flomat.rkt binding sites cited in include/flomat_cuda.h)
    "cblas_dgemm": (None, [c_int, c_int, c_int, c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]),
    "cblas_daxpy": (None, [c_int, c_double, _DP, c_int, _DP, c_int]),
    "cblas_dcopy": (None, [c_int, _DP, c_int, _DP, c_int]),
    "cblas_dscal": (None, [c_int, c_double, _DP, c_int]),
    "dlange_": (c_double, [c_char_p, POINTER(c_int), POINTER(c_int), _DP, POINTER(c_int), _DP]),
    "dgetrf_": (None, [POINTER(c_int), POINTER(c_int), _DP, POINTER(c_int), _IP, POINTER(c_int)]),
    "dgesv_": (None, [POINTER(c_int), POINTER(c_int), _DP, POINTER(c_int), _IP, _DP, POINTER(c_int), POINTER(c_int)]),
    "dgetri_": (None, [POINTER(c_int), _DP, POINTER(c_int), _IP, _DP, POINTER(c_int), POINTER(c_int)]),
    "cblas_dgemv": (None, [c_int, c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]),
    "cblas_ddot": (c_double, [c_int, _DP, c_int, _DP, c_int]),
    "cblas_dnrm2": (c_double, [c_int, _DP, c_int]),
    "cblas_idamax": (c_int, [c_int, _DP, c_int]),
    "cblas_dswap": (None, [c_int, _DP, c_int, _DP, c_int]),
    "dlarnv_": (None, [POINTER(c_int), _IP, POINTER(c_int), _DP]),
    "dpotrf_": (None, [c_char_p, POINTER(c_int), _DP, POINTER(c_int), POINTER(c_int)]),
    # tier 2
    "fm_init": (c_int, [c_int]),
    "fm_device_count": (c_int, []),
    "fm_set_stream": (c_int, [c_void_p]),
    "fm_sync": (c_int, []),
    "fm_last_error": (c_char_p, []),
    "fm_version": (c_char_p, []),
    "fm_kernel_launches": (c_uint64, []),
    "fm_alloc": (c_void_p, [c_int, c_int]),
    "fm_alloc_bytes": (c_void_p, [c_size_t]),
    "fm_free": (c_int, [c_void_p]),
    "fm_trim_cache": (c_int, []),
    "fm_upload": (c_int, [_DP, c_int, _DP, c_int, c_int, c_int]),
    "fm_download": (c_int, [_DP, c_int, _DP, c_int, c_int, c_int]),
    "fm_upload_async": (c_int, [_DP, c_int, _DP, c_int, c_int, c_int]),
    "fm_download_async": (c_int, [_DP, c_int, _DP, c_int, c_int, c_int]),
    "fm_transfer_sync": (c_int, []),
    "fm_get": (c_int, [_DP, c_size_t, POINTER(c_double)]),
    "fm_set": (c_int, [_DP, c_size_t, c_double]),
    "fm_upload_i32": (c_int, [_IP, _IP, c_size_t]),
    "fm_download_i32": (c_int, [_IP, _IP, c_size_t]),
    "fm_fill": (c_int, [_DP, c_int, c_int, c_int, c_double]),
    "fm_eye": (c_int, [_DP, c_int, c_int, c_int, c_int]),
    "fm_copy2d": (c_int, [_DP, c_int, _DP, c_int, c_int, c_int]),
    "fm_ewise": (c_int, [c_int, c_int, c_int, _DP, c_int, c_double, _DP, c_int, c_double, _DP, c_int]),
    "fm_axpy2d": (c_int, [c_int, c_int, c_double, _DP, c_int, _DP, c_int]),
    "fm_scal2d": (c_int, [c_int, c_int, c_double, _DP, c_int]),
    "fm_addsub": (c_int, [c_int, c_int, _DP, c_int, _DP, c_int, _DP, c_int, _DP, c_int]),
    "fm_norm": (c_int, [c_char, c_int, c_int, _DP, c_int, POINTER(c_double)]),
    "fm_transpose": (c_int, [c_int, c_int, _DP, c_int, _DP, c_int]),
    "fm_tri": (c_int, [c_int, c_int, c_int, c_int, _DP, c_int, _DP, c_int]),
    "fm_pivots_to_matrix": (c_int, [c_int, _IP, _DP, c_int]),
    "fm_map": (c_int, [c_int, c_int, c_int, _DP, c_int, _DP, c_int]),
    "fm_gemv": (c_int, [c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]),
    "fm_dot": (c_int, [c_int, _DP, c_int, _DP, c_int, POINTER(c_double)]),
    "fm_nrm2": (c_int, [c_int, _DP, c_int, POINTER(c_double)]),
    "fm_iamax": (c_int, [c_int, _DP, c_int, POINTER(c_int)]),
    "fm_swap": (c_int, [c_int, _DP, c_int, _DP, c_int]),
    "fm_colsums": (c_int, [c_int, c_int, _DP, c_int, _DP]),
    "fm_rowsums": (c_int, [c_int, c_int, _DP, c_int, _DP]),
    "fm_larnv": (c_int, [c_int, _IP, c_longlong, _DP]),
    "fm_potrf": (c_int, [c_char, c_int, _DP, c_int, _IP]),
    "fm_dgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]),
    "fm_dgemm_host": (c_int, [c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int, c_int]),
    "fm_dgemm_host_plan": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "fm_dgemm_tile_plan": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "fm_lu_host_plan": (c_int, [c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    "fm_dgemm_diag": (c_int, [c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int, c_double]),
    "fm_dgemm_batched": (c_int, [c_int, c_int, c_int, c_double, _DP, c_int, c_longlong, _DP, c_int, c_longlong, c_double, _DP, c_int,
                                 c_longlong, c_double, c_int]),
    "fm_getrf": (c_int, [c_int, c_int, _DP, c_int, _IP, _IP]),
    "fm_getrf_nopiv": (c_int, [c_int, _DP, c_int, _IP, _IP, _IP]),
    "fm_getrf_chunked": (c_int, [c_int, _DP, c_int, _IP, _IP, c_int]),
    "fm_getrs": (c_int, [c_int, c_int, _DP, c_int, _IP, _DP, c_int]),
    "fm_gesv": (c_int, [c_int, c_int, _DP, c_int, _IP, _DP, c_int, _IP]),
    "fm_getri": (c_int, [c_int, _DP, c_int, _IP, _IP]),
    "fm_det": (c_int, [c_int, _DP, c_int, _IP, POINTER(c_double)]),
    "fm_expm": (c_int, [c_int, _DP, c_int, _DP, c_int, c_int, POINTER(c_double), POINTER(c_int32)]),
    "fm_expm_batched": (c_int, [c_int, _DP, c_int, c_longlong, _DP, c_int, c_longlong, c_int, c_int, c_void_p]),
    # tier 3
    "fm_ipc_export": (c_int, [c_void_p, c_void_p, POINTER(c_size_t)]),
    "fm_ipc_open": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "fm_ipc_close": (c_int, [c_void_p]),
    "fm_dgemm_colshard": (c_int, [c_int, c_int, c_int, _DP, c_int, _DP, c_int, _DP, c_int, c_int, c_void_p, c_int]),
    # tier 3b: one process, one calling thread, all devices
    "fm_mg_init": (c_int, [c_int]),
    "fm_mg_device_count": (c_int, []),
    "fm_mg_dgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, _DP, c_int, _DP, c_int, c_double, _DP, c_int]),
    "fm_mg_expm_batched": (c_int, [c_int, _DP, c_int, c_longlong, _DP, c_int, c_longlong, c_int, c_int, c_void_p]),
    "fm_mg_expm": (c_int, [c_int, _DP, c_int, _DP, c_int, c_int, POINTER(c_double), POINTER(c_int32)]),
    "fm_mg_sync": (c_int, []),
    "fm_mg_column_block": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "fm_mg_item_block": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
}


def load():
    """dlopen the library and declare every prototype.  Raises if the .so is missing (build it with
    `python sci_b200/build.py` or `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlomatCudaError(
            f"{LIB_PATH} not found: build it with `python sci_b200/build.py`. There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the ABI and the header diverged
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().fm_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise FlomatCudaError(f"{what or 'libflomat_cuda'} failed with code {rc}: {last_error()}")


def init(device: int | None = None) -> None:
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("FLOMAT_CUDA_DEVICE", "0")))
    check(lib.fm_init(device), "fm_init")
