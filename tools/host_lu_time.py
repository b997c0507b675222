"""Wall-clock time of dgesv_ / dgetrf_ with PAGEABLE host operands (the reference's call, flomat.rkt:2140 / :1510), n = 8192, 64 rhs;
FLOMAT_CUDA_NO_LU_PIPELINE=1 gives the upload -> factor -> download sequence of round 1, FLOMAT_CUDA_LU_PIPELINE_CHUNK the chunk width."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
from ctypes import byref, c_int
import numpy as np, sci_b200 as fm
fm.init(0); lib = fm.load()
out = {"pipeline": os.environ.get("FLOMAT_CUDA_NO_LU_PIPELINE") is None, "chunk": os.environ.get("FLOMAT_CUDA_LU_PIPELINE_CHUNK", "default")}
for n, nrhs in ((4096, 64), (8192, 64), (16384, 64)):
    rng = np.random.default_rng(n)
    a = np.asfortranarray(rng.standard_normal((n, n))); b = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ipiv = np.zeros(n, dtype=np.int32); info = c_int(-99)
    best_sv = best_rf = 1e30
    for rep in range(4):
        A, B = a.copy(order="F"), b.copy(order="F")
        t0 = time.perf_counter()
        lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
        t = time.perf_counter() - t0
        if rep: best_sv = min(best_sv, t)
    resid = float(np.max(np.abs(a[:256, :] @ B - b[:256, :])))
    for rep in range(3):
        A = a.copy(order="F")
        t0 = time.perf_counter()
        lib.dgetrf_(byref(c_int(n)), byref(c_int(n)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, byref(info))
        t = time.perf_counter() - t0
        if rep: best_rf = min(best_rf, t)
    out[f"n{n}"] = {"dgesv_ms": round(best_sv * 1e3, 2), "dgetrf_ms": round(best_rf * 1e3, 2), "info": info.value, "max_residual_rows_0_255": resid,
                    "tflops_dgesv": round(((2 / 3) * n ** 3 + 2.0 * n * n * nrhs) / best_sv / 1e12, 2)}
print(json.dumps(out))
