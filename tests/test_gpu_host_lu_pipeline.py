"""dgesv_ / dgetrf_ with HOST operands at n >= 6144 (lowered here to 2048 through FLOMAT_CUDA_LU_PIPELINE_MIN_N, chunks of 256 / 512
columns, so that the oracle stays quick) go through the column-chunk pipeline (api.cu getrf_host_pipeline + lu.cu
k_getrf_streamed): A is uploaded chunk by chunk while the arrived chunks are factored (left-looking over the chunks), L\\U comes back
piecewise.  The reference makes exactly these calls with garbage-collected host memory (flomat.rkt:2140 `dgesv_`, :1510 `dgetrf_`).
Checked here against the oracle (LAPACK) on the same bytes: solution at 1x the north-star gate, P A = L U reconstruction, the pivot
vector, singular semantics (info > 0, b untouched), and equality with the unpipelined path up to rounding."""
import os
import subprocess
import sys
from ctypes import byref, c_int

import numpy as np
import pytest

from parity import gate, gate_solve, tol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def small_pipeline(monkeypatch):
    monkeypatch.setenv("FLOMAT_CUDA_LU_PIPELINE_MIN_N", "2048")
    yield


@pytest.fixture(params=["256", "512", None], ids=["chunk256", "chunk512", "default-chunks"])
def chunking(request, monkeypatch):
    if request.param: monkeypatch.setenv("FLOMAT_CUDA_LU_PIPELINE_CHUNK", request.param)
    return request.param


def rnd(seed, m, n):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)))


def call_dgesv(lib, a, b):
    n, nrhs = a.shape[0], b.shape[1]
    A, B = a.copy(order="F"), b.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    info = c_int(-99)
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    return A, B, ipiv, info.value


def unpack_plu(lu, ipiv):
    n = lu.shape[0]
    L = np.tril(lu, -1) + np.eye(n)
    U = np.triu(lu)
    perm = np.arange(n)
    for i, p in enumerate(ipiv):  # LAPACK: row i was interchanged with row ipiv[i] (1-based), in order
        perm[[i, p - 1]] = perm[[p - 1, i]]
    return perm, L, U


@pytest.mark.parametrize("n,nrhs", [(2048, 1), (2304, 3), (3000, 5)])
def test_dgesv_host_pipeline_matches_the_oracle(fm, ref, chunking, n, nrhs):
    lib = fm.load()
    a, b = rnd(7000 + n, n, n), rnd(7001 + n, n, nrhs)
    lu, x, ipiv, info = call_dgesv(lib, a, b)
    assert info == 0
    gate_solve(f"dgesv_(host pipeline) n={n}", a, b, x, ref.mldivide(a, b), n)
    want_lu, want_piv, _ = ref.lu(a)
    # the interchanges: partial pivoting on the same columns; a differing entry needs two candidates equal to rounding
    assert np.mean(ipiv == want_piv) >= 0.999, f"{np.sum(ipiv != want_piv)} of {n} pivots differ from LAPACK's"
    if np.array_equal(ipiv, want_piv):
        gate(f"L\\U(host pipeline) n={n}", lu, want_lu, n)
    perm, L, U = unpack_plu(lu, ipiv)
    resid = np.linalg.norm(a[perm, :] - L @ U, 1) / np.linalg.norm(a, 1)
    assert resid <= tol(n), f"||P A - L U||_1 / ||A||_1 = {resid:.3e} > {tol(n):.3e}"
    assert np.max(np.abs(np.tril(lu, -1))) <= 1.0  # partial pivoting: every multiplier is at most 1


def test_dgetrf_host_pipeline_matches_the_oracle(fm, ref, chunking):
    lib = fm.load()
    n = 2560
    a = rnd(7100, n, n)
    A = a.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    info = c_int(-99)
    lib.dgetrf_(byref(c_int(n)), byref(c_int(n)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, byref(info))
    assert info.value == 0
    want_lu, want_piv, _ = ref.lu(a)
    assert np.array_equal(ipiv, want_piv)
    gate("dgetrf_(host pipeline)", A, want_lu, n)


def test_host_pipeline_with_a_padded_leading_dimension(fm, ref):
    """lda > n on the host side (a view into a larger Racket matrix): only the n x n block is read and written."""
    lib = fm.load()
    n, lda, nrhs = 2176, 2200, 2
    big = np.asfortranarray(np.full((lda, n), 777.0))
    a = rnd(7200, n, n)
    big[:n, :] = a
    b = rnd(7201, n, nrhs)
    B = b.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    info = c_int(-99)
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), big.ctypes.data, byref(c_int(lda)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    assert info.value == 0
    assert np.all(big[n:, :] == 777.0)
    gate_solve("dgesv_(host pipeline, lda > n)", a, b, B, ref.mldivide(a, b), n)
    want_lu, want_piv, _ = ref.lu(a)
    assert np.array_equal(ipiv, want_piv)
    gate("L\\U(host pipeline, lda > n)", big[:n, :], want_lu, n)


def test_host_pipeline_singular_matrix_leaves_b_untouched(fm, ref, chunking):
    """LAPACK dgesv: info = first zero pivot, no solve, b as it was (the reference ignores info and returns the copy of b,
    flomat.rkt:2135-2142).  A zero column stays zero under every update, so its pivot is an exact zero in both implementations."""
    lib = fm.load()
    n = 2304
    a = rnd(7300, n, n)
    a[:, 2100] = 0.0
    b = rnd(7301, n, 2)
    lu, x, ipiv, info = call_dgesv(lib, a, b)
    want_lu, want_piv, want_info = ref.lu(a)
    assert info == want_info == 2101
    assert np.array_equal(x, b)


def test_host_pipeline_equals_the_plain_path_up_to_rounding(fm):
    """the same call with the pipeline switched off (a fresh process: the switch is read once): same pivots, L\\U within the gate"""
    n = 2304
    code = f"""
import sys, numpy as np
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
from ctypes import byref, c_int
import sci_b200 as fm
fm.init(0); lib = fm.load()
a = np.asfortranarray(np.random.default_rng(7400).standard_normal(({n}, {n})))
A = a.copy(order="F"); ipiv = np.zeros({n}, dtype=np.int32); info = c_int(-99)
lib.dgetrf_(byref(c_int({n})), byref(c_int({n})), A.ctypes.data, byref(c_int({n})), ipiv.ctypes.data, byref(info))
np.save(sys.argv[1], A); np.save(sys.argv[2], ipiv)
"""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = {}
        for tag, env in (("pipe", {"FLOMAT_CUDA_LU_PIPELINE_CHUNK": "384"}), ("plain", {"FLOMAT_CUDA_NO_LU_PIPELINE": "1"})):
            fa, fp = os.path.join(d, tag + "_a.npy"), os.path.join(d, tag + "_p.npy")
            subprocess.run([sys.executable, "-c", code, fa, fp], check=True, env={**os.environ, **env}, timeout=600)
            outs[tag] = (np.load(fa), np.load(fp))
    assert np.array_equal(outs["pipe"][1], outs["plain"][1])
    gate("L\\U pipeline vs plain", outs["pipe"][0], outs["plain"][0], n)


def test_dgesv_host_pipeline_at_config_3_size(fm, ref):
    """BASELINE config 3 through the reference's own call: n = 8192, 64 right-hand sides, pageable host memory, default chunking
    (1024, 1024, 2048, 2048, 2048 columns, a window of two chunks), against the oracle at 1x the gate."""
    lib = fm.load()
    n, nrhs = 8192, 64
    a, b = rnd(7500, n, n), rnd(7501, n, nrhs)
    lu, x, ipiv, info = call_dgesv(lib, a, b)
    assert info == 0
    gate_solve("dgesv_(host pipeline) n=8192", a, b, x, ref.mldivide(a, b), n)
    want_lu, want_piv, _ = ref.lu(a)
    assert np.mean(ipiv == want_piv) >= 0.999
    if np.array_equal(ipiv, want_piv):
        gate("L\\U(host pipeline) n=8192", lu, want_lu, n)


def test_chunked_factorisation_on_a_device_matrix(fm, ref):
    """fm_getrf_chunked: the pipeline's factorisation without the transfers (what tools/chunked_lu_time.py times)"""
    lib = fm.load()
    n = 1664
    a = rnd(7600, n, n)
    A = fm.Flomat.from_numpy(a)
    piv = fm.Pivots(n)
    assert lib.fm_getrf_chunked(n, A.a, A.lda, piv.ptr, piv.info_ptr, 384) == 0
    want_lu, want_piv, _ = ref.lu(a)
    assert np.array_equal(piv.to_numpy(), want_piv)
    gate("fm_getrf_chunked", A.to_numpy(), want_lu, n)


def test_host_pipeline_operand_placements(fm, ref):
    """What the compat symbols accept besides plain host buffers: no right-hand side at all, b already in HBM (solved only after info is
    known: dgesv must leave it untouched for a singular A), ipiv in HBM, and PINNED host memory (asynchronous copies, no staging)."""
    import torch

    lib = fm.load()
    n, nrhs = 2304, 4
    a, b = rnd(7700, n, n), rnd(7701, n, nrhs)
    want_x = ref.mldivide(a, b)
    want_lu, want_piv, _ = ref.lu(a)
    info = c_int(-99)
    # nrhs = 0
    A = a.copy(order="F"); ipiv = np.zeros(n, dtype=np.int32)
    lib.dgesv_(byref(c_int(n)), byref(c_int(0)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, None, byref(c_int(n)), byref(info))
    assert info.value == 0 and np.array_equal(ipiv, want_piv)
    gate("L\\U, nrhs = 0", A, want_lu, n)
    # b on the device
    A = a.copy(order="F"); B = fm.matrix(b)
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.a, byref(c_int(B.lda)), byref(info))
    assert info.value == 0
    gate_solve("dgesv_(host A, device b)", a, b, B.to_numpy(), want_x, n)
    sing = a.copy(order="F"); sing[:, 1000] = 0.0
    B = fm.matrix(b)
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), sing.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.a, byref(c_int(B.lda)), byref(info))
    assert info.value == 1001 and np.array_equal(B.to_numpy(), b)
    # ipiv on the device
    A = a.copy(order="F"); piv = fm.Pivots(n)
    lib.dgetrf_(byref(c_int(n)), byref(c_int(n)), A.ctypes.data, byref(c_int(n)), piv.ptr, byref(info))
    assert info.value == 0 and np.array_equal(piv.to_numpy(), want_piv)
    # pinned host memory
    pa = torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy().T  # F-contiguous view
    pb = torch.empty((nrhs, n), dtype=torch.float64, pin_memory=True).numpy().T
    pa[...] = a; pb[...] = b
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), pa.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, pb.ctypes.data, byref(c_int(n)), byref(info))
    assert info.value == 0 and np.array_equal(ipiv, want_piv)
    gate_solve("dgesv_(pinned host operands)", a, b, pb, want_x, n)
    gate("L\\U(pinned host operands)", pa, want_lu, n)
