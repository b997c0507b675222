"""Asynchronous host<->device transfers (fm_upload_async / fm_download_async / fm_transfer_sync) through the C ABI."""
import numpy as np
import pytest

from oracle.flomat_oracle import normwise_rel_err, tol

pytestmark = pytest.mark.gpu


def pinned(shape_f):
    """A pinned, column-major float64 host array of shape (m, n)."""
    import torch

    m, n = shape_f
    return torch.empty((n, m), dtype=torch.float64, pin_memory=True).numpy().T


def test_async_upload_is_visible_to_later_work_and_download_waits_for_earlier_work(fm):
    rng = np.random.default_rng(1)
    n = 1500
    ha, hb, hc = pinned((n, n)), pinned((n, n)), pinned((n, n))
    ha[...] = rng.standard_normal((n, n))
    hb[...] = rng.standard_normal((n, n))
    A, B = fm.Flomat.alloc(n, n), fm.Flomat.alloc(n, n)
    A.upload_async(ha)
    B.upload_async(hb)
    C = fm.times(A, B)       # queued after the uploads: sees the data
    C.download_async(hc)     # queued after the product: waits for it
    D = fm.plus(C, C)        # runs beside the download, must not disturb it
    fm.transfer_sync()
    want = ha @ hb
    assert np.abs(hc - want).max() <= 64 * n * 2.0 ** -52 * np.abs(want).max()
    assert np.array_equal(D.to_numpy(), 2.0 * hc)


def test_async_upload_into_recycled_block_waits_for_the_previous_owner(fm):
    """A block given back while a kernel that reads it is still queued, handed out again and filled by an asynchronous upload:
    the copy has to wait for that kernel (the allocator records an event when the block is freed)."""
    rng = np.random.default_rng(2)
    n = 2048
    ha, hz = pinned((n, n)), pinned((n, n))
    ha[...] = rng.standard_normal((n, n))
    hz[...] = 0.0
    fm.Flomat.alloc(8, 8).upload_async(pinned((8, 8)))  # transfers in use from here on: frees record events
    fm.sync()
    A = fm.Flomat.from_numpy(ha)
    ptr = A.a
    prods = [fm.times(A, A) for _ in range(6)]          # a queue of products that read A
    del A                                               # block returns to the cache while they are still queued
    Z = fm.Flomat.alloc(n, n)
    if Z.a != ptr:
        pytest.skip("the allocator did not hand the same block out again")
    Z.upload_async(hz)                                  # zeros into the block the queued products read
    fm.sync()
    want = ha @ ha
    for P in prods:
        assert np.abs(P.to_numpy() - want).max() <= 64 * n * 2.0 ** -52 * np.abs(want).max()
    assert not Z.to_numpy().any()


def test_flat_buffers_and_argument_checks(fm):
    m, n = 37, 21
    rng = np.random.default_rng(3)
    src = np.asfortranarray(rng.standard_normal((m, n)))
    flat = src.ravel(order="F").copy()
    A = fm.Flomat.alloc(m, n).upload_async(flat)         # pageable memory works too (no overlap)
    out = np.empty(m * n)
    A.download_async(out)
    fm.transfer_sync()
    assert np.array_equal(out, flat)
    with pytest.raises(ValueError):
        A.upload_async(np.zeros((n, m), order="F"))
    with pytest.raises(ValueError):
        A.download_async(np.zeros((m, n), dtype=np.float32, order="F"))


# ---------------------------------------------------------------- host-operand product: column-block pipeline (fm_dgemm_host)
@pytest.mark.parametrize("m,n,k,alpha,beta", [(1024, 1024, 1024, 1.0, 1.0), (700, 1300, 515, -0.5, 0.75), (129, 2055, 64, 2.0, 0.0),
                                              (256, 1024, 0, 1.0, 0.5), (5, 3, 4, 1.0, 1.0)])
def test_dgemm_host_pipeline_matches_oracle(fm, ref, m, n, k, alpha, beta):
    rng = np.random.default_rng(m + 3 * n + 7 * k)
    A, B = np.asfortranarray(rng.standard_normal((m, k))), np.asfortranarray(rng.standard_normal((k, n)))
    C0 = np.asfortranarray(rng.standard_normal((m, n)))
    got = fm.times_host(A, B, C0.copy(order="F"), alpha, beta)
    if k == 0:  # BLAS: C := beta*C
        assert np.array_equal(got, beta * C0)
        return
    want = ref.gemm(A, B, C0.copy(order="F"), alpha, beta)
    assert normwise_rel_err(got, want) <= tol(max(m, n, k))
    # the resident path on uploaded operands, same kernels: agreement to rounding of a different tile schedule at most
    Cd = fm.matrix(C0)
    fm.flomat_times(fm.matrix(A), fm.matrix(B), Cd, alpha, beta)
    assert normwise_rel_err(got, Cd.to_numpy()) <= tol(max(m, n, k))


def test_dgemm_host_async_and_views(fm, ref):
    """wait=False leaves C to fm_transfer_sync; leading dimensions larger than the row count (views, flomat.rkt:288-300)."""
    lib = fm.load()
    rng = np.random.default_rng(77)
    m, n, k, lda, ldb, ldc = 300, 1500, 200, 333, 256, 301
    A, B, C = (np.asfortranarray(rng.standard_normal(s)) for s in ((lda, k), (ldb, n), (ldc, n)))
    keep = C.copy(order="F")
    rc = lib.fm_dgemm_host(m, n, k, 1.0, A.ctypes.data, lda, B.ctypes.data, ldb, 0.0, C.ctypes.data, ldc, 1)
    assert rc == 0, lib.fm_last_error()
    fm.transfer_sync()
    want = ref.gemm(np.asfortranarray(A[:m]), np.asfortranarray(B[:k]))
    assert normwise_rel_err(C[:m], want) <= tol(max(m, n, k))
    assert np.array_equal(C[m:], keep[m:])  # rows outside the view untouched
    # a device pointer is refused, not misread
    d = fm.matrix(A)
    assert lib.fm_dgemm_host(m, n, k, 1.0, d.a, lda, B.ctypes.data, ldb, 0.0, C.ctypes.data, ldc, 0) != 0


def test_compat_cblas_dgemm_host_large_takes_pipeline(fm, ref):
    """cblas_dgemm with host pointers at a size past the pipeline threshold (N >= 1024, MNK >= 2^30), alpha = beta = 1 into zeros
    exactly as flomat.rkt:888 calls it."""
    lib = fm.load()
    n = 1024
    rng = np.random.default_rng(5)
    A, B = np.asfortranarray(rng.standard_normal((n, n))), np.asfortranarray(rng.standard_normal((n, n)))
    C = np.zeros((n, n), order="F")
    lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, A.ctypes.data, n, B.ctypes.data, n, 1.0, C.ctypes.data, n)
    assert normwise_rel_err(C, ref.gemm(A, B)) <= tol(n)


# ---------------------------------------------------------------- pageable host memory: the library's own staging (xfer.cu)
@pytest.mark.parametrize("m,n,ldh", [(1500, 1500, 1500), (4096, 5200, 4096), (1000, 3000, 1111), (2049, 1025, 2050)])
def test_pageable_upload_download_round_trip_is_exact(fm, m, n, ldh):
    """Unpinned host matrices (what the reference's GC malloc gives, flomat.rkt:437) go through pinned bounce buffers filled by
    several host threads; contiguous and strided (lda > m views) layouts, more chunks than bounce buffers."""
    lib = fm.load()
    rng = np.random.default_rng(m * 31 + n)
    H = np.asfortranarray(rng.standard_normal((ldh, n)))
    D = fm.Flomat.alloc(m, n)
    assert lib.fm_upload(D.a, m, H.ctypes.data, ldh, m, n) == 0, lib.fm_last_error()
    H2 = np.full((ldh, n), -7.0, order="F")
    assert lib.fm_download(H2.ctypes.data, ldh, D.a, m, m, n) == 0, lib.fm_last_error()
    assert np.array_equal(H2[:m], H[:m])
    assert np.all(H2[m:] == -7.0)  # rows between the view's end and the leading dimension are not written
    # and the device really holds the matrix: a kernel result computed from it
    assert np.array_equal(fm.dot_plus(D, D).to_numpy(), 2.0 * H[:m])


def test_pageable_staging_repeated_calls_do_not_mix_buffers(fm):
    """Back-to-back staged uploads and downloads of different matrices share the bounce ring; each result must be its own."""
    rng = np.random.default_rng(9)
    mats = [np.asfortranarray(rng.standard_normal((1200 + 64 * i, 1300))) for i in range(6)]
    dev = [fm.matrix(x) for x in mats]
    for x, d in zip(mats, dev):
        assert np.array_equal(d.to_numpy(), x)


def test_compat_dgesv_pageable_operands_through_staging(fm, ref):
    """dgesv_ as flomat.rkt:2140 calls it, at a size where the pageable operands take the library's staged copies (> 4 MB):
    X in B, L\\U in A, 1-based ipiv on the host -- against the oracle's dgesv on the same inputs."""
    from ctypes import byref, c_int

    lib = fm.load()
    n, nrhs = 1100, 600
    rng = np.random.default_rng(2024)
    a, b = np.asfortranarray(rng.standard_normal((n, n))), np.asfortranarray(rng.standard_normal((n, nrhs)))
    A, B = a.copy(order="F"), b.copy(order="F")
    ipiv, info = np.zeros(n, dtype=np.int32), c_int(-1)
    lib.dgesv_(byref(c_int(n)), byref(c_int(nrhs)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    assert info.value == 0
    want_lu, want_piv, _ = ref.lu(a)
    assert np.array_equal(ipiv, want_piv)
    assert normwise_rel_err(A, want_lu) <= tol(n)
    assert normwise_rel_err(B, ref.mldivide(a, b)) <= tol(n) * 50  # cond(A) slack as in test_compat_lapack_symbols_host_pointers
    assert np.abs(a @ B - b).max() <= 1e-9 * np.abs(b).max() * n


# ---------------------------------------------------------------- expm with the result delivered to host memory
@pytest.mark.parametrize("n,pin", [(8, False), (64, True), (700, False), (1500, True), (2304, False)])
@pytest.mark.parametrize("mode", [0, 1])
def test_expm_to_host_matches_device_result(fm, n, pin, mode):
    """fm_expm with a HOST result pointer: the last squaring runs in column blocks whose downloads overlap the remaining blocks.
    Same arithmetic as the resident call up to the tile configuration of the last product: normwise within 64 n eps."""
    rng = np.random.default_rng(n)
    a = np.asfortranarray(rng.standard_normal((n, n)) / np.sqrt(n))
    A = fm.matrix(a)
    want, s_want = fm.expm(A, mode, return_s=True)
    want = want.to_numpy()
    out = pinned((n, n)) if pin else np.empty((n, n), order="F")
    out[...] = np.nan
    _, s = fm.expm_to_host(A, out, mode, return_s=True)
    fm.transfer_sync()
    assert s == s_want
    assert np.isfinite(out).all()
    assert normwise_rel_err(out, want) <= tol(n)


def test_expm_to_host_without_squaring(fm):
    """s <= 0 (reference behaviour, expm.rkt:205-210: no squaring at all): the result comes straight from the solve."""
    a = np.asfortranarray(np.diag([0.1, 0.05, 0.02]))
    A = fm.matrix(a)
    want = fm.expm(A).to_numpy()
    out = np.empty((3, 3), order="F")
    fm.expm_to_host(A, out)
    fm.transfer_sync()
    assert np.array_equal(out, want)
