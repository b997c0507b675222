"""Asynchronous host<->device transfers (fm_upload_async / fm_download_async / fm_transfer_sync) through the C ABI."""
import numpy as np
import pytest

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
