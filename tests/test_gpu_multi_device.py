"""GPU: the single-process multi-GPU entry points (fm_mg_*, include/flomat_cuda.h tier 3b) -- ONE Python process, ONE thread, ctypes
only, no torchrun: the way `flomat*` (flomat.rkt:905-914) can reach every device of the box.  The group uses all visible devices
(1 on the round-end test box, where the calls take the single-device routes of the same entry points; 2..8 under `gpurun --gpus N`).
Results are compared with the oracle at 1 x 64 n eps; n = 16384 by Freivalds' check on the device."""
import numpy as np
import pytest

from oracle.flomat_oracle import normwise_rel_err, tol
from parity import gate

pytestmark = pytest.mark.gpu


def rnd(seed, m, n, scale=1.0):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((m, n)) * scale)


@pytest.fixture(scope="module")
def ndev(fm):
    return fm.mg_init(0)


def test_group_covers_visible_devices(fm, ndev):
    assert ndev == min(fm.load().fm_device_count(), 8) == fm.load().fm_mg_device_count()
    assert fm.mg_init(ndev) == ndev  # idempotent


@pytest.mark.parametrize("m,n,k", [(4096, 4096, 4096), (1000, 2304, 700), (513, 3000, 129), (256, 300, 64)])
def test_mg_dgemm_host_operands_vs_oracle(fm, ref, ndev, m, n, k):
    """Pageable host operands, as cblas_dgemm receives them from the reference; alpha/beta and ragged column blocks included."""
    a, b, c = rnd(1, m, k), rnd(2, k, n), rnd(3, m, n)
    got = fm.times_host_mg(a, b)
    gate(f"fm_mg_dgemm host {m}x{n}x{k} on {ndev} device(s)", got, ref.gemm(a, b), max(m, n, k))
    got = fm.times_host_mg(a, b, c.copy(order="F"), alpha=-0.75, beta=0.5)
    gate(f"fm_mg_dgemm host alpha/beta {m}x{n}x{k}", got, ref.gemm(a, b, c.copy(order="F"), alpha=-0.75, beta=0.5), max(m, n, k))


@pytest.mark.parametrize("m,n,k", [(4096, 4096, 4096), (777, 2500, 1030)])
def test_mg_dgemm_device_operands_vs_oracle(fm, ref, ndev, m, n, k):
    a, b, c = rnd(4, m, k), rnd(5, k, n), rnd(6, m, n)
    A, B = fm.matrix(a), fm.matrix(b)
    got = fm.times_mg(A, B).to_numpy()
    gate(f"fm_mg_dgemm device {m}x{n}x{k} on {ndev} device(s)", got, ref.gemm(a, b), max(m, n, k))
    C = fm.matrix(c)
    fm.times_mg(A, B, C, alpha=2.0, beta=-1.0)
    gate(f"fm_mg_dgemm device alpha/beta {m}x{n}x{k}", C.to_numpy(), ref.gemm(a, b, c.copy(order="F"), alpha=2.0, beta=-1.0), max(m, n, k))
    # the result is ordered on the caller's stream: a single-device product that consumes it right away sees all of it
    D = fm.times(fm.matrix(np.eye(m, order="F")), C).to_numpy()
    assert np.array_equal(D, C.to_numpy())


def test_mg_dgemm_config4_freivalds(fm, ndev):
    """BASELINE config 4 through ONE call: n = 16384 on all devices, device-generated inputs, Freivalds C x == A (B x)."""
    n = 16384
    A, B = fm.randn(n, n, iseed=np.array([41, 1, 2, 3], dtype=np.int32)), fm.randn(n, n, iseed=np.array([42, 1, 2, 3], dtype=np.int32))
    C = fm.times_mg(A, B)
    scale = fm.norm(A, 1) * fm.norm(B, 1)
    for k in range(2):
        x = fm.randn(n, 1, iseed=np.array([43 + k, 5, 6, 7], dtype=np.int32))
        lhs = fm.flomat_times_vector(C, x)
        rhs = fm.flomat_times_vector(A, fm.flomat_times_vector(B, x))
        assert fm.norm(fm.minus(lhs, rhs), 1) / (scale * fm.norm(x, 1)) <= tol(n)
    # and bit for bit what one device computes (same tiles, same k order, whichever device owns the column block)
    assert fm.norm(fm.minus(C, fm.times(A, B)), "max") == 0.0


def test_mg_dgemm_rejects_what_it_cannot_shard(fm, ndev):
    lib = fm.load()
    a = rnd(7, 300, 300)
    A = fm.matrix(a)
    c = np.zeros((300, 300), order="F")
    assert lib.fm_mg_dgemm(112, 111, 300, 300, 300, 1.0, a.ctypes.data, 300, a.ctypes.data, 300, 0.0, c.ctypes.data, 300) == -1002  # trans
    assert lib.fm_mg_dgemm(111, 111, 300, 300, 300, 1.0, A.a, 300, a.ctypes.data, 300, 0.0, c.ctypes.data, 300) == -1002  # mixed memory
    assert lib.fm_mg_dgemm(111, 111, 300, 300, 300, 1.0, a.ctypes.data, 200, a.ctypes.data, 300, 0.0, c.ctypes.data, 300) == -1002  # lda


@pytest.mark.parametrize("n,batch", [(64, 19), (256, 40)])
def test_mg_expm_batched_vs_oracle(fm, ref, ndev, n, batch):
    rng = np.random.default_rng(5001)
    host = rng.standard_normal(n * n * batch) / np.sqrt(n) * 2.0
    want = [ref.expm(np.asfortranarray(host[z * n * n:(z + 1) * n * n].reshape(n, n, order="F"))) for z in range(batch)]
    # host operands
    out, s_out = np.zeros_like(host), np.zeros(batch)
    fm.expm_batched_mg(host.ctypes.data, n, batch, out.ctypes.data, fm.EXPM_MIN_PRODUCTS, s_out)
    # device operands on the caller's device
    src = fm.Flomat.from_numpy(host.reshape(-1, 1))
    dst = fm.Flomat.alloc(n * n * batch, 1)
    s_dev = np.zeros(batch)
    fm.expm_batched_mg(src.a, n, batch, dst.a, fm.EXPM_MIN_PRODUCTS, s_dev)
    got_dev = dst.to_numpy().ravel()
    for z in range(batch):
        item = np.asfortranarray(host[z * n * n:(z + 1) * n * n].reshape(n, n, order="F"))
        assert s_out[z] == s_dev[z] == ref.scaling_exponent(item)
        for got in (out, got_dev):
            assert normwise_rel_err(got[z * n * n:(z + 1) * n * n].reshape(n, n, order="F"), want[z]) <= tol(n), z


def test_second_device_context(fm, ref):
    """fm_init on another device of the same process gets its own context (allocator, workspaces, kernel attributes);
    going back to the first device keeps working.  Needs 2 devices."""
    lib = fm.load()
    if lib.fm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    a, b = rnd(8, 700, 700), rnd(9, 700, 700)
    want = ref.gemm(a, b)
    try:
        fm.init(1)
        A1, B1 = fm.matrix(a), fm.matrix(b)
        assert normwise_rel_err(fm.times(A1, B1).to_numpy(), want) <= tol(700)
        assert normwise_rel_err(fm.mldivide(A1, B1).to_numpy(), ref.mldivide(a, b)) <= tol(700)
        assert normwise_rel_err(fm.expm(fm.dot_times(A1, 1.0 / 26.0)).to_numpy(), ref.expm(np.asfortranarray(a / 26.0))) <= tol(700)
        del A1, B1
    finally:
        fm.init(0)
    assert normwise_rel_err(fm.times(fm.matrix(a), fm.matrix(b)).to_numpy(), want) <= tol(700)
