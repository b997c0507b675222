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


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n", [1024, 1100, 2304])
def test_mg_expm_single_matrix_vs_oracle(fm, ref, ndev, n, mode):
    """fm_mg_expm (one large expm over the group, every product column-sharded with the fused all-gather, LU replicated, solves
    sharded by right-hand sides) against the oracle at 1x the gate, device and host operands, both product orders; ragged column
    blocks at n = 1100.  (One device: the same entry point takes the one-device route.)"""
    a = rnd(70 + n, n, n, 1.0 / np.sqrt(n))
    want, s_ref = ref.expm(a), ref.scaling_exponent(a)  # (both product orders are gated against the reference's own order)
    got, s = fm.expm_mg(fm.matrix(a), mode=mode, return_s=True)
    assert s == s_ref
    gate(f"fm_mg_expm device n={n} mode={mode} on {ndev} device(s)", got.to_numpy(), want, n)
    got_h, s = fm.expm_mg(a, mode=mode, return_s=True)
    assert s == s_ref
    gate(f"fm_mg_expm host n={n} mode={mode} on {ndev} device(s)", got_h, want, n)
    # against the one-device pipeline: the same algorithm, other tile shapes
    one = fm.expm(fm.matrix(a), mode).to_numpy()
    gate(f"fm_mg_expm vs fm_expm n={n} mode={mode}", got_h, one, n)


def test_mg_expm_is_repeatable_and_leaves_the_input_alone(fm, ndev):
    n = 1536
    a = rnd(81, n, n, 1.0 / np.sqrt(n))
    A = fm.matrix(a)
    first = fm.expm_mg(A).to_numpy()
    again = fm.expm_mg(A).to_numpy()
    assert np.array_equal(first, again)  # fixed tile order, fixed split: bit-identical from call to call
    assert np.array_equal(A.to_numpy(), a)


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


@pytest.mark.parametrize("npeers", [2, 7])
def test_peer_store_epilogue_under_lsu_pressure_single_gpu(fm, npeers):
    """Regression guard for the stage-release hazard of the GEMM main loop (dgemm.cu: the mbarrier arrive that frees a pipeline stage
    must not overtake the stage's last LDS; seen in round 1 only with NVLink stores in flight on 2 GPUs).  One GPU is enough to
    exercise it: the "peers" are other buffers of the same device, so the peer-store epilogue runs (consumer-side stores for 2 peers,
    the forwarding warps for 7), while a second stream saturates the memory pipes with elementwise traffic.  120 products per
    configuration must equal a quiet run bit for bit, in the local C and in every "peer"."""
    import ctypes

    import torch

    lib = fm.load()
    n, nloc = 2048, 1024
    A = fm.randn(n, n, iseed=np.array([7, 7, 7, 7], dtype=np.int32))
    B = fm.randn(n, nloc, iseed=np.array([9, 9, 9, 9], dtype=np.int32))
    quiet = fm.Flomat.alloc(n, n)
    fm._lib.check(lib.fm_fill(quiet.a, n, n, n, 0.0), "fm_fill")
    fm._lib.check(lib.fm_dgemm_colshard(n, nloc, n, A.a, n, B.a, n, quiet.a, n, 512, None, 0), "fm_dgemm_colshard")
    want = quiet.to_numpy()
    C = fm.Flomat.alloc(n, n)
    peers = [fm.Flomat.alloc(n, n) for _ in range(npeers)]
    arr = (ctypes.c_void_p * npeers)(*[p.a for p in peers])
    side = torch.cuda.Stream()
    x = torch.randn(1 << 26, dtype=torch.float64, device="cuda")
    y = torch.randn(1 << 26, dtype=torch.float64, device="cuda")
    bad = 0
    for rep in range(120):
        for buf in [C] + peers:
            fm._lib.check(lib.fm_fill(buf.a, n, n, n, 0.0), "fm_fill")
        fm.sync()
        with torch.cuda.stream(side):
            for _ in range(6):
                x.add_(y)  # 1.6 GB of traffic per pass beside the product
        fm._lib.check(lib.fm_dgemm_colshard(n, nloc, n, A.a, n, B.a, n, C.a, n, 512, arr, npeers), "fm_dgemm_colshard")
        torch.cuda.synchronize()
        if rep % 8 == 0 or rep == 119:  # full comparison of every copy on a sample of the repetitions, checksums on the rest
            for buf in [C] + peers:
                bad += int(not np.array_equal(buf.to_numpy(), want))
        else:
            for buf in [C] + peers:
                bad += int(fm.norm(fm.minus(buf, quiet), "max") != 0.0)
    assert bad == 0
