"""GPU: the configurations north_star names, compared with the ORACLE on the same bytes at their full sizes, gated at 1 x 64 n eps.

Inputs are SURVEY 8(d)'s: numpy PCG64 `standard_normal` with the listed seeds, column-major.  The oracle is the reference's call
sequence on OpenBLAS (oracle/flomat_oracle.py); its cost on the 16 host cores of a GPU box: dgesv n = 8192 ~1.5 s, dgemm n = 8192 ~2 s,
expm n = 4096 ~5 s, expm n = 8192 ~35 s.  The n = 16384 product is checked on sampled columns against the oracle's dgemm over the full A
(a column of C depends on all of A and one column of B), plus Freivalds over the whole result in tests/test_gpu_parity.py.
"""
import math

import numpy as np
import pytest

from oracle.flomat_oracle import normwise_rel_err, tol
from parity import gate, gate_solve

pytestmark = pytest.mark.gpu


def randn(seed, m, n, scale=1.0):
    a = np.random.default_rng(seed).standard_normal((n, m))  # C-order (n, m) buffer = the column-major m x n matrix's bytes
    if scale != 1.0:
        a *= scale
    return a.T  # F-contiguous view, no copy


def test_mldivide_config3_vs_oracle(fm, ref):
    """BASELINE config 3: `mldivide` (copies + dgesv, flomat.rkt:3289 -> :2145), n = 8192, 64 right-hand sides, seeds 3001 / 3002."""
    n, nrhs = 8192, 64
    a, b = randn(3001, n, n), randn(3002, n, nrhs)
    x = fm.mldivide(fm.matrix(a), fm.matrix(b)).to_numpy()
    want = ref.mldivide(a, b)
    assert ref.last_info == 0
    gate_solve("config3 mldivide n=8192 nrhs=64", a, b, x, want, n)


def test_times_metric_size_vs_oracle(fm, ref):
    """The metric's `times` (n = 8192, seeds 6001 / 6002) and the full n = 4096 product against the oracle's cblas_dgemm, every entry."""
    for n, sa, sb in ((4096, 4001, 4002), (8192, 6001, 6002)):
        a, b = randn(sa, n, n), randn(sb, n, n)
        got = fm.times(fm.matrix(a), fm.matrix(b)).to_numpy()
        want = ref.times(a, b)
        gate(f"times n={n}", got, want, n)
        # entrywise as well: |c_ij - ref_ij| <= 64 n eps (|A| |B|)_ij on sampled entries (the normwise gate cannot hide a bad tile)
        rng = np.random.default_rng(n)
        for _ in range(256):
            i, j = rng.integers(0, n, 2)
            assert abs(got[i, j] - want[i, j]) <= tol(n) * float(np.abs(a[i, :]) @ np.abs(b[:, j]))
        del got, want


def test_times_config4_sampled_columns_vs_oracle(fm, ref):
    """BASELINE config 4 on one GPU (n = 16384): inputs from the device generator (LAPACK's dlarnv stream), columns J of C against the
    oracle's dgemm A * B[:, J] over the FULL A ("oracle check on sampled rows/cols")."""
    n = 16384
    A = fm.randn(n, n, iseed=np.array([41, 1, 2, 3], dtype=np.int32))
    B = fm.randn(n, n, iseed=np.array([42, 1, 2, 3], dtype=np.int32))
    C = fm.times(A, B)
    a = A.to_numpy()
    cols = sorted({0, 1, 127, 128, n // 2 - 1, n // 2, n - 129, n - 1} | {int(c) for c in np.random.default_rng(16384).integers(0, n, 40)})
    bj = np.asfortranarray(np.stack([B.sub(0, c, n, 1).to_numpy()[:, 0] for c in cols], axis=1))
    cj = np.asfortranarray(np.stack([C.sub(0, c, n, 1).to_numpy()[:, 0] for c in cols], axis=1))
    want = ref.times(np.asfortranarray(a), bj)
    gate(f"times n=16384, {len(cols)} sampled columns", cj, want, n)
    # sampled rows too: C[I, :] = A[I, :] * B needs all of B on the host
    b = B.to_numpy()
    rows = sorted({0, 127, 128, n - 1} | {int(r) for r in np.random.default_rng(61).integers(0, n, 28)})
    ci = np.asfortranarray(np.stack([C.sub(r, 0, 1, n).to_numpy()[0, :] for r in rows], axis=0))
    want = ref.times(np.asfortranarray(a[rows, :]), np.asfortranarray(b))
    gate(f"times n=16384, {len(rows)} sampled rows", ci, want, n)


@pytest.mark.parametrize("n,s_want", [(4096, 7.0), (8192, 8.0)])
def test_expm_metric_size_vs_oracle(fm, ref, n, s_want):
    """The metric's expm (n = 8192, A = N(0,1)/sqrt(n), seed 6003, s = 8) and n = 4096 against the oracle's expm of the same bytes
    (11 + s dgemms + dgesv on OpenBLAS), both product orders of the device pipeline, at 1 x 64 n eps."""
    a = randn(6003, n, n, 1.0 / math.sqrt(n))
    want = ref.expm(np.asfortranarray(a))
    s_ref = ref.scaling_exponent(np.asfortranarray(a))
    assert s_ref == s_want
    A = fm.matrix(a)
    for mode, name in ((fm.EXPM_MIN_PRODUCTS, "5+s"), (fm.EXPM_REFERENCE_ORDER, "7+s")):
        E, s = fm.expm(A, mode, return_s=True)
        assert s == s_ref
        gate(f"expm n={n} ({name} products)", E.to_numpy(), want, n)
        del E


def test_expm_batched_config5_every_item_of_a_sample_vs_oracle(fm, ref):
    """BASELINE config 5 generator (seed 5001, N(0,1)/16, 256 x 256 items): 64 consecutive items through fm_expm_batched, EVERY one
    compared with the oracle's expm."""
    n, batch = 256, 64
    host = np.random.default_rng(5001).standard_normal(n * n * batch) / 16.0
    src = fm.Flomat.from_numpy(host.reshape(-1, 1))
    dst = fm.Flomat.alloc(n * n * batch, 1)
    s_out = np.zeros(batch)
    fm.expm_batched(src.a, n, batch, dst.a, fm.EXPM_MIN_PRODUCTS, s_out)
    got = dst.to_numpy().ravel()
    worst = 0.0
    for z in range(batch):
        item = np.asfortranarray(host[z * n * n:(z + 1) * n * n].reshape(n, n, order="F"))
        assert s_out[z] == ref.scaling_exponent(item)
        worst = max(worst, normwise_rel_err(got[z * n * n:(z + 1) * n * n].reshape(n, n, order="F"), ref.expm(item)))
    assert worst <= tol(n), f"worst item: {worst / tol(n):.3f} x 64 n eps"
