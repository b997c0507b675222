"""Developer tool: per-block timeline of fm_dgemm_host (FLOMAT_HOST_GEMM_TRACE=1), pinned and pageable operands."""
import os, sys
os.environ["FLOMAT_HOST_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch
import sci_b200 as fm
fm.init(0)
lib = fm.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pin = lambda: torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy().T
rng = np.random.default_rng(1)
hA, hB, hC = pin(), pin(), pin()
hA[...] = rng.standard_normal((n, n)); hB[...] = rng.standard_normal((n, n))
pA, pB, pC = hA.copy(order="F"), hB.copy(order="F"), np.zeros((n, n), order="F")
for rep in range(2):  # the second call is the warm one
    print(f"# pinned operands, beta = 0, call {rep}", file=sys.stderr, flush=True)
    fm.times_host(hA, hB, hC, 1.0, 0.0)
for rep in range(2):
    pC.fill(0.0)
    print(f"# pageable operands, alpha = beta = 1 into zeros (the reference's call), call {rep}", file=sys.stderr, flush=True)
    lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, pA.ctypes.data, n, pB.ctypes.data, n, 1.0, pC.ctypes.data, n)
