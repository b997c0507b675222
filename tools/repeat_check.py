"""Race hunt by repetition: the paths with asynchronous hand-overs (programmatic-dependent split-K chains, the Neumann-product solve
inside expm, the host-operand LU pipeline, the look-ahead LU, fm_mg_expm when several devices are visible) are run many times on the
same input and every result is compared BIT FOR BIT with the first one -- all of them are deterministic by construction (fixed tile and
summation order), so any difference is an ordering bug."""
import math, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
from ctypes import byref, c_int
import numpy as np, torch, sci_b200 as fm
fm.init(0); lib = fm.load()
rng = np.random.default_rng(11)
def rnd(m, n, scale=1.0): return np.asfortranarray(rng.standard_normal((m, n)) * scale)
bad = {}
def repeat(name, fn, reps):
    first = fn(); diff = 0
    for _ in range(reps - 1):
        if not np.array_equal(fn(), first): diff += 1
    bad[name] = diff
    print(f"{name}: {reps} runs, {diff} differ", flush=True)
# split-K products back to back inside one stream, other work interleaved
for n in (256, 512, 768):
    A, B = fm.matrix(rnd(n, n)), fm.matrix(rnd(n, n))
    def f():
        C = fm.times(A, B); D = fm.times(C, B); E = fm.plus(D, C); return fm.times(E, A).to_numpy()
    repeat(f"split-K chain n={n}", f, 300)
for n in (200, 512, 1024):
    X = fm.matrix(rnd(n, n, 1 / math.sqrt(n)))
    repeat(f"expm min-products n={n}", lambda: fm.expm(X, 1).to_numpy(), 200)
    repeat(f"expm reference order n={n}", lambda: fm.expm(X, 0).to_numpy(), 100)
for n in (1000, 2304, 4096):
    A = fm.matrix(rnd(n, n))
    def g():
        P, L, U = fm.flomat_plu(A); return np.concatenate([L.to_numpy().ravel(), U.to_numpy().ravel(), P.to_numpy().ravel()])
    repeat(f"look-ahead LU n={n}", g, 40 if n < 4096 else 15)
os.environ["FLOMAT_CUDA_LU_PIPELINE_MIN_N"] = "2048"
n = 3072
a, b = rnd(n, n), rnd(n, 4)
def h():
    A, B = a.copy(order="F"), b.copy(order="F"); ipiv = np.zeros(n, dtype=np.int32); info = c_int(-1)
    lib.dgesv_(byref(c_int(n)), byref(c_int(4)), A.ctypes.data, byref(c_int(n)), ipiv.ctypes.data, B.ctypes.data, byref(c_int(n)), byref(info))
    return np.concatenate([A.ravel(), B.ravel(), ipiv.astype(np.float64)])
repeat("dgesv_ host pipeline n=3072", h, 25)
if fm.load().fm_device_count() > 1:
    fm.mg_init(0)
    for n in (1100, 2304):
        X = fm.matrix(rnd(n, n, 1 / math.sqrt(n)))
        repeat(f"fm_mg_expm n={n} mode 0", lambda: fm.expm_mg(X, mode=0).to_numpy(), 40)
        repeat(f"fm_mg_expm n={n} mode 1", lambda: fm.expm_mg(X, mode=1).to_numpy(), 40)
    A, B = fm.matrix(rnd(2304, 2304)), fm.matrix(rnd(2304, 2304))
    repeat("fm_mg_dgemm device operands n=2304", lambda: fm.times_mg(A, B).to_numpy(), 40)
print("ALL IDENTICAL" if not any(bad.values()) else f"DIFFERENCES: {bad}")
