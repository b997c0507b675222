"""Developer tool: `times` on HOST operands -- upload / product / download one after the other against the column-block pipeline
of fm_dgemm_host (pinned and pageable host memory).  Wall clock around synchronous calls."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch
import sci_b200 as fm
fm.init(0)
lib = fm.load()
ns = [int(x) for x in sys.argv[1:]] or [4096, 8192]
for n in ns:
    pin = lambda: torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy().T
    rng = np.random.default_rng(n)
    hA, hB, hC = pin(), pin(), pin()
    hA[...] = rng.standard_normal((n, n)); hB[...] = rng.standard_normal((n, n))
    pA, pB, pC = np.asfortranarray(hA.copy(order="F")), np.asfortranarray(hB.copy(order="F")), np.zeros((n, n), order="F")

    def serial():
        A, B = fm.Flomat.alloc(n, n), fm.Flomat.alloc(n, n)
        lib.fm_upload(A.a, n, hA.ctypes.data, n, n, n); lib.fm_upload(B.a, n, hB.ctypes.data, n, n, n)
        C = fm.flomat_times(A, B)
        lib.fm_download(hC.ctypes.data, n, C.a, n, n, n)

    def best(fn, reps=4):
        fn(); fm.sync()
        b = 1e9
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); fm.sync(); b = min(b, time.perf_counter() - t0)
        return b * 1e3

    flop = 2.0 * n ** 3
    for name, fn in [("serial upload/product/download (pinned)", serial),
                     ("fm_dgemm_host pipeline (pinned)", lambda: fm.times_host(hA, hB, hC, 1.0, 0.0)),
                     ("cblas_dgemm host pointers (pageable)", lambda: lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, pA.ctypes.data, n, pB.ctypes.data, n, 0.0, pC.ctypes.data, n))]:
        ms = best(fn)
        print(f"n={n} {name}: {ms:.2f} ms  {flop / ms / 1e9:.2f} TFLOP/s end to end", flush=True)
    ref = fm.flomat_times(fm.matrix(hA), fm.matrix(hB)).to_numpy()
    print("  pinned result equals resident product:", np.array_equal(hC, ref), " pageable:", np.array_equal(pC, ref), flush=True)
    # the reference's exact call (flomat.rkt:888 via :905): alpha = beta = 1 into a zero matrix, pageable memory
    def ref_call():
        pC.fill(0.0)
        t0 = time.perf_counter()
        lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, pA.ctypes.data, n, pB.ctypes.data, n, 1.0, pC.ctypes.data, n)
        return (time.perf_counter() - t0) * 1e3
    ref_call()
    ms = min(ref_call() for _ in range(4))
    print(f"n={n} cblas_dgemm alpha=beta=1, C=zeros (pageable): {ms:.2f} ms  {flop / ms / 1e9:.2f} TFLOP/s end to end", flush=True)
    # raw staged copies of one n x n matrix
    D = fm.Flomat.alloc(n, n)
    def up():
        lib.fm_upload(D.a, n, pA.ctypes.data, n, n, n); fm.sync()
    def down():
        lib.fm_download(pC.ctypes.data, n, D.a, n, n, n)
    for name, fn in (("fm_upload pageable", up), ("fm_download pageable", down)):
        ms = best(fn)
        print(f"n={n} {name}: {ms:.2f} ms  {n * n * 8 / ms / 1e6:.1f} GB/s", flush=True)
