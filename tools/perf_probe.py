"""Developer tool: device-resident timings of every hot-path entry point (CUDA events, best of N)."""
import sys, os, math, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import numpy as np, torch
import sci_b200 as fm
fm.init(0)
lib = fm.load()
PEAK = 37.0

def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e18
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def randmat(n, m=None, scale=1.0, seed=0):
    m = m or n
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    t = torch.randn(n * m, dtype=torch.float64, device="cuda", generator=g) * scale
    out = fm.Flomat.alloc(n, m)
    lib.fm_copy2d(out.a, n, t.data_ptr(), n, n, m); fm.sync()
    return out

what = sys.argv[1:] or ["ewise", "level12", "lu", "expm", "batched"]
res = {}
if "ewise" in what:
    for n in (8192,):
        A, B = randmat(n, seed=1), randmat(n, seed=2); C = fm.Flomat.alloc(n, n)
        for name, fn, bpe in [("plus", lambda: fm.dot_plus(A, B, C), 24), (".*", lambda: fm.dot_times(A, B, C), 24),
                              ("scale", lambda: fm.dot_times(A, 2.0, C), 16), ("copy", lambda: lib.fm_copy2d(C.a, n, A.a, n, n, n), 16),
                              ("fill", lambda: lib.fm_fill(C.a, n, n, n, 1.0), 8), ("norm1", lambda: fm.norm(A, 1), 8),
                              ("axpy!", lambda: fm.flomat_plus_bang(A, C), 24)]:
            ms = timeit(fn, reps=5, warm=2)
            print(f"ewise {name} n={n}: {ms:.3f} ms  {bpe*n*n/ms/1e6:.0f} GB/s", flush=True)
            res[f"ewise_{name}_{n}"] = bpe * n * n / ms / 1e6
        del A, B, C
if "level12" in what:
    n = 8192
    A = randmat(n, seed=11); B = fm.Flomat.alloc(n, n); x = randmat(n, 1, seed=12); y = fm.Flomat.alloc(n, 1)
    iseed = np.array([1, 2, 3, 5], dtype=np.int32)
    for name, fn, bpe in [("transpose", lambda: lib.fm_transpose(n, n, A.a, n, B.a, n), 16),
                          ("triu", lambda: lib.fm_tri(0, 0, n, n, A.a, n, B.a, n), 16),
                          (".exp", lambda: lib.fm_map(11, n, n, A.a, n, B.a, n), 16),
                          (".sin", lambda: lib.fm_map(5, n, n, A.a, n, B.a, n), 16),
                          (".sqrt", lambda: lib.fm_map(9, n, n, A.a, n, B.a, n), 16),
                          ("gemv N", lambda: lib.fm_gemv(111, n, n, 1.0, A.a, n, x.a, 1, 0.0, y.a, 1), 8),
                          ("gemv T", lambda: lib.fm_gemv(112, n, n, 1.0, A.a, n, x.a, 1, 0.0, y.a, 1), 8),
                          ("colsums", lambda: lib.fm_colsums(n, n, A.a, n, y.a), 8),
                          ("rowsums", lambda: lib.fm_rowsums(n, n, A.a, n, y.a), 8),
                          ("nrm2 (n^2)", lambda: fm.flcolumn_norm(fm.Flomat(n * n, 1, A.a, n * n, A._buf)), 8),
                          ("rand", lambda: lib.fm_larnv(1, iseed.ctypes.data, n * n, B.a), 8),
                          ("randn", lambda: lib.fm_larnv(3, iseed.ctypes.data, n * n, B.a), 8)]:
        ms = timeit(fn, reps=5, warm=2)
        print(f"level12 {name} n={n}: {ms:.3f} ms  {bpe*n*n/ms/1e6:.0f} GB/s", flush=True)
        res[f"l12_{name}"] = bpe * n * n / ms / 1e6
    del A, B, x, y
if "lu" in what:
    for n in [int(x) for x in os.environ.get("LU_N", "1024,2048,4096,8192").split(",")]:
        A0 = randmat(n, seed=3); A = fm.Flomat.alloc(n, n)
        piv = fm.Pivots(n)
        def f():
            lib.fm_copy2d(A.a, n, A0.a, n, n, n)
            lib.fm_getrf(n, n, A.a, n, piv.ptr, piv.info_ptr)
        ms_copy = timeit(lambda: lib.fm_copy2d(A.a, n, A0.a, n, n, n), reps=3)
        ms = timeit(f, reps=3) - ms_copy
        print(f"getrf n={n}: {ms:.3f} ms  {(2/3)*n**3/ms/1e9:.2f} TFLOP/s ({(2/3)*n**3/ms/1e9/PEAK*100:.1f}%)", flush=True)
        res[f"getrf_{n}"] = ms
        Bm = randmat(n, 64, seed=4); X = fm.Flomat.alloc(n, 64)
        def g():
            lib.fm_copy2d(X.a, n, Bm.a, n, n, 64)
            lib.fm_getrs(n, 64, A.a, n, piv.ptr, X.a, n)
        ms = timeit(g, reps=3)
        print(f"getrs n={n} nrhs=64: {ms:.3f} ms", flush=True)
        ms = timeit(lambda: fm.mldivide(A0, Bm), reps=3)
        w = (2/3)*n**3 + 2*n*n*64
        print(f"mldivide n={n} nrhs=64 (copies+gesv): {ms:.3f} ms  {w/ms/1e9:.2f} TFLOP/s ({w/ms/1e9/PEAK*100:.1f}%)", flush=True)
        res[f"mldivide_{n}"] = ms
        del A0, A, Bm, X
if "inv" in what:
    for n in [int(x) for x in os.environ.get("INV_N", "2048,8192").split(",")]:
        A0 = randmat(n, seed=31)
        ms = timeit(lambda: fm.inv(A0), reps=3)
        print(f"inv n={n} (copy + getrf + getri): {ms:.3f} ms  {2*n**3/ms/1e9:.2f} TFLOP/s on 2n^3 ({2*n**3/ms/1e9/PEAK*100:.1f}%)", flush=True)
        res[f"inv_{n}"] = ms
        del A0
if "chol" in what:
    for n in [int(x) for x in os.environ.get("CHOL_N", "4096,8192").split(",")]:
        M = randmat(n, seed=21)
        A0 = fm.plus(fm.times(M, fm.transpose(M)), fm.dot_times(fm.eye(n), float(n))); del M
        A = fm.Flomat.alloc(n, n); info = fm.Pivots(0)
        def f():
            lib.fm_copy2d(A.a, n, A0.a, n, n, n)
            lib.fm_potrf(b"L", n, A.a, n, info.info_ptr)
        ms = timeit(f, reps=3) - timeit(lambda: lib.fm_copy2d(A.a, n, A0.a, n, n, n), reps=3)
        print(f"potrf n={n}: {ms:.3f} ms  {n**3/3/ms/1e9:.2f} TFLOP/s on n^3/3 ({n**3/3/ms/1e9/PEAK*100:.1f}%), info {info.info()}", flush=True)
        res[f"potrf_{n}"] = ms
        del A0, A
if "expm" in what:
    for n in [int(x) for x in os.environ.get("EXPM_N", "512,2048,8192").split(",")]:
        X = randmat(n, scale=1/math.sqrt(n), seed=5)
        for mode in (0, 1):
            _, s = fm.expm(X, mode, return_s=True)
            ms = timeit(lambda: fm.expm(X, mode), reps=2)
            wmin = (5+s)*2*n**3 + 8/3*n**3
            print(f"expm n={n} mode={mode} s={s}: {ms:.3f} ms  W_min {wmin/ms/1e9:.2f} TFLOP/s ({wmin/ms/1e9/PEAK*100:.1f}%)", flush=True)
            res[f"expm_{n}_{mode}"] = ms
        del X
if "batched" in what:
    n, batch = 256, int(os.environ.get("BATCH", "512"))
    g = torch.Generator(device="cuda"); g.manual_seed(5001)
    src = torch.randn(n*n*batch, dtype=torch.float64, device="cuda", generator=g) / 16.0
    dst = torch.empty_like(src)
    s_out = np.zeros(batch)
    for mode in (0, 1):
        fm.expm_batched(src.data_ptr(), n, batch, dst.data_ptr(), mode, s_out)
        ms = timeit(lambda: fm.expm_batched(src.data_ptr(), n, batch, dst.data_ptr(), mode, s_out), reps=2)
        s = s_out.max()
        wmin = batch * ((5+s)*2*n**3 + 8/3*n**3)
        print(f"expm_batched {batch}x{n} mode={mode} s={s}: {ms:.3f} ms  W_min {wmin/ms/1e9:.2f} TFLOP/s ({wmin/ms/1e9/PEAK*100:.1f}%)", flush=True)
        res[f"expm_batched_{mode}"] = ms
json.dump(res, open("gpurun_out/perf_probe.json", "w"))
