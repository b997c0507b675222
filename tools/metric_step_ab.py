import math, os, sys, json
sys.path.insert(0, "/root/repo")
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
n = 8192
g = torch.Generator(device="cuda"); g.manual_seed(6003)
t = torch.randn(n * n, dtype=torch.float64, device="cuda", generator=g) / math.sqrt(n)
X = fm.Flomat.alloc(n, n); lib.fm_copy2d(X.a, n, t.data_ptr(), n, n, n); fm.sync()
A = fm.Flomat.alloc(n, n); lib.fm_copy2d(A.a, n, t.data_ptr(), n, n, n)
ts = []
for i in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fm.times(A, X); fm.expm(X, 1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(json.dumps({"tail_rule": os.environ.get("FLOMAT_GEMM_NO_TAIL_RULE") is None, "spread": os.environ.get("FLOMAT_GEMM_NO_SPREAD") is None, "step_ms": [round(x, 2) for x in ts]}))
