"""Does PCIe traffic slow the (launch-latency bound) LU panel chain?  fm_getrf / fm_getrf_chunked on a device matrix, alone and while a
second stream copies pinned host memory to the device (h2d), back (d2h) or both."""
import json, os, sys, threading
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
dev = torch.device("cuda", 0)
n = 8192
t = torch.randn(n * n, dtype=torch.float64, device=dev)
LU = fm.Flomat.alloc(n, n); piv = fm.Pivots(n)
host = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); host2 = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(64 << 20, dtype=torch.uint8, device=dev); dbuf2 = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn):
    lib.fm_copy2d(LU.a, n, t.data_ptr(), n, n, n); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
def load(mode):
    # ~100 ms of queued copies
    if "h" in mode:
        with torch.cuda.stream(s_up):
            for _ in range(60): dbuf.copy_(host, non_blocking=True)
    if "d" in mode:
        with torch.cuda.stream(s_dn):
            for _ in range(60): host2.copy_(dbuf2, non_blocking=True)
out = {}
for name, fn in (("fm_getrf", lambda: lib.fm_getrf(n, n, LU.a, n, piv.ptr, piv.info_ptr)),
                 ("chunked_2048", lambda: lib.fm_getrf_chunked(n, LU.a, n, piv.ptr, piv.info_ptr, 2048))):
    row = {}
    for mode in ("", "h", "d", "hd"):
        best = 1e30
        for _ in range(3):
            torch.cuda.synchronize(); load(mode); best = min(best, run(fn)); torch.cuda.synchronize()
        row["alone" if not mode else mode] = round(best, 2)
    out[name] = row
print(json.dumps(out))
