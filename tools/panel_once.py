import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
m, w = int(os.environ.get("PM", "8192")), int(os.environ.get("PW", "128"))
g = torch.Generator(device="cuda"); g.manual_seed(3)
piv = fm.Pivots(128)
for _ in range(4):
    t = torch.randn(m * w, dtype=torch.float64, device="cuda", generator=g)
    lib.fm_getrf(m, w, t.data_ptr(), m, piv.ptr, piv.info_ptr)
fm.sync(); print("ok", piv.info())
