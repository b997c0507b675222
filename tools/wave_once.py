"""One 8192 x 1024 x 8192 product (the column block of an 8-GPU sharded product: 3.46 waves of 128x128 tiles) for ncu."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, sci_b200 as fm
fm.init(0); lib = fm.load()
m, n, k = 8192, 1024, 8192
a = torch.randn(m * k, dtype=torch.float64, device="cuda"); b = torch.randn(k * n, dtype=torch.float64, device="cuda"); c = torch.empty(m * n, dtype=torch.float64, device="cuda")
for _ in range(2): lib.fm_dgemm(111, 111, m, n, k, 1.0, a.data_ptr(), m, b.data_ptr(), k, 0.0, c.data_ptr(), m)
fm.sync()
