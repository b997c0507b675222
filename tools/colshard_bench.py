"""torchrun helper: times the column-sharded `times` (fused all-gather epilogue) at n = 16384 on every rank and verifies it.
FLOMAT_GEMM_NO_COPIER=1 selects the round-1 epilogue (consumer warps store to the peers themselves) for A/B runs."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir))
import torch
import torch.distributed as dist

import sci_b200 as fm
from sci_b200.multigpu import ColShardTimes

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
fm.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cs = ColShardTimes(n, world, rank)
ms = cs._time(cs.step_fused, 5, 3)
ok = cs.verify()
cs.c_tensor().zero_()
ms_n = cs._time(cs.step_nccl, 5, 2)
ok_n = cs.verify()
if rank == 0:
    print(json.dumps({"n": n, "n_gpus": world, "copier": os.environ.get("FLOMAT_GEMM_NO_COPIER") is None, "fused_ms": ms, "fused_ok": ok,
                      "tflops_per_gpu": 2.0 * n ** 3 / ms / 1e9 / world, "nccl_ms": ms_n, "nccl_ok": ok_n}), flush=True)
cs.close()
dist.destroy_process_group()
sys.exit(0 if (ok and ok_n) else 1)
