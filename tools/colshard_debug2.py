import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, torch.distributed as dist
import sci_b200 as fm
from sci_b200.multigpu import ColShardTimes
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); fm.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("CS_N", "16384"))
cs = ColShardTimes(n, world, rank)
c = cs.c_tensor()
for trial in range(3):
    c.fill_(-7.0)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 1:   # only rank 1 computes and pushes into rank 0
        cs.check(cs.lib.fm_dgemm_colshard(n, cs.nloc, n, cs.A.data_ptr(), n, cs.B.data_ptr(), n, cs.c_ptr, n, cs.j0, cs.peer_array, len(cs.peers)), "x")
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        cv = c.view(n, n)
        mine = cv[:n // 2]; theirs = cv[n // 2:]
        touched = (mine != -7.0).nonzero()
        print(f"trial {trial}: rank0 own block entries touched by rank 1: {touched.shape[0]} first={touched[:10].tolist()}; their block untouched entries: {(theirs == -7.0).sum().item()}", flush=True)
    dist.barrier()
    # now check rank 1's copy on rank 0 against rank 1's local copy
    blk = c[(n // 2) * n:].clone()
    if rank == 1:
        dist.send(blk, dst=0)
    else:
        other = torch.empty_like(blk); dist.recv(other, src=1)
        diff = (other != blk).nonzero().flatten()
        print(f"trial {trial}: rank0 copy of block1 vs rank1 local: mismatches={diff.numel()} first={[(int(i) // n, int(i) % n) for i in diff[:10].tolist()]}", flush=True)
        if diff.numel():
            i = diff[0].item(); print("   values", blk[i].item(), other[i].item())
cs.close(); dist.destroy_process_group()
