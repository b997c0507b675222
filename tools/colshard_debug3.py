import os, sys, ctypes
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
a = cs.A.view(n, n).t()
bl = cs.B.view(cs.nloc, n).t()
want = (a @ bl).t().contiguous().view(-1)     # column-major block
selfpeer = torch.empty(n * n, dtype=torch.float64, device="cuda")
sp_arr = (ctypes.c_void_p * 1)(selfpeer.data_ptr())
def run(mode):
    c.fill_(-7.0); torch.cuda.synchronize(); dist.barrier()
    if mode == "nopeer": arr, k = None, 0
    elif mode == "selfpeer": arr, k = sp_arr, 1
    else: arr, k = cs.peer_array, len(cs.peers)
    cs.check(cs.lib.fm_dgemm_colshard(n, cs.nloc, n, cs.A.data_ptr(), n, cs.B.data_ptr(), n, cs.c_ptr, n, cs.j0, arr, k), "x")
    torch.cuda.synchronize(); dist.barrier()
    got = c[cs.j0 * n:(cs.j0 + cs.nloc) * n]
    bad = ((got - want).abs() > 1e-6).nonzero().flatten()
    tiles = sorted({(int(i) % n // 128, int(i) // n // 128) for i in bad[:4000].tolist()})
    tile_ids = [tn * 16 + tm if tm < 16 else None for tm, tn in tiles]   # band-0 tile index
    if mode == "remote" and rank == 0 and bad.numel():
        idx = bad.tolist()
        tm0, tn0 = int(idx[0]) % n // 128, int(idx[0]) // n // 128
        intile = [(i % n - tm0 * 128, i // n - tn0 * 128) for i in idx if (i % n) // 128 == tm0 and (i // n) // 128 == tn0]
        rows = sorted({r for r, _ in intile}); cols = sorted({cc for _, cc in intile})
        print(f"   first bad tile ({tm0},{tn0}): {len(intile)} bad entries; rows={rows}; cols={cols}", flush=True)
        i0 = idx[0]
        print(f"   entry {i0}: got={got[i0].item():.6f} want={want[i0].item():.6f}", flush=True)
    # compare with the peer's copy of my block
    mine = got.clone()
    if world == 2 and mode == "remote":
        other = torch.empty_like(mine)
        if rank == 0:
            dist.send(mine, dst=1); dist.recv(other, src=1)
        else:
            dist.recv(other, src=0); dist.send(mine, dst=0)
        # `other` = the peer's LOCAL block; my copy of it lives in c at the peer's j0
        pj0 = (1 - rank) * (n // 2)
        mycopy = c[pj0 * n:(pj0 + n // 2) * n]
        print(f"rank {rank}: my copy of peer block vs peer local: mismatches={(mycopy != other).sum().item()}", flush=True)
    print(f"rank {rank} mode={mode}: bad={bad.numel()} tiles(tm,tn)={tiles[:12]} maxdiff={(got-want).abs().max().item():.3e}", flush=True)
for mode in ("nopeer", "nopeer", "selfpeer", "selfpeer", "remote", "remote", "remote"):
    run(mode)
cs.close(); dist.destroy_process_group()
