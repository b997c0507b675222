import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import torch, torch.distributed as dist
import sci_b200 as fm
from sci_b200.multigpu import ColShardTimes
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); fm.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("CS_N", "1024"))
cs = ColShardTimes(n, world, rank)
c = cs.c_tensor(); c.fill_(-7.0)
sumA0 = cs.A.sum().item(); sumB0 = cs.B.sum().item()
torch.cuda.synchronize(); dist.barrier()
cs.step_fused()
torch.cuda.synchronize(); dist.barrier()
cview = c.view(n, n)
blocks = [(r, *__import__("sci_b200.multigpu", fromlist=["column_block"]).column_block(n, world, r)) for r in range(world)]
for r, j0, nl in blocks:
    blk = cview[j0:j0 + nl]
    print(f"rank {rank}: block of rank {r} cols [{j0},{j0+nl}): untouched={(blk == -7.0).float().mean().item():.3f} finite={torch.isfinite(blk).all().item()} abs-sum={blk.abs().sum().item():.6e}", flush=True)
a = cs.A.view(n, n).t(); bcol = cs.B[:n]; want = a @ bcol; got = c[cs.j0 * n:(cs.j0 + 1) * n]
print(f"rank {rank}: own-block diff-sum={(got - want).abs().sum().item():.3e} want-sum={want.abs().sum().item():.3e} tol={64 * n * 2.0 ** -52 * 4 * want.abs().sum().item():.3e} maxdiff={(got-want).abs().max().item():.3e}", flush=True)
bad = ((got - want).abs() > 1e-6).nonzero().flatten()
print(f"rank {rank}: A unchanged={cs.A.sum().item() == sumA0} B unchanged={cs.B.sum().item() == sumB0}", flush=True)
for i in bad[:6].tolist():
    print(f"rank {rank}: row {i}: got={got[i].item():.9f} want={want[i].item():.9f} diff={got[i].item()-want[i].item():.3e}", flush=True)
print(f"rank {rank}: bad rows in col j0: count={bad.numel()} first={bad[:20].tolist()} last={bad[-5:].tolist()}", flush=True)
for jj in (1, 127, 128, 4000, cs.nloc - 1):
    w_ = a @ cs.B[jj * n:(jj + 1) * n]; g_ = c[(cs.j0 + jj) * n:(cs.j0 + jj + 1) * n]
    b_ = ((g_ - w_).abs() > 1e-6).nonzero().flatten()
    print(f"rank {rank}: col j0+{jj}: bad={b_.numel()} first={b_[:8].tolist()}", flush=True)
cs.step_fused(); torch.cuda.synchronize(); dist.barrier()
got2 = c[cs.j0 * n:(cs.j0 + 1) * n]
bad2 = ((got2 - want).abs() > 1e-6).nonzero().flatten()
print(f"rank {rank}: second run bad count={bad2.numel()} first={bad2[:20].tolist()}", flush=True)
cs.peer_array_saved = cs.peer_array; saved = cs.peers; cs.peers = []; cs.peer_array = None
cs.step_fused(); torch.cuda.synchronize(); dist.barrier()
bad3 = ((c[cs.j0 * n:(cs.j0 + 1) * n] - want).abs() > 1e-6).nonzero().flatten()
print(f"rank {rank}: run without peers bad count={bad3.numel()}", flush=True)
cs.peers = saved; cs.peer_array = cs.peer_array_saved
want2 = (cs.A.view(n, n)[:, :] * 1.0).t().contiguous() @ bcol
print(f"rank {rank}: contiguous-A check diff-sum={(got - want2).abs().sum().item():.3e}", flush=True)
print(f"rank {rank} peers={[hex(p) for p in cs.peers]} c_ptr={hex(cs.c_ptr)} verify={cs.verify()}", flush=True)
cs.close()
dist.destroy_process_group()
