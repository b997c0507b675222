"""Multi-GPU paths (one process per GPU, torch.distributed for the plumbing).

* `times` column-sharded (BASELINE config 4): A is replicated, rank r owns the contiguous column block
  [j0, j0+nloc) of B and C.  C is reassembled on every rank.  Two implementations of the reassembly:
    - "fused":  the GEMM epilogue itself stores every finished C tile into the local C *and* into every peer's C through
      CUDA-IPC mapped peer pointers over NVLink (fm_dgemm_colshard) -- the all-gather rides inside the GEMM, tile by
      tile, and only a barrier follows;
    - "nccl":   local GEMM, then an in-place ncclAllGather (the baseline the fused kernel is compared with).
* batched expm shards by items and needs no collective (`item_block`).

The partitioning helpers are pure host code and are exercised on CPU with the gloo backend (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np


def column_block(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous column block of rank `rank`: (first column, number of columns).  Blocks are multiples of 128 columns
    (the GEMM tile width) except possibly the last, so no tile straddles two owners."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    tiles = -(-n // 128)
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    j0, j1 = min(t0 * 128, n), min(t1 * 128, n)
    return j0, j1 - j0


def item_block(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of batch items owned by `rank` (batched expm): (first item, count)."""
    base, extra = divmod(batch, world)
    z0 = rank * base + min(rank, extra)
    return z0, base + (1 if rank < extra else 0)


def sub_blocks(nloc: int, max_blocks: int = 4, min_cols: int = 1024) -> list[tuple[int, int]]:
    """The sub-blocks (first column, width) a rank's column block is cut into for the end-to-end pipeline (upload i+1 | product i |
    download i-1): at most `max_blocks`, each a whole number of 128-column tile columns and at least `min_cols` wide, so that every
    product still fills whole waves of the persistent GEMM."""
    nsub = min(max_blocks, max(1, nloc // min_cols))
    cuts = [(nloc * i // nsub) // 128 * 128 for i in range(nsub)] + [nloc]
    return [(cuts[i], cuts[i + 1] - cuts[i]) for i in range(nsub) if cuts[i + 1] > cuts[i]]


def allgather_columns(c_full, n_rows: int, n_cols: int, world: int, rank: int, dist):
    """Reassemble a column-major (n_rows x n_cols) matrix stored in the flat torch tensor `c_full` whose own column block
    is already in place.  Works for any backend (NCCL on GPUs, gloo on CPU tensors).  Equal blocks use one in-place
    all_gather_into_tensor; ragged blocks fall back to per-rank broadcasts."""
    blocks = [column_block(n_cols, world, r) for r in range(world)]
    sizes = {b[1] for b in blocks}
    j0, nloc = blocks[rank]
    if len(sizes) == 1:
        mine = c_full[j0 * n_rows:(j0 + nloc) * n_rows]
        dist.all_gather_into_tensor(c_full[: n_rows * n_cols], mine)
    else:
        for r, (b0, bn) in enumerate(blocks):
            if bn:
                dist.broadcast(c_full[b0 * n_rows:(b0 + bn) * n_rows], src=r)
    return c_full


class ColShardTimes:
    """C = A * B at size n with B and C column-sharded over `world` ranks (device-resident, synthetic N(0,1) data)."""

    def __init__(self, n: int, world: int, rank: int, seed: int = 4001):
        import torch
        import torch.distributed as dist

        from . import _lib

        self.torch, self.dist, self.lib = torch, dist, _lib.load()
        self.check = _lib.check
        self.n, self.world, self.rank = n, world, rank
        self.j0, self.nloc = column_block(n, world, rank)
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev)
        g.manual_seed(seed)  # same A on every rank
        self.A = torch.randn(n * n, dtype=torch.float64, device=dev, generator=g)
        g.manual_seed(seed + 1 + rank)
        self.B = torch.randn(n * max(self.nloc, 1), dtype=torch.float64, device=dev, generator=g)
        # C comes from cudaMalloc directly (fm_alloc) so that it can be exported through CUDA IPC
        self.c_ptr = self.lib.fm_alloc(n, n)
        if not self.c_ptr:
            raise _lib.FlomatCudaError("fm_alloc failed: " + _lib.last_error())
        self.peers = []
        self.peer_array = None
        if world > 1:
            handle = ctypes.create_string_buffer(64)
            off = ctypes.c_size_t()
            self.check(self.lib.fm_ipc_export(self.c_ptr, handle, ctypes.byref(off)), "fm_ipc_export")
            mine = (bytes(handle.raw), int(off.value))
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            for r, (h, o) in enumerate(everyone):
                if r == rank:
                    continue
                p = ctypes.c_void_p()
                self.check(self.lib.fm_ipc_open(h, o, ctypes.byref(p)), "fm_ipc_open")
                self.peers.append(p.value)
            self.peer_array = (ctypes.c_void_p * len(self.peers))(*self.peers)
        self._sync_token = torch.zeros(1, device=dev)

    def c_tensor(self):
        """The full C as a flat torch tensor aliasing the fm_alloc'ed buffer (for NCCL and checks)."""
        torch = self.torch

        class _Wrap:
            pass

        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (self.n * self.n,), "typestr": "<f8", "data": (int(self.c_ptr), False), "version": 2}
        return torch.as_tensor(w, device=torch.device("cuda", torch.cuda.current_device()))

    def step_fused(self):
        n = self.n
        self.check(self.lib.fm_dgemm_colshard(n, self.nloc, n, self.A.data_ptr(), n, self.B.data_ptr(), n, self.c_ptr, n, self.j0,
                                              self.peer_array, len(self.peers)), "fm_dgemm_colshard")
        if self.world > 1:
            self.dist.all_reduce(self._sync_token)  # stream-ordered barrier: every peer's stores have landed

    def step_nccl(self):
        n = self.n
        self.check(self.lib.fm_dgemm_colshard(n, self.nloc, n, self.A.data_ptr(), n, self.B.data_ptr(), n, self.c_ptr, n, self.j0,
                                              None, 0), "fm_dgemm_colshard")
        if self.world > 1:
            allgather_columns(self.c_tensor(), n, n, self.world, self.rank, self.dist)

    def _time(self, fn, steps, warmup):
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=self.A.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def checksum(self):
        """Column-block checksums of the assembled C: identical on every rank iff the reassembly is complete."""
        c = self.c_tensor().view(self.n, self.n)  # row r of this view = column r of C
        return c.abs().sum(dim=1)

    def verify(self) -> bool:
        torch, dist = self.torch, self.dist
        cs = self.checksum()
        ok = bool(torch.isfinite(cs).all().item())
        if self.world > 1:
            ref = cs.clone()
            dist.broadcast(ref, src=0)
            ok = ok and bool(torch.equal(ref, cs))
        # own block against an independent product (cuBLAS via torch is only the CHECKER here, never the product)
        if self.nloc:
            n = self.n
            a = self.A.view(n, n).t()  # column-major -> torch row-major transpose
            cols = sorted({0, 1, self.nloc // 2, self.nloc - 1} | {(k * 131) % self.nloc for k in range(28)})
            b = torch.stack([self.B[jj * n:(jj + 1) * n] for jj in cols], dim=1)
            want = a @ b
            c = self.c_tensor()
            got = torch.stack([c[(self.j0 + jj) * n:(self.j0 + jj + 1) * n] for jj in cols], dim=1)
            err = (got - want).abs().sum(dim=0)
            ok = ok and bool((err <= 64 * n * 2.0 ** -52 * want.abs().sum(dim=0)).all().item())
        flag = torch.tensor([1.0 if ok else 0.0], device=self.A.device)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item() == 1.0)

    def bench(self, steps=3, warmup=2) -> dict:
        n = self.n
        flops = 2.0 * n ** 3
        out = {"n": n, "n_gpus": self.world, "scaling": "strong", "columns_per_gpu": self.nloc}
        ms = self._time(self.step_fused, steps, warmup)
        out["fused_ok"] = self.verify()
        out["fused_ms"] = ms
        out["fused_tflops_total"] = flops / (ms * 1e-3) / 1e12
        if self.world > 1:
            self.c_tensor().zero_()
            ms2 = self._time(self.step_nccl, steps, warmup)
            out["nccl_ok"] = self.verify()
            out["nccl_ms"] = ms2
            out["nccl_tflops_total"] = flops / (ms2 * 1e-3) / 1e12
        out["tflops_per_gpu"] = out["fused_tflops_total"] / self.world
        return out

    def close(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        for p in self.peers:
            self.lib.fm_ipc_close(p)
        self.peers = []
        self.lib.fm_free(self.c_ptr)
        self.c_ptr = None


class ShardedJob:
    """BASELINE configs 4 and 5 as ONE job sharded over `world` ranks (the N > 1 workload of bench.py, total work fixed):

      * `times` at size n: A replicated, B and C in column blocks, the all-gather of C fused into the GEMM epilogue (ColShardTimes);
      * `items` independent expm of n_item x n_item matrices, `item_block` of them per rank, no collective.

    step()      device-resident inputs (what `value` times)
    e2e_step()  the same job from pinned HOST buffers: every rank uploads ITS column block of A and of B and its items; A is then
                all-gathered over NVLink (NCCL, in place) instead of travelling world times over PCIe; the rank's block of C and
                its expm results go back to the host beside the kernels that follow them.
    """

    def __init__(self, n: int, items: int, n_item: int, world: int, rank: int, mode: int, seed: int = 4001):
        import torch

        from . import flomat as F

        self.F, self.torch = F, torch
        self.cs = ColShardTimes(n, world, rank, seed=seed)
        self.lib, self.check, self.dist = self.cs.lib, self.cs.check, self.cs.dist
        self.n, self.world, self.rank, self.mode = n, world, rank, mode
        self.n_item = n_item
        self.z0, self.nitems = item_block(items, world, rank)
        self.items_total = items
        dev = self.cs.A.device
        g = torch.Generator(device=dev)
        g.manual_seed(5001 + rank)
        per = n_item * n_item
        self.src = torch.randn(per * max(self.nitems, 1), dtype=torch.float64, device=dev, generator=g) / 16.0
        self.dst = torch.empty_like(self.src)
        self.s_out = np.zeros(max(self.nitems, 1))
        self._host = None

    # ---- device-resident step ------------------------------------------------------------------------------------------------
    def step(self):
        self.cs.step_fused()
        if self.nitems:
            self.F.expm_batched(self.src.data_ptr(), self.n_item, self.nitems, self.dst.data_ptr(), self.mode, self.s_out)

    def flops(self, expm_flops_min) -> float:
        """Algorithmic work of this rank's share: its column block of the product + W_min of its items (needs one step() first)."""
        w = 2.0 * self.n * self.n * self.cs.nloc
        if self.nitems:
            w += float(sum(expm_flops_min(self.n_item, sv) for sv in self.s_out[: self.nitems]))
        return w

    # ---- end to end from host buffers ----------------------------------------------------------------------------------------
    def _host_buffers(self):
        if self._host is None:
            torch, n, nloc, per = self.torch, self.n, max(self.cs.nloc, 1), self.n_item * self.n_item
            pin = lambda count: torch.empty(count, dtype=torch.float64, pin_memory=True)  # noqa: E731
            h = {"a": pin(n * nloc), "b": pin(n * nloc), "c": pin(n * nloc), "x": pin(per * max(self.nitems, 1)), "e": pin(per * max(self.nitems, 1))}
            j0 = self.cs.j0
            h["a"].copy_(self.cs.A[j0 * n:(j0 + nloc) * n])
            h["b"].copy_(self.cs.B[: n * nloc])
            h["x"].copy_(self.src)
            torch.cuda.synchronize()
            self._host = h
        return self._host

    def e2e_bytes(self):
        n, nloc, per = self.n, self.cs.nloc, self.n_item * self.n_item
        return {"h2d": 8 * (2 * n * nloc + per * self.nitems), "d2h": 8 * (n * nloc + per * self.nitems)}

    def e2e_step(self):
        lib, cs, n, nloc, j0 = self.lib, self.cs, self.n, self.cs.nloc, self.cs.j0
        h = self._host_buffers()
        if nloc:
            self.check(lib.fm_upload_async(cs.A.data_ptr() + 8 * j0 * n, n, h["a"].data_ptr(), n, n, nloc), "fm_upload_async")
        if self.world > 1:
            allgather_columns(cs.A, n, n, self.world, self.rank, self.dist)  # A over NVLink instead of `world` times over PCIe
        # the rank's block of B / C in sub-blocks of >= 1024 columns (whole waves of the persistent GEMM): sub-block i+1 goes up and
        # sub-block i-1 comes down while sub-block i is multiplied
        subs = sub_blocks(nloc)
        stage = {}

        def up(i):
            c0, w = subs[i]
            stage[i] = self.F.Flomat.alloc(n, w)  # a recycled block: the allocator orders the upload behind its last readers
            self.check(lib.fm_upload_async(stage[i].a, n, h["b"].data_ptr() + 8 * c0 * n, n, n, w), "fm_upload_async")

        if subs:
            up(0)
        for i, (c0, w) in enumerate(subs):
            if i + 1 < len(subs):
                up(i + 1)
            self.check(lib.fm_dgemm_colshard(n, w, n, cs.A.data_ptr(), n, stage[i].a, n, cs.c_ptr, n, j0 + c0, cs.peer_array, len(cs.peers)),
                       "fm_dgemm_colshard")
            self.check(lib.fm_download_async(h["c"].data_ptr() + 8 * c0 * n, n, cs.c_ptr + 8 * (j0 + c0) * n, n, n, w), "fm_download_async")
        # (the staging blocks are given back when the step returns: fm_free orders the compute stream behind the transfers in flight,
        #  which would serialise the next product with this sub-block's download)
        if self.world > 1:
            self.dist.all_reduce(cs._sync_token)  # stream-ordered barrier: every peer's stores have landed
        per = self.n_item * self.n_item
        if self.nitems:
            x = self.F.Flomat.alloc(per, self.nitems)  # a recycled block: the allocator orders the upload behind its last readers
            self.check(lib.fm_upload_async(x.a, per, h["x"].data_ptr(), per, per, self.nitems), "fm_upload_async")
            # two halves: the results of the first half travel to the host while the second half is computed
            half = self.nitems // 2 if self.nitems >= 128 else self.nitems
            for z0, cnt in ((0, half), (half, self.nitems - half)):
                if cnt:
                    self.F.expm_batched(x.a + 8 * per * z0, self.n_item, cnt, self.dst.data_ptr() + 8 * per * z0, self.mode, None)
                    self.check(lib.fm_download_async(h["e"].data_ptr() + 8 * per * z0, per, self.dst.data_ptr() + 8 * per * z0, per, per, cnt),
                               "fm_download_async")
        self.check(lib.fm_transfer_sync(), "fm_transfer_sync")

    # ---- verification (asserted by bench.py) -----------------------------------------------------------------------------------
    def verify(self, sample_items: int = 16) -> dict:
        """times: every rank's assembled C is complete, identical across ranks and right on sampled columns (ColShardTimes.verify);
        expm: exp(A_z) exp(-A_z) == I within 64 n eps ||.||_1 ||.||_1 for a sample of this rank's items (all on the device)."""
        torch, F = self.torch, self.F
        out = {"times_ok": self.cs.verify(), "expm_ok": True}
        k = min(sample_items, self.nitems)
        if k:
            ni, per = self.n_item, self.n_item * self.n_item
            neg = (-self.src[: per * k]).contiguous()
            e_pos, e_neg = torch.empty_like(neg), torch.empty_like(neg)
            F.expm_batched(self.src.data_ptr(), ni, k, e_pos.data_ptr(), self.mode, None)
            F.expm_batched(neg.data_ptr(), ni, k, e_neg.data_ptr(), self.mode, None)
            p = torch.bmm(e_pos.view(k, ni, ni), e_neg.view(k, ni, ni))  # (row-major views = transposes: E_neg^T E_pos^T)^T, still I)
            eye = torch.eye(ni, dtype=torch.float64, device=p.device)
            err = (p - eye).abs().sum(dim=1).amax(dim=1)
            bound = 64 * ni * 2.0 ** -52 * e_pos.view(k, ni, ni).abs().sum(dim=2).amax(dim=1) * e_neg.view(k, ni, ni).abs().sum(dim=2).amax(dim=1)
            # the timed run's results for these items (a different batch size picks other tile shapes: same values within rounding)
            d = (e_pos - self.dst[: per * k]).view(k, ni, ni).abs().sum(dim=2).amax(dim=1)
            same = bool((d <= 64 * ni * 2.0 ** -52 * e_pos.view(k, ni, ni).abs().sum(dim=2).amax(dim=1)).all().item())
            out["expm_ok"] = bool((err <= bound).all().item()) and bool(torch.isfinite(p).all().item()) and same
        flag = torch.tensor([1.0 if (out["times_ok"] and out["expm_ok"]) else 0.0], device=self.cs.A.device)
        if self.world > 1:
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        out["all_ranks_ok"] = bool(flag.item() == 1.0)
        return out

    def close(self):
        self.cs.close()


def bind_host_near_device(device: int) -> dict:
    """Pin the calling thread (and every thread it creates later: the library's copy pool, torch's helpers) to the cores of the NUMA
    node the GPU hangs off, so that pinned host buffers are first-touched there and the staging copies do not cross the socket link.
    One process per GPU only (a single process that drives all devices must stay unbound).  Best effort: returns what it found and
    did; a box that does not expose the topology (numa_node = -1, as most virtual machines) is left alone."""
    import os

    info = {"device": device, "bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else device
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        node = int(open(path + "/numa_node").read().strip())
        cpulist = open(path + "/local_cpulist").read().strip()
        info.update({"pci": bus, "numa_node": node, "local_cpulist": cpulist})
        if node < 0 or not cpulist:
            return info
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
            info["cpus"] = len(cpus)
    except Exception as e:  # noqa: BLE001  (no NVML, no sysfs, no permission: stay unbound)
        info["error"] = f"{type(e).__name__}: {e}"
    return info
