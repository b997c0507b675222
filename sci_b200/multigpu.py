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
        out["pct_of_fp64_peak_per_gpu"] = 100.0 * out["tflops_per_gpu"] / 37.0
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
