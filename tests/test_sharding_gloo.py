"""CPU, world_size=2 over gloo: the host-side sharding logic of the multi-GPU paths (column blocks of `times`, item blocks
of batched expm, reassembly of C).  The per-rank product is done by the ORACLE here -- this test covers the plumbing
around the kernel, not the kernel."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))


def test_partitions_cover_exactly_once():
    from sci_b200.multigpu import column_block, item_block

    for n in (1, 100, 128, 129, 1000, 8192, 16384, 16400):
        for world in (1, 2, 3, 4, 8):
            blocks = [column_block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (a0, an), (b0, bn) in zip(blocks, blocks[1:]):
                assert a0 + an == b0
            assert blocks[-1][0] + blocks[-1][1] == n
            assert all(b[0] % 128 == 0 for b in blocks if b[1])
    for batch in (1, 7, 4096):
        for world in (1, 2, 8):
            blocks = [item_block(batch, world, r) for r in range(world)]
            assert sum(b[1] for b in blocks) == batch and blocks[0][0] == 0
            assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1
    assert [item_block(4096, 8, r)[1] for r in range(8)] == [512] * 8  # BASELINE config 5


def test_single_process_partitions_match_the_launcher_ones():
    """fm_mg_dgemm / fm_mg_expm_batched (one process, all devices; csrc/mg.cu) cut B / C and the batch exactly like the
    one-process-per-GPU path does (sci_b200/multigpu.py): host arithmetic, no GPU needed."""
    from ctypes import byref, c_int

    from sci_b200 import _lib
    from sci_b200.multigpu import column_block, item_block

    lib = _lib.load()
    a, b = c_int(), c_int()
    for n in (1, 100, 128, 129, 1000, 2500, 8192, 16384, 16400):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                assert lib.fm_mg_column_block(n, world, r, byref(a), byref(b)) == 0
                assert (a.value, b.value) == column_block(n, world, r), (n, world, r)
                assert lib.fm_mg_item_block(n, world, r, byref(a), byref(b)) == 0
                assert (a.value, b.value) == item_block(n, world, r), (n, world, r)
    assert lib.fm_mg_column_block(100, 4, 4, byref(a), byref(b)) == -1002 and lib.fm_mg_item_block(-1, 2, 0, byref(a), byref(b)) == -1002
    assert lib.fm_mg_device_count() == 0  # no group without fm_mg_init (and no GPU is touched by asking)


def test_e2e_sub_blocks_cover_the_column_block():
    from sci_b200.multigpu import sub_blocks

    for nloc in (0, 1, 127, 128, 1000, 1024, 2048, 2176, 4096, 8192, 16384):
        subs = sub_blocks(nloc)
        assert sum(w for _, w in subs) == nloc and len(subs) <= 4
        pos = 0
        for c0, w in subs:
            assert c0 == pos and c0 % 128 == 0 and w > 0
            pos += w
        if nloc >= 2048:
            assert min(w for _, w in subs) >= 1024  # whole waves of 128 x 128 tiles at m = 16384
    assert sub_blocks(2048) == [(0, 1024), (1024, 1024)] and sub_blocks(8192) == [(0, 2048), (2048, 2048), (4096, 2048), (6144, 2048)]


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from sci_b200.multigpu import column_block, allgather_columns
    from oracle.flomat_oracle import Flomat, CProvider
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ref = Flomat(CProvider(threads=1))
    for n in (256, 300):            # equal blocks (in-place all_gather) and ragged blocks (broadcasts)
        rng = np.random.default_rng(4001)
        a = np.asfortranarray(rng.standard_normal((n, n)))
        b = np.asfortranarray(rng.standard_normal((n, n)))
        j0, nloc = column_block(n, world, rank)
        c = torch.zeros(n * n, dtype=torch.float64)
        if nloc:
            blk = ref.gemm(a, np.asfortranarray(b[:, j0:j0 + nloc]))      # this rank's column block of C
            c[j0 * n:(j0 + nloc) * n] = torch.from_numpy(blk.ravel(order="F").copy())
        allgather_columns(c, n, n, world, rank, dist)
        full = ref.gemm(a, b)
        got = c.numpy().reshape(n, n, order="F")
        assert np.array_equal(got, full), (rank, n)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_column_sharded_times_world2_gloo(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
