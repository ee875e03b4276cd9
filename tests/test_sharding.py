"""CPU, world_size 2, gloo: the host-side sharding of batched solves (source partition + final row gather).
Each rank 'solves' its shard with the CPU oracle; the gathered matrix must equal the serial one row for row."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gproshan_b200 import meshgen as mg
from gproshan_b200.sharding import gather_rows, gather_rows_ragged, shard_bounds, shard_sources


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 128, 1024, 1021):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_weak_scaling_shards_are_prefixes_of_the_1024_job():
    src = mg.random_sources(1024, 1024, 1998092, unique=True)
    for w in (1, 2, 4, 8):
        got = np.concatenate([shard_sources(src, r, w, per_rank=128) for r in range(w)])
        assert np.array_equal(got, src[:128 * w])
    with pytest.raises(ValueError):
        shard_sources(src[:100], 0, 2, per_rank=128)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_src, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_lib import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    mesh = mg.icosphere(6, 5e-3, seed=4).astype(np.float32)
    srcs = mg.random_sources(11, n_src, mesh.n_vertices, unique=True)

    def solve(s):
        t, srt, lim = orc.compute_toplesets(mesh, [s])
        return orc.ptp_cpu(mesh, [s], lim, srt)[0]

    mine = shard_sources(srcs, rank, world)
    rows = torch.from_numpy(np.stack([solve(s) for s in mine]) if mine.size else np.zeros((0, mesh.n_vertices), np.float32))
    if srcs.size % world == 0:
        full = gather_rows(rows, world)
    else:
        full = gather_rows_ragged(rows, srcs.size, rank, world)
    if rank == 0:
        want = np.stack([solve(s) for s in srcs])
        q.put(bool(np.array_equal(full.numpy(), want)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_src", [8, 7])
def test_two_rank_gather_equals_serial(n_src):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_src, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
