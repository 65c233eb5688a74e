"""Host-side logic of the N > 1 path on CPU: world_size-2 and -3 `gloo` groups exercise the plumbing that carries the
CUDA IPC handles between ranks (gather_handles) and the row partition every rank derives independently."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import svgf, ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nbytes, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import svgf as _svgf
    m = _svgf()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = np.full(nbytes, rank + 1, np.uint8); mine[0] = 100 + rank
    allh = m.gather_handles(mine, dist, world)
    rs = m.row_partition(1080, world)
    q.put((rank, allh.tolist(), rs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_handle_gather_is_rank_ordered(world):
    m = svgf()
    nbytes = m.lib().svgf_ipc_handles_size()
    assert nbytes == 17 * 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nbytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    for rank, allh, rs in res:
        a = np.array(allh, np.uint8).reshape(world, nbytes)
        for r in range(world):
            assert a[r, 0] == 100 + r and (a[r, 1:] == r + 1).all(), "rank %d sees rank %d's handles in the wrong slot" % (rank, r)
        assert rs == m.row_partition(1080, world)


@pytest.mark.parametrize("H,world", [(1080, 1), (1080, 2), (1080, 8), (2160, 8), (7, 8), (1, 3)])
def test_row_partition_covers_the_frame_once(H, world):
    m = svgf()
    rs = m.row_partition(H, world)
    assert len(rs) == world + 1 and rs[0] == 0 and rs[-1] == H
    assert all(rs[i] <= rs[i + 1] for i in range(world))
    sizes = [rs[i + 1] - rs[i] for i in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_balanced_partition_equalises_cost():
    m = svgf()
    rs = m.row_partition(2160, 8)
    # a cost profile that grows down the frame (more geometry in the lower rows)
    cost = [1.0, 1.1, 1.4, 1.8, 2.2, 2.4, 2.0, 1.5]
    new = m.balanced_partition(rs, cost)
    assert new[0] == 0 and new[-1] == 2160 and all(new[i] < new[i + 1] for i in range(8))
    dens = np.repeat(np.array(cost) / 270.0, 270)
    per = [dens[new[r]:new[r + 1]].sum() for r in range(8)]
    assert max(per) / min(per) < 1.03
    # uniform cost keeps equal strips
    assert m.balanced_partition(rs, [1.0] * 8) == rs
