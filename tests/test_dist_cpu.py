"""CPU: the N > 1 host logic (block ranges, ordered gather of record lengths) on gloo, world_size 2."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from fastqueeze_b200.shard import block_range, ordered_offsets


def test_block_range_partitions():
    for n in (0, 1, 7, 8, 274, 275):
        for world in (1, 2, 4, 8):
            spans = [block_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, nblocks, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = block_range(nblocks, world, rank)
    local = [1000 + 7 * i for i in range(lo, hi)]          # stand-in record lengths of this rank's blocks
    lens, offs = ordered_offsets(local, nblocks)
    q.put((rank, lens, offs))
    dist.destroy_process_group()


@pytest.mark.parametrize("nblocks", [5, 8])
def test_ordered_gather_world2(nblocks):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, nblocks, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [1000 + 7 * i for i in range(nblocks)]
    offs, o = [], 16
    for ln in want:
        offs.append(o)
        o += ln
    for _, lens, got in res:
        assert lens == want and got == offs


def test_identity_without_process_group():
    lens, offs = ordered_offsets([10, 20, 30], 3)
    assert lens == [10, 20, 30] and offs == [16, 26, 46]
