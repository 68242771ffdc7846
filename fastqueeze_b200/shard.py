"""Sharding blocks over the GPUs of one box.

Blocks are independent units (every model resets per block), so there is no data-path collective:
each rank encodes a contiguous range of block indices.  The only exchange is host-side and tiny:
record lengths are all-gathered so that every rank knows the file offset of its records, which is
what keeps the layout identical to the reference at -t 1 (first record at offset 16, each next at
prev.offset + prev.len, addBlockPEInfo @0x412f6d-0x412f76).
"""


def block_range(nblocks: int, world: int, rank: int):
    """Contiguous, balanced range [lo, hi) of block indices for `rank` (first ranks take the remainder)."""
    base, rem = divmod(nblocks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ordered_offsets(local_lens, nblocks: int, first_offset: int = 16, group=None):
    """All-gather the record lengths of every rank and return (global_lens, offsets) in block order.

    local_lens: lengths of this rank's records, in its block_range order.  Works on any
    torch.distributed backend (gloo on CPU, nccl on GPU); with no process group it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        lens = list(local_lens)
    else:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        width = max(block_range(nblocks, world, r)[1] - block_range(nblocks, world, r)[0] for r in range(world))
        mine = torch.zeros(width, dtype=torch.int64, device=dev)
        mine[:len(local_lens)] = torch.tensor(list(local_lens), dtype=torch.int64, device=dev)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        lens = []
        for r in range(world):
            lo, hi = block_range(nblocks, world, r)
            lens += [int(x) for x in parts[r][:hi - lo].cpu()]
    offs, o = [], first_offset
    for ln in lens:
        offs.append(o)
        o += ln
    return lens, offs
