"""Multi-GPU plumbing on the host side: one process per GPU, torch.distributed carries the
128-byte NCCL id and the barriers; the data-path collectives run inside libplangpu (NCCL).

Sharding: orders and lineitem are row-range sharded by ORDER INDEX, so a rank
holds whole orders with all their lineitems (co-partitioned on the order key), and shards
are contiguous in rank order, which keeps the order-dependent decimal rounding well defined.
"""
import ctypes as C
import os


def shard_range(n, rank, world):
    """[lo, hi) of `n` items for `rank` -- contiguous, disjoint, covering, balanced to +-1."""
    return n * rank // world, n * (rank + 1) // world


def init_comm(lib, check, backend_device="cuda"):
    """Bring up the library's NCCL communicator using torch.distributed for the id exchange.
    Requires torch.distributed to be initialised.  Returns (world, rank)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return 1, 0
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        check(lib.pg_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    if backend_device == "cuda":
        uid = uid.cuda()
    dist.broadcast(uid, 0)
    raw = bytes(uid.cpu().tolist())
    check(lib.pg_comm_init(world, rank, raw))
    return world, rank


def broadcast_bytes(data, nbytes, src=0, device="cpu"):
    """Broadcast a small byte string (the NCCL id) over the initialised process group."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t = torch.tensor(list(data), dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def merge_lowcard_partials(per_rank):
    """Rank-ordered exact merge of per-rank low-cardinality partial aggregates (restates the
    C++ merge of scanagg.cu for host-side tests): per_rank[r] = {group key: {"count": n,
    "sums": [ints], "first_row": i}}.  Sums are Python ints (exact); first_row takes the min."""
    out = {}
    for part in per_rank:
        for key, g in part.items():
            o = out.setdefault(key, {"count": 0, "sums": [0] * len(g["sums"]), "first_row": None})
            o["count"] += g["count"]
            o["sums"] = [a + b for a, b in zip(o["sums"], g["sums"])]
            if g["first_row"] is not None and (o["first_row"] is None or g["first_row"] < o["first_row"]):
                o["first_row"] = g["first_row"]
    return out
