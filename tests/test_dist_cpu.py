"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- shard ranges, id
broadcast, rank-ordered exact merge of partial aggregates (checked against the oracle on the
whole table)."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    from oracle import oracle as O
    from plan_b200 import dist as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # id broadcast (the 128 bytes of ncclUniqueId travel the same way)
        blob = bytes(range(128)) if rank == 0 else None
        assert D.broadcast_bytes(blob, 128) == bytes(range(128))
        sf = 0.02
        n_orders = O.lib().tg_num_orders(sf)
        lo, hi = D.shard_range(n_orders, rank, world)
        _, line = O.gen_orders_lineitem(sf, lo, hi)           # this rank's shard only
        first = O.lib().tg_count_lineitems(sf, 0, lo)
        res = O.q1(line)
        mine = {(g["l_returnflag"], g["l_linestatus"]): {
            "count": g["count_order"], "first_row": first + g["first_row"],
            "sums": [g["x_qty"], g["x_base"], g["x_disc_price"], g["x_charge"], g["x_disc"]]} for g in res["groups"]}
        r6 = O.q6(line)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, r6["exact"], r6["rows_selected"], len(line["l_orderkey"])))
        merged = D.merge_lowcard_partials([g[0] for g in gathered])
        _, whole = O.gen_orders_lineitem(sf)
        ref = O.q1(whole)
        # the same merge through the PRODUCT's own code (pg_host_merge_partials: the function the pipelines call after
        # their all-gather; host arithmetic only, no GPU): per-rank records in the library's wire layout
        import ctypes as C
        from plan_b200 import _lib as L
        keys = sorted({k for g in gathered for k in g[0]})
        nvals, ng = len(keys) * 6, len(keys)
        rank_bytes = nvals * 16 + ng * 8
        blob = bytearray()
        for part, _, _, _ in gathered:
            tot, first = [], []
            for k in keys:
                p = part.get(k)
                vals = [0] * 6 if p is None else [p["count"]] + list(p["sums"])
                tot += vals
                first.append(0x7f7f7f7f7f7f7f7f if p is None else p["first_row"])
            for v in tot:
                blob += int(v & ((1 << 128) - 1)).to_bytes(16, "little")
            for f in first:
                blob += int(f).to_bytes(8, "little", signed=True)
        out_tot = (C.c_uint64 * (2 * nvals))()
        out_first = (C.c_int64 * ng)()
        buf = (C.c_char * len(blob)).from_buffer(blob)
        L.check(L.lib().pg_host_merge_partials(C.addressof(buf), rank_bytes, world, nvals, ng, C.addressof(out_tot), C.addressof(out_first)))
        for i, k in enumerate(keys):
            got = [out_tot[2 * (6 * i + j)] | (out_tot[2 * (6 * i + j) + 1] << 64) for j in range(6)]
            assert got == [merged[k]["count"]] + merged[k]["sums"], (k, got)
            assert out_first[i] == merged[k]["first_row"]
        assert sum(g[3] for g in gathered) == len(whole["l_orderkey"])
        assert len(merged) == len(ref["groups"])
        for g in ref["groups"]:
            m = merged[(g["l_returnflag"], g["l_linestatus"])]
            assert m["count"] == g["count_order"] and m["first_row"] == g["first_row"]
            assert m["sums"] == [g["x_qty"], g["x_base"], g["x_disc_price"], g["x_charge"], g["x_disc"]]
        ref6 = O.q6(whole)
        assert sum(g[1] for g in gathered) == ref6["exact"] and sum(g[2] for g in gathered) == ref6["rows_selected"]
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_exactly():
    sys.path.insert(0, ROOT)
    from plan_b200 import dist as D
    for n in (0, 1, 7, 1500000, 150000001):
        for world in (1, 2, 3, 4, 8):
            rs = [D.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def test_two_rank_merge_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29611
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
