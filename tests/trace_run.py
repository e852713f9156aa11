"""Diagnostic (not a test): per-phase breakdown of Q6/Q1/Q3 at a given SF with PG_TRACE=1."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import time
    import torch
    import torch.distributed as dist
    from plan_b200 import _lib as L, compute as X, dist as D, tpch as T
    sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    queries = sys.argv[2].split(",") if len(sys.argv) > 2 else ["q6", "q1", "q3"]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    lib = L.lib()
    L.check(lib.pg_init(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        D.init_comm(lib, L.check)
    n = lib.pg_tpch_num_orders(sf)
    lo, hi = D.shard_range(n, rank, world)
    tables = T.generate_device_tables(sf, lo, hi, want=T.ALL_TABLES if "q9" in queries else ("lineitem", "orders", "customer"))
    for name in ("customer", "part", "supplier", "partsupp", "nation"):
        if name in tables:
            tables[name].set_replicated()
    plans = {"q6": T.q6_plan, "q1": T.q1_plan, "q3": T.q3_plan, "q3k": lambda: T.q3_topk_plan(10), "q18": T.q18_plan, "q9": T.q9_plan,
             "gok": lambda: T.groupby_plan(key="l_orderkey", value="l_quantity", having_gt=314, topk=100),
             "gpk": lambda: T.groupby_plan(key="l_partkey", value="l_quantity", topk=100),
             "gsk": lambda: T.groupby_plan(key="l_suppkey", value="l_extendedprice", topk=100)}
    for q in queries:
        ex = X.gpuPipelineExec(plans[q](), tables)
        ex.Init()
        for it in range(3):
            ex.Reset()
            if world > 1:
                dist.barrier()
            os.environ["PG_TRACE"] = "1" if it == 2 else "0"
            t0 = time.perf_counter()
            chunks = X.drain(ex)
            dt = (time.perf_counter() - t0) * 1e3
            if it == 2:
                print("[r%d] %s: drain %.3f ms exec %.3f ms kernels %.3f ms main %.3f ms rows_out %d | %s" % (
                    rank, q, dt, ex.stats.exec_ms, ex.stats.kernel_ms, ex.stats.main_kernel_ms,
                    sum(c.Card() for c in chunks), ex.Explain()[:100]), flush=True)
        os.environ["PG_TRACE"] = "0"
        ex.Close()


if __name__ == "__main__":
    main()
