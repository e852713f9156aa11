"""Launched by tests/test_gpu_multi.py: ONE process drives N GPUs (pg_init_devices / pg_use_device, include/plangpu.h),
one host thread per device -- the shape the single-process reference needs (the Go shim runs one goroutine per device
locked to its OS thread).  Every thread builds its row-range shard in HBM and runs Q6 / Q1 / Q3 top-10 / Q9 (partsupp
sharded: the all-to-all row exchange) through the C ABI; the NCCL collectives inside the plans meet across the threads.
Every thread's merged result is compared with the CPU oracle over the WHOLE table.  Exit code 0 = parity everywhere."""
import ctypes as C
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from oracle import oracle as O
    from plan_b200 import _lib as L, compute as X, dist as D, tpch as T
    import test_gpu_scanagg as SA
    import test_gpu_join as J
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    # the C oracle keeps static scratch: its calls are serialised (only the CPU reference, never the GPU work, so the
    # threads still meet inside the plans' collectives)
    import types
    olock = threading.RLock()       # oracle functions call each other

    def locked(fn):
        def call(*a, **kw):
            with olock:
                return fn(*a, **kw)
        return call
    for name, fn in list(vars(O).items()):
        if isinstance(fn, types.FunctionType) and not name.startswith("_") and name not in ("lib", "build"):
            setattr(O, name, locked(fn))
    sf = 0.05
    lib = L.lib()
    devs = (C.c_int * world)(*range(world))
    L.check(lib.pg_init_devices(world, devs))
    assert lib.pg_num_devices() == world
    orders, line = O.gen_orders_lineitem(sf)
    host = {"orders": orders, "lineitem": line, "customer": O.gen_customer(sf)}
    q9_want = O.q9(O.gen_part(sf, "pink"), O.gen_supplier(sf), O.gen_partsupp(sf), orders, line, like_word="pink")
    n_orders = lib.pg_tpch_num_orders(sf)
    errors = []

    def worker(rank):
        try:
            L.check(lib.pg_use_device(rank))
            lo, hi = D.shard_range(n_orders, rank, world)
            tables = T.generate_device_tables(sf, lo, hi, want=T.ALL_TABLES, partsupp_shard=(rank, world) if world > 1 else None)
            for name in ("customer", "part", "supplier", "nation") + (("partsupp",) if world == 1 else ()):
                tables[name].set_replicated()
            SA.check_q6(O, tables, line)
            SA.check_q1(O, tables, line)
            J.check_q3_topk(O, tables, host, 10)
            chunks, st, explain = J._run(T.q9_plan("pink"), tables)
            assert ("ROW EXCHANGE" in explain) == (world > 1), explain
            assert J._q9_rows(chunks) == q9_want
            for t in tables.values():
                t.free()
        except BaseException as e:      # noqa: BLE001 -- reported by the main thread
            import traceback
            errors.append((rank, traceback.format_exc()))
            raise e

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    if errors or any(t.is_alive() for t in threads):
        for r, tb in errors:
            sys.stderr.write("device %d:\n%s\n" % (r, tb))
        os._exit(1)          # a failed rank leaves its peers inside a collective: do not wait for them
    L.check(lib.pg_shutdown())
    print("single process, %d device(s): parity OK" % world)


if __name__ == "__main__":
    main()
