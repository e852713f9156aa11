"""Kernel tuning harness (a script, not a test): times the scan-aggregate pipelines of Q6 / Q1 (and Q3) under
environment-selected kernel variants, data resident in HBM.
    python tests/tune_scan.py <sf> q6,q1 "PG_LC_UNROLL=2" "PG_LC_UNROLL=4" ...
Each further argument is a space-separated list of VAR=value settings applied for one measurement."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    sf = float(sys.argv[1])
    queries = sys.argv[2].split(",")
    configs = sys.argv[3:] or [""]
    from plan_b200 import _lib as L, compute as X, tpch as T
    lib = L.lib()
    L.check(lib.pg_init(0))
    want = ["lineitem"] + (["orders", "customer"] if "q3" in queries else [])
    tables = T.generate_device_tables(sf, want=tuple(want))
    plans = {"q6": T.q6_plan, "q1": T.q1_plan, "q3": lambda: T.q3_topk_plan(10)}
    for cfg in configs:
        saved = {}
        for kv in cfg.split():
            k, v = kv.split("=", 1)
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        for q in queries:
            ex = X.gpuPipelineExec(plans[q](), tables)
            ex.Init()
            for _ in range(3):
                ex.Reset(); X.drain(ex)
            main_ms, kern_ms, exec_ms = [], [], []
            for _ in range(10):
                ex.Reset(); X.drain(ex)
                main_ms.append(ex.stats.main_kernel_ms); kern_ms.append(ex.stats.kernel_ms); exec_ms.append(ex.stats.exec_ms)
            gb = ex.stats.main_kernel_bytes / 1e9
            m = sorted(main_ms)[len(main_ms) // 2]
            print("[%s] %s main %.3f ms (%.2f GB -> %.0f GB/s) kernels %.3f exec %.3f | %s" % (
                cfg, q, m, gb, gb / m * 1e3, sorted(kern_ms)[5], sorted(exec_ms)[5], ex.Explain()[:230]), flush=True)
            ex.Close()
        for k, v in saved.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


if __name__ == "__main__":
    main()
