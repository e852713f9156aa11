"""Generates tests/golden/oracle_sf100_q18.txt: the CPU oracle's TPC-H Q18 at the headline scale (SF100),
in the reference's result-file format.  CPU only; order ranges are generated and folded chunk by chunk
(orders never straddle a chunk), so memory stays below ~3 GB.  About 6 minutes.

    python tests/golden/make_sf100_q18_fixture.py [sf]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O   # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
tag = ("%g" % sf).replace(".", "p")
here = os.path.dirname(os.path.abspath(__file__))
t0 = time.time()
n = O.lib().tg_num_orders(sf)
step = 5_000_000
rows = []
for lo in range(0, n, step):
    hi = min(n, lo + step)
    orders, line = O.gen_orders_lineitem(sf, lo, hi, lineitem_cols=["l_orderkey", "l_quantity"],
                                         orders_cols=["o_orderkey", "o_custkey", "o_orderdate", "o_totalprice"])
    # the customer side of the join: every o_custkey of dbgen exists (1..150000*sf); q18() checks membership
    ck = np.unique(orders["o_custkey"])
    rows += O.q18({"c_custkey": ck}, orders, line, limit=None)
    print("orders [%d, %d): %d qualifying so far, %.0f s" % (lo, hi, len(rows), time.time() - t0), flush=True)
ncust = O.lib().tg_num_customers(sf)
assert all(1 <= r["c_custkey"] <= ncust for r in rows)
rows.sort(key=lambda r: (-r["o_totalprice"], r["o_orderdate"]))
txt = O.q18_text(rows[:100])
open(os.path.join(here, "oracle_sf%s_q18.txt" % tag), "w").write(txt)
print(txt)
