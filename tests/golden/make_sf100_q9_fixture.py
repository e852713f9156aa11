"""Generates tests/golden/oracle_sf100_q9.txt: the CPU oracle's TPC-H Q9 at the headline scale (SF100) in the
reference's result-file format.  CPU only; lineitem/orders are generated and folded in order-range chunks
(memory stays below ~6 GB: partsupp is 80 M rows).  About 8 minutes.

    python tests/golden/make_sf100_q9_fixture.py [sf]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O   # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
tag = ("%g" % sf).replace(".", "p")
here = os.path.dirname(os.path.abspath(__file__))
t0 = time.time()
part, supplier, partsupp = O.gen_part(sf, "pink"), O.gen_supplier(sf), O.gen_partsupp(sf)
part["p_name"] = None          # 20 M python objects are not needed once the LIKE flags exist
print("dimension tables %.0f s" % (time.time() - t0), flush=True)
names = O.nation_names()
n = O.lib().tg_num_orders(sf)
step = 5_000_000
tot = {}
for lo in range(0, n, step):
    hi = min(n, lo + step)
    orders, line = O.gen_orders_lineitem(sf, lo, hi, lineitem_cols=["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount"],
                                         orders_cols=["o_orderkey", "o_orderdate"])
    for nation, year, v in O.q9(part, supplier, partsupp, orders, line):
        tot[(nation, year)] = tot.get((nation, year), 0) + v
    print("orders [%d, %d): %.0f s" % (lo, hi, time.time() - t0), flush=True)
rows = sorted(((k[0], k[1], v) for k, v in tot.items()), key=lambda r: (r[0], -r[1]))
txt = O.q9_text(rows)
open(os.path.join(here, "oracle_sf%s_q9.txt" % tag), "w").write(txt)
print(txt)
