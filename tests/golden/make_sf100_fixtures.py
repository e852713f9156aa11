"""Generates tests/golden/oracle_sf100_q{6,1,3}.txt: the CPU oracle's results at the headline scale
(SF100, 600,037,902 lineitem rows) in the reference's result-file format.  CPU only (about 15 minutes,
~35 GB of RAM); the GPU tests then compare the CUDA path with these files at full size, which pins
the order-dependent >19-digit rounding of Q1's sum_charge against the sequential fold.

    python tests/golden/make_sf100_fixtures.py [sf]
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
orders, line = O.gen_orders_lineitem(
    sf, lineitem_cols=["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag",
                       "l_linestatus", "l_shipdate"],
    orders_cols=["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"])
cust = O.gen_customer(sf)
print("generated %d lineitem rows in %.0f s" % (len(line["l_orderkey"]), time.time() - t0), flush=True)
for name, fn in (("q6", lambda: O.q6_text(O.q6(line))),
                 ("q1", lambda: O.q1_text(O.q1(line))),
                 ("q3", lambda: O.q3_text(O.q3(cust, orders, line, capacity=4000000)))):
    t = time.time()
    txt = fn()
    open(os.path.join(here, "oracle_sf%s_%s.txt" % (tag, name)), "w").write(txt)
    print(name, "%.0f s" % (time.time() - t), flush=True)
    print(txt, flush=True)
