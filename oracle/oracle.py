"""ctypes driver for the CPU oracle (liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Restates, for result comparison, the reference's output path:
  Vector.GetValue   /root/reference/pkg/chunk/vector.go:76-186
  Value.String      /root/reference/pkg/chunk/value.go:26-70
  Chunk.SaveToFile  /root/reference/pkg/chunk/chunk.go:196-220 (tab separated rows)
  ORDER BY keys     /root/reference/pkg/compute/sort_encoder.go:65-81
  LIMIT             /root/reference/pkg/compute/executor_limit.go:120-137
"""
import ctypes as C
import datetime
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CMP_EQ, CMP_NE, CMP_LT, CMP_LE, CMP_GT, CMP_GE = range(6)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


class Dec(C.Structure):
    _fields_ = [("coef", C.c_uint64), ("scale", C.c_int8), ("neg", C.c_uint8)]

    def tuple(self):
        return (int(self.coef), int(self.scale), int(self.neg))


class Huge(C.Structure):
    _fields_ = [("lower", C.c_uint64), ("upper", C.c_int64)]

    def value(self):
        return (int(self.upper) << 64) + int(self.lower)


class Q6Result(C.Structure):
    _fields_ = [("rows_in", C.c_int64), ("rows_selected", C.c_int64), ("sum", Dec),
                ("has_row", C.c_int), ("exact_lo", C.c_uint64), ("exact_hi", C.c_int64),
                ("error", C.c_int)]


class Q1Group(C.Structure):
    _fields_ = [("rf", C.c_uint8), ("ls", C.c_uint8), ("sum_qty", Huge),
                ("sum_base", Dec), ("sum_disc_price", Dec), ("sum_charge", Dec),
                ("avg_qty_sum", C.c_double), ("avg_price_sum", Dec), ("avg_disc_sum", Dec),
                ("count", C.c_uint64), ("avg_qty", C.c_double), ("avg_price", Dec), ("avg_disc", Dec),
                ("x_base_lo", C.c_uint64), ("x_disc_price_lo", C.c_uint64), ("x_charge_lo", C.c_uint64),
                ("x_disc_lo", C.c_uint64), ("x_base_hi", C.c_int64), ("x_disc_price_hi", C.c_int64),
                ("x_charge_hi", C.c_int64), ("x_disc_hi", C.c_int64), ("x_qty", C.c_int64),
                ("first_row", C.c_int64)]


class Q1Result(C.Structure):
    _fields_ = [("rows_in", C.c_int64), ("rows_selected", C.c_int64), ("ngroups", C.c_int),
                ("error", C.c_int), ("g", Q1Group * 64)]


class Q3Group(C.Structure):
    _fields_ = [("orderkey", C.c_int64), ("orderdate", C.c_int32), ("shippriority", C.c_int32),
                ("revenue", Dec), ("x_rev_lo", C.c_uint64), ("x_rev_hi", C.c_int64),
                ("first_row", C.c_int64)]


class Q3Result(C.Structure):
    _fields_ = [("n_cust_sel", C.c_int64), ("n_orders_sel", C.c_int64), ("n_orders_joined", C.c_int64),
                ("n_line_sel", C.c_int64), ("n_line_joined", C.c_int64), ("ngroups", C.c_int64),
                ("nout", C.c_int64), ("error", C.c_int)]


class StatsGroup(C.Structure):
    _fields_ = [("rf", C.c_uint8), ("min_ext", Dec), ("max_ext", Dec), ("max_disc", Dec), ("sum_tax", Dec),
                ("avg_tax", Dec), ("sum_taxed", Dec), ("count", C.c_uint64), ("first_row", C.c_int64),
                ("n_ext", C.c_uint64), ("n_tax", C.c_uint64), ("n_taxed", C.c_uint64)]


class StatsResult(C.Structure):
    _fields_ = [("rows_in", C.c_int64), ("rows_selected", C.c_int64), ("ngroups", C.c_int), ("error", C.c_int),
                ("g", StatsGroup * 16)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.tg_num_orders.restype = C.c_int64
        L.tg_num_orders.argtypes = [C.c_double]
        L.tg_num_customers.restype = C.c_int64
        L.tg_num_customers.argtypes = [C.c_double]
        L.tg_count_lineitems.restype = C.c_int64
        L.tg_count_lineitems.argtypes = [C.c_double, C.c_int64, C.c_int64]
        L.tg_gen_orders_lineitem.restype = C.c_int64
        L.tg_gen_orders_lineitem.argtypes = [C.c_double, C.c_int64, C.c_int64] + [C.c_void_p] * 19
        L.tg_gen_customer.restype = None
        L.tg_gen_customer.argtypes = [C.c_double, C.c_int64, C.c_int64] + [C.c_void_p] * 3
        L.tg_segment_name.restype = C.c_char_p
        L.tg_segment_name.argtypes = [C.c_int]
        L.orc_q6.restype = None
        L.orc_q6.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_double, C.c_double,
                                                             C.c_int32, C.POINTER(Q6Result)]
        L.orc_q1.restype = None
        L.orc_q1.argtypes = [C.c_int64] + [C.c_void_p] * 7 + [C.c_int32, C.POINTER(Q1Result)]
        L.orc_q3.restype = None
        L.orc_q3.argtypes = ([C.c_int64, C.c_void_p, C.c_void_p, C.c_uint8] +
                             [C.c_int64] + [C.c_void_p] * 4 + [C.c_int32] +
                             [C.c_int64] + [C.c_void_p] * 4 + [C.c_int32] +
                             [C.c_void_p, C.c_int64, C.POINTER(Q3Result)])
        L.orc_stats.restype = None
        L.orc_stats.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_int32] * 6 + [C.c_int64] + [C.c_void_p] * 6 + [C.POINTER(StatsResult)]
        assert L.orc_sizeof_stats_result() == C.sizeof(StatsResult)
        L.orc_format_decimal.restype = C.c_int
        L.orc_format_decimal.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.orc_decimal_sortkey.restype = C.c_int
        L.orc_decimal_sortkey.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        for name in ("orc_dec_add", "orc_dec_mul", "orc_dec_quo"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                          C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_dec_float64.restype = C.c_double
        L.orc_dec_float64.argtypes = [C.c_uint64, C.c_int, C.c_int]
        assert L.orc_sizeof_q1_group() == C.sizeof(Q1Group), (L.orc_sizeof_q1_group(), C.sizeof(Q1Group))
        assert L.orc_sizeof_q1_result() == C.sizeof(Q1Result)
        assert L.orc_sizeof_q3_group() == C.sizeof(Q3Group)
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ data --

def days(y, m, d):
    return (datetime.date(y, m, d) - datetime.date(1970, 1, 1)).days


SEGMENTS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]

LINEITEM_COLS = [("l_orderkey", np.int64), ("l_partkey", np.int32), ("l_suppkey", np.int32),
                 ("l_linenumber", np.int32), ("l_quantity", np.int32), ("l_extendedprice", np.int64),
                 ("l_discount", np.int64), ("l_tax", np.int64), ("l_returnflag", np.uint8),
                 ("l_linestatus", np.uint8), ("l_shipdate", np.int32), ("l_commitdate", np.int32),
                 ("l_receiptdate", np.int32)]
ORDERS_COLS = [("o_orderkey", np.int64), ("o_custkey", np.int32), ("o_orderdate", np.int32),
               ("o_shippriority", np.int32), ("o_totalprice", np.int64), ("o_orderstatus", np.uint8)]


def gen_orders_lineitem(sf, o_lo=0, o_hi=None, lineitem_cols=None, orders_cols=None):
    """dbgen-equivalent orders [o_lo,o_hi) and their lineitems as numpy columns."""
    L = lib()
    if o_hi is None:
        o_hi = L.tg_num_orders(sf)
    nline = L.tg_count_lineitems(sf, o_lo, o_hi)
    want_l = set(c for c, _ in LINEITEM_COLS) if lineitem_cols is None else set(lineitem_cols)
    want_o = set(c for c, _ in ORDERS_COLS) if orders_cols is None else set(orders_cols)
    orders = {c: (np.empty(o_hi - o_lo, dtype=t) if c in want_o else None) for c, t in ORDERS_COLS}
    line = {c: (np.empty(nline, dtype=t) if c in want_l else None) for c, t in LINEITEM_COLS}
    args = [_p(orders[c]) for c, _ in ORDERS_COLS] + [_p(line[c]) for c, _ in LINEITEM_COLS]
    n = L.tg_gen_orders_lineitem(sf, o_lo, o_hi, *args)
    assert n == nline
    return ({k: v for k, v in orders.items() if v is not None},
            {k: v for k, v in line.items() if v is not None})


def gen_customer(sf, c_lo=0, c_hi=None):
    L = lib()
    if c_hi is None:
        c_hi = L.tg_num_customers(sf)
    out = {"c_custkey": np.empty(c_hi - c_lo, np.int32), "c_mktsegment": np.empty(c_hi - c_lo, np.uint8),
           "c_nationkey": np.empty(c_hi - c_lo, np.int32)}
    L.tg_gen_customer(sf, c_lo, c_hi, _p(out["c_custkey"]), _p(out["c_mktsegment"]), _p(out["c_nationkey"]))
    out["c_name"] = np.array([customer_name(k).encode() for k in out["c_custkey"]], dtype=object)
    return out


def gen_part(sf, like_word=None):
    """dbgen-equivalent part: p_partkey, p_name (numpy object array of bytes) and, for `p_name like '%word%'`
    (wildcardMatch, function_operator_boolean.go:336-377: a substring test for this pattern), the match flag."""
    L = lib()
    n = L.tg_num_parts_pub(C.c_double(sf))
    keys = np.empty(n, np.int32)
    buf = C.create_string_buffer(56 * n)
    off = np.empty(n + 1, np.int64)
    has = np.zeros(n, np.uint8)
    L.tg_gen_part(C.c_double(sf), C.c_int64(0), C.c_int64(n), _p(keys), buf, _p(off),
                  like_word.encode() if like_word else None, _p(has))
    raw = buf.raw
    names = np.array([raw[off[i]:off[i + 1]] for i in range(n)], dtype=object)
    return {"p_partkey": keys, "p_name": names, "p_name_like": has.astype(bool), "like_word": like_word}


def gen_supplier(sf):
    L = lib()
    n = L.tg_num_supp_pub(C.c_double(sf))
    out = {"s_suppkey": np.empty(n, np.int32), "s_nationkey": np.empty(n, np.int32)}
    L.tg_gen_supplier(C.c_double(sf), C.c_int64(0), C.c_int64(n), _p(out["s_suppkey"]), _p(out["s_nationkey"]))
    return out


def gen_partsupp(sf):
    L = lib()
    n = L.tg_num_parts_pub(C.c_double(sf))
    out = {"ps_partkey": np.empty(4 * n, np.int32), "ps_suppkey": np.empty(4 * n, np.int32), "ps_supplycost": np.empty(4 * n, np.int64)}
    L.tg_gen_partsupp(C.c_double(sf), C.c_int64(0), C.c_int64(n), _p(out["ps_partkey"]), _p(out["ps_suppkey"]), _p(out["ps_supplycost"]))
    return out


# dbgen distributions (dists.dss) the draws of tg_gen_q12_q14_draws index; member order pinned by the reference's
# golden cases/tpch/1g/plan/q12.txt (FOB / TRUCK counts) and q14.txt (PROMO share)
SHIPMODES = ["REG AIR", "AIR", "RAIL", "TRUCK", "MAIL", "FOB", "SHIP"]
PRIORITIES = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]
PTYPES = ["%s %s %s" % (a, b, c) for a in ("STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO")
          for b in ("ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED") for c in ("TIN", "NICKEL", "BRASS", "STEEL", "COPPER")]


BRANDS = ["Brand#%d%d" % (m, n) for m in range(1, 6) for n in range(1, 6)]
CONTAINERS = ["%s %s" % (a, b) for a in ("SM", "LG", "MED", "JUMBO", "WRAP") for b in ("CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM")]
SHIPINSTRUCT = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]     # only index 0 is pinned (by q19.txt)


def gen_q19_columns(sf):
    """dictionary codes of l_shipinstruct (per lineitem row), p_brand / p_container (per part) and p_size"""
    L = lib()
    L.tg_gen_q19_draws.restype = None
    L.tg_gen_q19_draws.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    no = L.tg_num_orders(sf)
    nl = L.tg_count_lineitems(sf, 0, no)
    npart = L.tg_num_parts_pub(C.c_double(sf))
    out = {"l_shipinstruct": np.empty(nl, np.uint8), "p_brand": np.empty(npart, np.uint8), "p_size": np.empty(npart, np.int32),
           "p_container": np.empty(npart, np.uint8)}
    L.tg_gen_q19_draws(sf, 0, no, _p(out["l_shipinstruct"]), 0, npart, _p(out["p_brand"]), _p(out["p_size"]), _p(out["p_container"]))
    return out


Q19_GROUPS = (("Brand#23", ("SM CASE", "SM BOX", "SM PACK", "SM PKG"), 5, 5), ("Brand#15", ("MED BAG", "MED BOX", "MED PKG", "MED PACK"), 14, 10),
              ("Brand#44", ("LG CASE", "LG BOX", "LG PACK", "LG PKG"), 28, 15))


def q19(line, extra12, extra19):
    """cases/tpch/query/q19.sql: exact revenue (scale 4).  'AIR REG' is not a ship mode dbgen produces ('REG AIR' is): only AIR matches."""
    pk = line["l_partkey"] - 1
    base = (extra19["l_shipinstruct"] == SHIPINSTRUCT.index("DELIVER IN PERSON")) & np.isin(extra12["l_shipmode"], [SHIPMODES.index("AIR")])
    m = np.zeros(len(pk), bool)
    for brand, conts, q0, smax in Q19_GROUPS:
        m |= (base & (extra19["p_brand"][pk] == BRANDS.index(brand)) & np.isin(extra19["p_container"][pk], [CONTAINERS.index(c) for c in conts]) &
              (line["l_quantity"] >= q0) & (line["l_quantity"] <= q0 + 10) & (extra19["p_size"][pk] >= 1) & (extra19["p_size"][pk] <= smax))
    rev = line["l_extendedprice"][m].astype(object) * (100 - line["l_discount"][m].astype(object))
    return {"revenue": int(rev.sum()) if m.any() else 0, "rows": int(m.sum())}


def gen_q12_q14_columns(sf):
    """dictionary codes of l_shipmode (per lineitem row), o_orderpriority (per order) and p_type (per part)"""
    L = lib()
    L.tg_gen_q12_q14_draws.restype = None
    L.tg_gen_q12_q14_draws.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    no = L.tg_num_orders(sf)
    nl = L.tg_count_lineitems(sf, 0, no)
    npart = L.tg_num_parts_pub(C.c_double(sf))
    out = {"o_orderpriority": np.empty(no, np.uint8), "l_shipmode": np.empty(nl, np.uint8), "p_type": np.empty(npart, np.uint8)}
    L.tg_gen_q12_q14_draws(sf, 0, no, _p(out["o_orderpriority"]), _p(out["l_shipmode"]), 0, npart, _p(out["p_type"]))
    return out


def q12(orders, line, extra, modes=("FOB", "TRUCK"), year=1996):
    """cases/tpch/query/q12.sql restated with exact integer arithmetic: rows (l_shipmode, high_line_count, low_line_count)
    in l_shipmode order; sum(INTEGER) is a HUGEINT (function_aggr.go:620-650)."""
    sm, pr = extra["l_shipmode"], extra["o_orderpriority"]
    m = ((line["l_commitdate"] < line["l_receiptdate"]) & (line["l_shipdate"] < line["l_commitdate"]) &
         (line["l_receiptdate"] >= days(year, 1, 1)) & (line["l_receiptdate"] < days(year + 1, 1, 1)) &
         np.isin(sm, [SHIPMODES.index(x) for x in modes]))
    oidx = np.searchsorted(orders["o_orderkey"], line["l_orderkey"][m])
    ok = (oidx < len(orders["o_orderkey"])) & (orders["o_orderkey"][np.minimum(oidx, len(orders["o_orderkey"]) - 1)] == line["l_orderkey"][m])
    lp, lm = pr[oidx[ok]], sm[m][ok]
    rows = []
    for name in sorted(modes):
        sel = lm == SHIPMODES.index(name)
        if not sel.any():
            continue
        high = int(((lp[sel] == 0) | (lp[sel] == 1)).sum())
        rows.append((name, high, int(sel.sum()) - high))
    return rows


def q4(orders, line, extra, date_lo=None, date_hi=None):
    """cases/tpch/query/q4.sql: orders of the quarter with at least one late lineitem, counted per priority"""
    date_lo = days(1997, 7, 1) if date_lo is None else date_lo
    date_hi = days(1997, 10, 1) if date_hi is None else date_hi
    late = np.unique(line["l_orderkey"][line["l_commitdate"] < line["l_receiptdate"]])
    m = (orders["o_orderdate"] >= date_lo) & (orders["o_orderdate"] < date_hi) & np.isin(orders["o_orderkey"], late)
    pr = extra["o_orderpriority"][m]
    return [(PRIORITIES[p], int((pr == p).sum())) for p in range(5) if (pr == p).any()]


def q4_text(rows):
    return "#\t\n" + "".join("%s\t%d\n" % r for r in rows)


def q12_text(rows):
    return "#\t\t\n" + "".join("%s\t%d\t%d\n" % r for r in rows)


def q14(line, extra, date_lo=None, date_hi=None, like_prefix="PROMO"):
    """cases/tpch/query/q14.sql: the two DECIMAL sums exactly (scale 4), and promo_revenue the way the reference computes it
    above the aggregate: the literal 100.00 is FLOAT, so both sums are cast to FLOAT and the arithmetic is float32
    (builder_binder.go:264-273; binFloat32MultiOp / binFloat32DivOp)."""
    date_lo = days(1996, 4, 1) if date_lo is None else date_lo
    date_hi = days(1996, 5, 1) if date_hi is None else date_hi
    m = (line["l_shipdate"] >= date_lo) & (line["l_shipdate"] < date_hi)
    rev = line["l_extendedprice"][m].astype(object) * (100 - line["l_discount"][m].astype(object))
    ptype = extra["p_type"][line["l_partkey"][m] - 1]
    promo_codes = [i for i, t in enumerate(PTYPES) if t.startswith(like_prefix)]
    promo = int(rev[np.isin(ptype, promo_codes)].sum()) if m.any() else 0
    total = int(rev.sum()) if m.any() else 0
    return {"promo": promo, "total": total, "rows": int(m.sum())}


def q14_promo_revenue(promo, total):
    """100.00 * promo / total in float32, printed like a Go float64 holding that float32 (Value.String of a FLOAT)"""
    with np.errstate(all="ignore"):
        f = np.float32(100.0) * np.float32(float(C.c_double(lib().orc_dec_float64(abs(promo), 4, int(promo < 0))).value)) / \
            np.float32(float(C.c_double(lib().orc_dec_float64(abs(total), 4, int(total < 0))).value))
    return fmt_double(float(f))


def gen_q11_q22_columns(sf):
    """c_acctbal (cents, per customer) and ps_availqty (per partsupp row)"""
    L = lib()
    L.tg_gen_q11_q22_draws.restype = None
    L.tg_gen_q11_q22_draws.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    nc, npart = L.tg_num_customers(sf), L.tg_num_parts_pub(C.c_double(sf))
    out = {"c_acctbal": np.empty(nc, np.int64), "ps_availqty": np.empty(4 * npart, np.int32)}
    L.tg_gen_q11_q22_draws(sf, 0, nc, _p(out["c_acctbal"]), 0, npart, _p(out["ps_availqty"]))
    return out


# dbgen's fixed nation -> region assignment (nations of dists.dss); regions: 0 AFRICA, 1 AMERICA, 2 ASIA, 3 EUROPE, 4 MIDDLE EAST
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
NATION_REGION = [0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1]


def _year(d):
    return d.astype("datetime64[D]").astype("datetime64[Y]").astype(np.int64) + 1970


def _line_dims(cust, supp, orders, line):
    """per lineitem row: index of its order, customer nation, supplier nation (the joins every query below shares)"""
    oidx = np.searchsorted(orders["o_orderkey"], line["l_orderkey"])
    return oidx, cust["c_nationkey"][orders["o_custkey"][oidx] - 1], supp["s_nationkey"][line["l_suppkey"] - 1]


def _revenue(line, m):
    return (line["l_extendedprice"][m] * (100 - line["l_discount"][m])).astype(object)          # scale 4, exact


def q5(cust, supp, orders, line, region="AMERICA", year=1994):
    """cases/tpch/query/q5.sql: revenue (scale 4) per nation of the region where customer and supplier share the nation; revenue desc"""
    names, nreg = nation_names(), np.array(NATION_REGION)
    oidx, cn, sn = _line_dims(cust, supp, orders, line)
    od = orders["o_orderdate"][oidx]
    m = (od >= days(year, 1, 1)) & (od < days(year + 1, 1, 1)) & (cn == sn) & (nreg[sn] == REGIONS.index(region))
    rev, key = _revenue(line, m), sn[m]
    rows = [(names[k], int(rev[key == k].sum())) for k in np.unique(key)]
    return sorted(rows, key=lambda r: -r[1])


def q7(cust, supp, orders, line, a="FRANCE", b="ARGENTINA"):
    """cases/tpch/query/q7.sql: (supp_nation, cust_nation, l_year, revenue) in key order"""
    names = nation_names()
    _, cn, sn = _line_dims(cust, supp, orders, line)
    A, B = names.index(a), names.index(b)
    sd = line["l_shipdate"]
    m = (sd >= days(1995, 1, 1)) & (sd <= days(1996, 12, 31)) & (((sn == A) & (cn == B)) | ((sn == B) & (cn == A)))
    out = {}
    for s, c, y, v in zip(sn[m], cn[m], _year(sd[m]), _revenue(line, m)):
        k = (names[s], names[c], int(y))
        out[k] = out.get(k, 0) + int(v)
    return [k + (out[k],) for k in sorted(out)]


def q8(cust, supp, orders, line, extra12, nation="ARGENTINA", region="AMERICA", ptype="ECONOMY BURNISHED TIN"):
    """cases/tpch/query/q8.sql: (o_year, mkt_share, the two exact sums at scale 4); the share is a DECIMAL quotient of the two exact sums (govalues Quo), printed at
    the type's scale 4"""
    names, nreg = nation_names(), np.array(NATION_REGION)
    oidx, cn, sn = _line_dims(cust, supp, orders, line)
    od = orders["o_orderdate"][oidx]
    m = ((od >= days(1995, 1, 1)) & (od <= days(1996, 12, 31)) & (extra12["p_type"][line["l_partkey"] - 1] == PTYPES.index(ptype)) &
         (nreg[cn] == REGIONS.index(region)))
    rev, yr, mine = _revenue(line, m), _year(od[m]), sn[m] == names.index(nation)
    rows = []
    for y in np.unique(yr):
        part, tot = int(rev[(yr == y) & mine].sum()) if ((yr == y) & mine).any() else 0, int(rev[yr == y].sum())
        oc, os_, on = C.c_uint64(), C.c_int(), C.c_int()
        assert lib().orc_dec_quo(C.c_uint64(part), 4, 0, C.c_uint64(tot), 4, 0, C.byref(oc), C.byref(os_), C.byref(on)) == 0
        rows.append((int(y), (oc.value, os_.value, on.value), part, tot))
    return rows


def q11(supp, partsupp, extra11, nation="JAPAN", fraction_inv=10000):
    """cases/tpch/query/q11.sql: (ps_partkey, value) with value = sum(ps_supplycost * ps_availqty) (scale 2) above total * 0.0001,
    value desc.  The threshold literal is FLOAT in the reference (the comparison runs in float32); at SF1 no group lies within 1300.00
    of the threshold (float32 resolves 0.5 there), so the exact comparison selects the same rows."""
    m = supp["s_nationkey"][partsupp["ps_suppkey"] - 1] == nation_names().index(nation)
    val = partsupp["ps_supplycost"][m] * extra11["ps_availqty"][m].astype(np.int64)
    u, inv = np.unique(partsupp["ps_partkey"][m], return_inverse=True)
    sums = np.zeros(len(u), np.int64)
    np.add.at(sums, inv, val)
    tot = int(val.sum())
    return [(int(u[i]), int(sums[i])) for i in np.lexsort((u, -sums)) if int(sums[i]) * fraction_inv > tot]


def q17(line, extra19, brand="Brand#54", container="LG BAG"):
    """cases/tpch/query/q17.sql: sum(l_extendedprice) over the lines whose quantity is below 0.2 * avg(l_quantity) of their part,
    divided by the FLOAT literal 7.0 in float32 (like Q14's 100.00).  At SF1 no line sits on the 0.2 * avg boundary: float32, float64
    and exact rational comparison select the same 558 lines, so the golden file pins the sum and the float32 division only."""
    pk = line["l_partkey"] - 1
    npart = len(extra19["p_brand"])
    sq, cq = np.bincount(pk, weights=line["l_quantity"], minlength=npart), np.bincount(pk, minlength=npart)
    m = ((extra19["p_brand"] == BRANDS.index(brand)) & (extra19["p_container"] == CONTAINERS.index(container)))[pk]
    q, p = line["l_quantity"][m], pk[m]
    sel = q.astype(np.float32) < np.float32(0.2) * (sq[p] / cq[p]).astype(np.float32)
    s = int(line["l_extendedprice"][m][sel].sum())
    f = np.float32(float(C.c_double(lib().orc_dec_float64(abs(s), 2, 0)).value)) / np.float32(7.0)
    return {"sum": s, "rows": int(sel.sum()), "avg_yearly": fmt_double(float(f))}


def q21(supp, orders, line, nation="BRAZIL", limit=100):
    """cases/tpch/query/q21.sql: suppliers of the nation that were the ONLY late supplier of a multi-supplier order with status F;
    (s_name, numwait) by numwait desc, s_name"""
    ok, sk = line["l_orderkey"], line["l_suppkey"]
    late = line["l_receiptdate"] > line["l_commitdate"]
    oid = np.searchsorted(orders["o_orderkey"], ok)
    no = len(orders["o_orderkey"])

    def suppliers_per_order(mask):
        pair = np.unique(oid[mask].astype(np.int64) * (1 << 32) + sk[mask])
        return np.bincount(pair >> 32, minlength=no)
    nsup, nlate = suppliers_per_order(np.ones(len(ok), bool)), suppliers_per_order(late)
    m = (late & (orders["o_orderstatus"][oid] == ord("F")) & (supp["s_nationkey"][sk - 1] == nation_names().index(nation)) &
         (nsup[oid] > 1) & (nlate[oid] == 1))
    cnt = np.bincount(sk[m])
    rows = sorted((-int(c), "Supplier#%09d" % s) for s, c in enumerate(cnt) if c)[:limit]
    return [(n, -c) for c, n in rows]


def q22(cust, orders, extra22, codes=(10, 11, 26, 22, 19, 20, 27)):
    """cases/tpch/query/q22.sql: customers of the listed country codes (substring(c_phone, 1, 2) = 10 + c_nationkey in dbgen) without
    orders whose balance exceeds the average positive balance of those codes; (cntrycode, numcust, totacctbal scale 2).
    c_acctbal > avg is decided exactly (acctbal * n > sum): avg is a 19-digit quotient, a 2-digit balance never ties with it unless
    the quotient is exact, where both forms agree."""
    ab = extra22["c_acctbal"]
    code = cust["c_nationkey"] + 10
    inl = np.isin(code, list(codes))
    pos = inl & (ab > 0)
    s, n = int(ab[pos].sum()), int(pos.sum())
    has = np.zeros(len(ab) + 1, bool)
    has[orders["o_custkey"]] = True
    m = inl & (ab * n > s) & ~has[cust["c_custkey"]]
    return [(int(c), int((m & (code == c)).sum()), int(ab[m & (code == c)].sum())) for c in sorted(set(code[m].tolist()))]


def _addresses(buf, n):
    raw = buf.raw
    return [raw[41 * i:41 * i + 41].split(b"\0")[0].decode() for i in range(n)]


def _phones(nat, ph):
    return ["%d-%d-%d-%d" % (nat[i] + 10, ph[3 * i], ph[3 * i + 1], ph[3 * i + 2]) for i in range(len(nat))]


def gen_supplier_text(sf):
    """s_address / s_phone (lists of str), s_acctbal (cents) and the `s_comment like '%Customer%Complaints%'` flag per supplier"""
    L = lib()
    n = L.tg_num_supp_pub(C.c_double(sf))
    buf, ph, cp, ab = C.create_string_buffer(41 * n), np.empty(3 * n, np.int32), np.empty(n, np.uint8), np.empty(n, np.int64)
    L.tg_gen_supplier_text.restype = None
    L.tg_gen_supplier_text.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tg_gen_supplier_text(sf, 0, n, buf, _p(ph), _p(cp), _p(ab))
    return {"s_address": _addresses(buf, n), "s_phone": _phones(gen_supplier(sf)["s_nationkey"], ph), "complaint": cp.astype(bool), "s_acctbal": ab}


def gen_customer_text(sf, cust):
    """c_address / c_phone (lists of str) per customer"""
    L = lib()
    n = len(cust["c_custkey"])
    buf, ph = C.create_string_buffer(41 * n), np.empty(3 * n, np.int32)
    L.tg_gen_customer_text.restype = None
    L.tg_gen_customer_text.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    L.tg_gen_customer_text(sf, 0, n, buf, _p(ph))
    return {"c_address": _addresses(buf, n), "c_phone": _phones(cust["c_nationkey"], ph)}


def q15(line, stext, date_lo=None, date_hi=None):
    """cases/tpch/query/q15.sql: the supplier(s) with the largest revenue (scale 4) of the quarter;
    (s_suppkey, s_name, s_address, s_phone, total_revenue)"""
    date_lo = days(1995, 12, 1) if date_lo is None else date_lo
    date_hi = days(1996, 3, 1) if date_hi is None else date_hi
    m = (line["l_shipdate"] >= date_lo) & (line["l_shipdate"] < date_hi)
    rev = np.zeros(len(stext["s_address"]) + 1, np.int64)          # < 2^63: a supplier holds ~600 lines of < 1e11 each
    np.add.at(rev, line["l_suppkey"][m], line["l_extendedprice"][m] * (100 - line["l_discount"][m]))
    best = int(rev.max())
    return [(int(k), "Supplier#%09d" % k, stext["s_address"][k - 1], stext["s_phone"][k - 1], best) for k in np.nonzero(rev == best)[0]]


def q16(partsupp, extra12, extra19, stext, brand="Brand#35", type_prefix="ECONOMY BURNISHED", sizes=(14, 7, 21, 24, 35, 33, 2, 20)):
    """cases/tpch/query/q16.sql: count(distinct ps_suppkey) per (p_brand, p_type, p_size) without the complaint suppliers;
    supplier_cnt desc, then the keys"""
    pk = partsupp["ps_partkey"] - 1
    b, t, z = extra19["p_brand"][pk], extra12["p_type"][pk], extra19["p_size"][pk]
    m = ((b != BRANDS.index(brand)) & ~np.isin(t, [i for i, n in enumerate(PTYPES) if n.startswith(type_prefix)]) & np.isin(z, list(sizes)) &
         ~stext["complaint"][partsupp["ps_suppkey"] - 1])
    key = (b[m].astype(np.int64) * 150 + t[m]) * 64 + z[m]
    g, cnt = np.unique(np.unique(key * (1 << 32) + partsupp["ps_suppkey"][m]) >> 32, return_counts=True)
    rows = sorted((-int(c), BRANDS[int(k) // 64 // 150], PTYPES[int(k) // 64 % 150], int(k) % 64) for k, c in zip(g, cnt))
    return [(bn, tn, sz, -c) for c, bn, tn, sz in rows]


def q20(part, supp, partsupp, line, extra11, stext, prefix=b"lime", nation="VIETNAM", year=1993):
    """cases/tpch/query/q20.sql: suppliers of the nation holding more than 0.5 * (the year's shipped quantity) of a part named
    prefix%; (s_name, s_address) by s_name.  0.5 is a FLOAT literal: the comparison is float32, exact here (0.5 and sums < 2^24)."""
    lime = np.array([nm.startswith(prefix) for nm in part["p_name"]])
    m = (line["l_shipdate"] >= days(year, 1, 1)) & (line["l_shipdate"] < days(year + 1, 1, 1)) & lime[line["l_partkey"] - 1]
    u, inv = np.unique(line["l_partkey"][m].astype(np.int64) * (1 << 32) + line["l_suppkey"][m], return_inverse=True)
    sq = np.zeros(len(u), np.int64)
    np.add.at(sq, inv, line["l_quantity"][m])
    pkey = partsupp["ps_partkey"].astype(np.int64) * (1 << 32) + partsupp["ps_suppkey"]
    pos = np.minimum(np.searchsorted(u, pkey), max(len(u) - 1, 0))
    sel = (u[pos] == pkey) & (extra11["ps_availqty"].astype(np.float32) > np.float32(0.5) * sq[pos].astype(np.float32))   # no lines: NULL, not true
    nk = nation_names().index(nation)
    return [("Supplier#%09d" % k, stext["s_address"][k - 1]) for k in np.unique(partsupp["ps_suppkey"][sel]).tolist() if supp["s_nationkey"][k - 1] == nk]


COMMENT_STREAMS = {"c_comment": (1335826707, 73), "s_comment": (1341315363, 63), "o_comment": (276090261, 49)}   # (stream seed, average length)


def comments(column, rows):
    """the generated text of `column` for the given 0-based rows (dbgen cuts every comment out of one 300 MiB text pool;
    oracle/tpchgen.c restates the pool).  The ~10 in 10000 suppliers whose comment dbgen overwrites with "Customer ... Complaints /
    Recommends" (gen_supplier_text's flag covers the Complaints half) come back with their un-overwritten text."""
    L = lib()
    L.tg_text_pool.restype = C.c_void_p
    L.tg_comment_spans.restype = None
    L.tg_comment_spans.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    base = L.tg_text_pool()
    assert base, "text pool allocation failed"
    seed, avg = COMMENT_STREAMS[column]
    out = []
    off, ln = np.empty(1, np.int64), np.empty(1, np.int32)
    for r in rows:
        L.tg_comment_spans(seed, avg, int(r), int(r) + 1, _p(off), _p(ln))
        out.append(C.string_at(base + int(off[0]), int(ln[0])).decode())
    return out


def q13(orders, ncust, w1=b"pending", w2=b"accounts"):
    """cases/tpch/query/q13.sql: customers LEFT JOIN orders (o_comment not like '%pending%accounts%'), count(o_orderkey) per
    customer, then customers per count; custdist desc, c_count desc.  A customer without a surviving order has c_count NULL in the
    reference (count over no non-NULL input, function_aggr.go), printed "NULL"."""
    L = lib()
    L.tg_comments_like2.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_char_p, C.c_char_p, C.c_void_p]
    flags = np.empty(len(orders["o_custkey"]), np.uint8)
    seed, avg = COMMENT_STREAMS["o_comment"]
    assert L.tg_comments_like2(seed, avg, 0, len(flags), w1, w2, _p(flags)) == 0
    cnt = np.bincount(orders["o_custkey"][flags == 0], minlength=ncust + 1)[1:]
    dist = np.bincount(cnt)
    rows = sorted((-int(n), -int(c)) for c, n in enumerate(dist) if n)
    return [(None if c == 0 else -c, -n) for n, c in rows]


def q10(cust, orders, line, extra22, ctext, date_lo=None, date_hi=None, limit=20):
    """cases/tpch/query/q10.sql: (c_custkey, c_name, revenue scale 4, c_acctbal scale 2 signed, n_name, c_address, c_phone, c_comment),
    revenue desc, limit 20"""
    date_lo = days(1993, 3, 1) if date_lo is None else date_lo
    date_hi = days(1993, 6, 1) if date_hi is None else date_hi
    oidx = np.searchsorted(orders["o_orderkey"], line["l_orderkey"])
    od = orders["o_orderdate"][oidx]
    m = (od >= date_lo) & (od < date_hi) & (line["l_returnflag"] == ord("R"))
    rev = np.zeros(len(cust["c_custkey"]) + 1, np.int64)
    np.add.at(rev, orders["o_custkey"][oidx][m], line["l_extendedprice"][m] * (100 - line["l_discount"][m]))
    names = nation_names()
    top = sorted((-int(rev[k]), int(k)) for k in np.nonzero(rev)[0])[:limit]
    cm = comments("c_comment", [k - 1 for _, k in top])
    return [(k, "Customer#%09d" % k, -r, int(extra22["c_acctbal"][k - 1]), names[cust["c_nationkey"][k - 1]], ctext["c_address"][k - 1],
             ctext["c_phone"][k - 1], cm[i]) for i, (r, k) in enumerate(top)]


def q2(supp, partsupp, extra12, extra19, stext, size=48, type_suffix="TIN", region="MIDDLE EAST", limit=100):
    """cases/tpch/query/q2.sql: (s_acctbal scale 2 signed, s_name, n_name, p_partkey, p_mfgr, s_address, s_phone, s_comment) of the
    region's cheapest supplier(s) of every matching part; s_acctbal desc, n_name, s_name, p_partkey"""
    names, nreg = nation_names(), np.array(NATION_REGION)
    in_reg = nreg[supp["s_nationkey"][partsupp["ps_suppkey"] - 1]] == REGIONS.index(region)
    mincost = np.full(len(extra19["p_size"]) + 1, 1 << 62, np.int64)
    np.minimum.at(mincost, partsupp["ps_partkey"][in_reg], partsupp["ps_supplycost"][in_reg])
    pk = partsupp["ps_partkey"] - 1
    suffix = np.array([t.endswith(type_suffix) for t in PTYPES])
    m = in_reg & (extra19["p_size"][pk] == size) & suffix[extra12["p_type"][pk]] & (partsupp["ps_supplycost"] == mincost[partsupp["ps_partkey"]])
    rows = []
    for p, k in zip(partsupp["ps_partkey"][m].tolist(), partsupp["ps_suppkey"][m].tolist()):
        rows.append((-int(stext["s_acctbal"][k - 1]), names[supp["s_nationkey"][k - 1]], "Supplier#%09d" % k, p,
                     "Manufacturer#%d" % (extra19["p_brand"][p - 1] // 5 + 1), stext["s_address"][k - 1], stext["s_phone"][k - 1], k))
    rows = sorted(rows)[:limit]
    cm = comments("s_comment", [r[7] - 1 for r in rows])
    return [(-b, sn, nn, p, mf, a, ph, cm[i]) for i, (b, nn, sn, p, mf, a, ph, _k) in enumerate(rows)]


def rows_text(header_tabs, rows):
    """the reference's result file: '#' + one tab per column after the first, then tab-separated rows"""
    return "#" + "\t" * header_tabs + "\n" + "".join("\t".join(str(x) for x in r) + "\n" for r in rows)


def nation_names():
    L = lib()
    L.tg_nation_name.restype = C.c_char_p
    return [L.tg_nation_name(i).decode() for i in range(25)]


# --------------------------------------------------------------- queries --

def _i128(lo, hi):
    return (int(hi) << 64) + int(lo)


def q6(line, date_lo=days(1994, 1, 1), date_hi=days(1995, 1, 1), disc_lit=0.03, disc_eps=0.01, qty_lt=24):
    r = Q6Result()
    n = len(line["l_shipdate"])
    lib().orc_q6(n, _p(line["l_shipdate"]), _p(line["l_discount"]), _p(line["l_quantity"]),
                 _p(line["l_extendedprice"]), date_lo, date_hi, disc_lit, disc_eps, qty_lt, C.byref(r))
    assert r.error == 0
    return {"rows_in": r.rows_in, "rows_selected": r.rows_selected, "has_row": bool(r.has_row),
            "sum": r.sum.tuple(), "exact": _i128(r.exact_lo, r.exact_hi)}


def q1(line, ship_le=days(1998, 8, 11)):
    r = Q1Result()
    n = len(line["l_shipdate"])
    lib().orc_q1(n, _p(line["l_shipdate"]), _p(line["l_returnflag"]), _p(line["l_linestatus"]),
                 _p(line["l_quantity"]), _p(line["l_extendedprice"]), _p(line["l_discount"]),
                 _p(line["l_tax"]), ship_le, C.byref(r))
    assert r.error == 0, r.error
    groups = []
    for i in range(r.ngroups):
        g = r.g[i]
        groups.append({
            "l_returnflag": chr(g.rf), "l_linestatus": chr(g.ls),
            "sum_qty": g.sum_qty.value(), "sum_base_price": g.sum_base.tuple(),
            "sum_disc_price": g.sum_disc_price.tuple(), "sum_charge": g.sum_charge.tuple(),
            "avg_qty": g.avg_qty, "avg_price": g.avg_price.tuple(), "avg_disc": g.avg_disc.tuple(),
            "count_order": int(g.count),
            "x_qty": int(g.x_qty), "x_base": _i128(g.x_base_lo, g.x_base_hi),
            "x_disc_price": _i128(g.x_disc_price_lo, g.x_disc_price_hi),
            "x_charge": _i128(g.x_charge_lo, g.x_charge_hi), "x_disc": _i128(g.x_disc_lo, g.x_disc_hi),
            "first_row": int(g.first_row)})
    return {"rows_in": r.rows_in, "rows_selected": r.rows_selected, "groups": groups}


def q3(cust, orders, line, segment="HOUSEHOLD", odate_lt=days(1995, 3, 29), ship_gt=days(1995, 3, 29),
       capacity=None):
    r = Q3Result()
    if isinstance(segment, str):
        seg_code = SEGMENTS.index(segment) if segment in SEGMENTS else 255     # unknown literal matches nothing
    else:
        seg_code = int(segment)
    if capacity is None:
        capacity = max(len(orders["o_orderkey"]), 1)
    out = (Q3Group * capacity)()
    lib().orc_q3(len(cust["c_custkey"]), _p(cust["c_custkey"]), _p(cust["c_mktsegment"]), seg_code,
                 len(orders["o_orderkey"]), _p(orders["o_orderkey"]), _p(orders["o_custkey"]),
                 _p(orders["o_orderdate"]), _p(orders["o_shippriority"]), odate_lt,
                 len(line["l_orderkey"]), _p(line["l_orderkey"]), _p(line["l_extendedprice"]),
                 _p(line["l_discount"]), _p(line["l_shipdate"]), ship_gt,
                 C.cast(out, C.c_void_p), capacity, C.byref(r))
    assert r.error == 0, r.error
    groups = [{"l_orderkey": int(g.orderkey), "revenue": g.revenue.tuple(), "o_orderdate": int(g.orderdate),
               "o_shippriority": int(g.shippriority), "x_revenue": _i128(g.x_rev_lo, g.x_rev_hi),
               "first_row": int(g.first_row)} for g in out[:r.nout]]
    stats = {k: int(getattr(r, k)) for k in ("n_cust_sel", "n_orders_sel", "n_orders_joined",
                                              "n_line_sel", "n_line_joined", "ngroups")}
    return {"groups": groups, "stats": stats}


def _lut(chars):
    if chars is None:
        return None
    t = np.zeros(256, dtype=np.uint8)
    for ch in chars:
        t[ord(ch)] = 1
    return t


def stats(line, d0, d1, d2, d3, q0, q1, disc_gt_cents, valid=None, linestatus_in=None, returnflag_in=None):
    """The wider 'stats' shape (min/max/sum/avg/count, 7 comparisons, 1 key) -- see refexec.c orc_stats.
    valid: optional {column: bool array} for l_quantity / l_extendedprice / l_tax (False = NULL); aggregate
    results whose inputs were all NULL come back as None."""
    r = StatsResult()
    valid = valid or {}
    vb = {k: np.ascontiguousarray(v, dtype=np.uint8) for k, v in valid.items()}
    lib().orc_stats(len(line["l_shipdate"]), _p(line["l_shipdate"]), _p(line["l_commitdate"]), _p(line["l_receiptdate"]),
                    _p(line["l_quantity"]), _p(line["l_extendedprice"]), _p(line["l_discount"]), _p(line["l_tax"]),
                    _p(line["l_returnflag"]), d0, d1, d2, d3, q0, q1, disc_gt_cents,
                    _p(vb.get("l_quantity")), _p(vb.get("l_extendedprice")), _p(vb.get("l_tax")),
                    _p(line["l_linestatus"]), _p(_lut(linestatus_in)), _p(_lut(returnflag_in)), C.byref(r))
    assert r.error == 0, r.error
    groups = [{"l_returnflag": chr(g.rf),
               "min_ext": g.min_ext.tuple() if g.n_ext else None, "max_ext": g.max_ext.tuple() if g.n_ext else None,
               "max_disc": g.max_disc.tuple(), "sum_tax": g.sum_tax.tuple() if g.n_tax else None,
               "avg_tax": g.avg_tax.tuple() if g.n_tax else None,
               "sum_taxed": g.sum_taxed.tuple() if g.n_taxed else None,
               "count": int(g.count), "first_row": int(g.first_row)} for g in r.g[:r.ngroups]]
    return {"rows_selected": int(r.rows_selected), "groups": groups}


def groupby_sum(line, key="l_orderkey", value="l_quantity", having_gt=None, ship_le=None):
    """numpy restatement of a high-cardinality hash aggregate (exact int64 arithmetic):
    GroupedAggrHashTable.FindOrCreateGroups (aggregate_hash.go:201-391) + sum(INT32)->HUGEINT /
    sum(DECIMAL) / count (function_aggr.go:620-650, 950-962) + HAVING via greatHugeintOp
    (function_operator_boolean.go:255-263).  Returns {key: (sum, count)}."""
    k, v = line[key], line[value].astype(np.int64)
    if ship_le is not None:
        m = line["l_shipdate"] <= ship_le
        k, v = k[m], v[m]
    uk, inv = np.unique(k, return_inverse=True)
    sums = np.zeros(len(uk), dtype=np.int64)
    np.add.at(sums, inv, v)
    cnts = np.bincount(inv, minlength=len(uk))
    out = {int(a): (int(b), int(c)) for a, b, c in zip(uk, sums, cnts)}
    if having_gt is not None:
        out = {a: bc for a, bc in out.items() if bc[0] > having_gt}
    return out


def semi_groupby(orders, line, anti=False, odate_lt=days(1995, 3, 29), ship_gt=days(1995, 3, 29),
                 valid_okey=None, valid_lkey=None):
    """numpy restatement of a SEMI / ANTI hash join under a hash aggregate: a probe row is emitted once
    iff some / no build row has its key (Scan.NextSemiOrAntiJoin, join_scan.go:90-165; NULL keys never
    match, join_table.go:152-195).  Returns {o_custkey: (sum(o_totalprice) cents, count)}."""
    bm = line["l_shipdate"] > ship_gt
    if valid_lkey is not None:
        bm = bm & valid_lkey                      # NULL build keys are dropped (join_table.go:152-195)
    build_keys = line["l_orderkey"][bm]
    m = orders["o_orderdate"] < odate_lt
    hit = np.isin(orders["o_orderkey"], build_keys)
    if valid_okey is not None:
        hit = hit & valid_okey                    # a NULL probe key finds no match: SEMI drops it, ANTI keeps it
    m &= ~hit if anti else hit
    k, v = orders["o_custkey"][m], orders["o_totalprice"][m].astype(np.int64)
    uk, inv = np.unique(k, return_inverse=True)
    sums = np.zeros(len(uk), dtype=np.int64)
    np.add.at(sums, inv, v)
    cnts = np.bincount(inv, minlength=len(uk))
    return {int(a): (int(b), int(c)) for a, b, c in zip(uk, sums, cnts)}


# ------------------------------------------------------------- formatting --

def fmt_decimal(dec, type_scale):
    buf = C.create_string_buffer(64)
    n = lib().orc_format_decimal(dec[0], dec[1], dec[2], type_scale, buf, 64)
    assert n >= 0
    return buf.value.decode()


def decimal_sortkey(dec):
    w, f = C.c_int64(), C.c_int64()
    ok = lib().orc_decimal_sortkey(dec[0], dec[1], dec[2], C.byref(w), C.byref(f))
    assert ok
    return (w.value, f.value)


def fmt_date(d):
    return (datetime.date(1970, 1, 1) + datetime.timedelta(days=int(d))).isoformat()


def fmt_double(x):
    """Go fmt.Sprintf("%v", float64) (chunk/value.go:55-58) = strconv.FormatFloat(x, 'g', -1, 64): the shortest digits that
    round-trip, in %e form when the decimal exponent is < -4 or >= 6 -- strconv's formatDigits takes eprec = 6 when the precision is
    "shortest" (float64(time.Second) prints 1e+09, 123456789.0 prints 1.23456789e+08; the 1e21 cut-off is encoding/json's and
    JavaScript's, not fmt's).  %e: d.ddde+XX with at least two exponent digits; %f: plain digits without a trailing ".0".
    No golden file of the reference holds a DOUBLE of 1e6 or more: the upper cut-off is restated from the Go library, not pinned."""
    import math
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "+Inf" if x > 0 else "-Inf"
    sign = "-" if math.copysign(1.0, x) < 0 else ""
    if x == 0:
        return sign + "0"
    m, _, e = ("%r" % abs(x)).partition("e")            # Python's repr is the shortest round-trip digit string too
    ip, _, fp = m.partition(".")
    e10 = int(e) if e else 0
    if ip.strip("0"):
        dp = len(ip) + e10                              # value = 0.DIGITS * 10^dp
    else:
        dp = e10 - (len(fp) - len(fp.lstrip("0")))
    digits = (ip + fp).strip("0") or "0"
    exp = dp - 1
    if exp < -4 or exp >= 6:
        return "%s%s%se%s%02d" % (sign, digits[0], "." + digits[1:] if len(digits) > 1 else "", "+" if exp >= 0 else "-", abs(exp))
    if dp <= 0:
        return sign + "0." + "0" * (-dp) + digits
    if len(digits) <= dp:
        return sign + digits + "0" * (dp - len(digits))
    return sign + digits[:dp] + "." + digits[dp:]


def q1_text(res):
    """Rows as the reference writes them (ORDER BY l_returnflag, l_linestatus)."""
    rows = sorted(res["groups"], key=lambda g: (g["l_returnflag"], g["l_linestatus"]))
    lines = ["#" + "\t" * 9]
    for g in rows:
        lines.append("\t".join([
            g["l_returnflag"], g["l_linestatus"], str(g["sum_qty"]),
            fmt_decimal(g["sum_base_price"], 2), fmt_decimal(g["sum_disc_price"], 4),
            fmt_decimal(g["sum_charge"], 8), fmt_double(g["avg_qty"]),
            fmt_decimal(g["avg_price"], 2), fmt_decimal(g["avg_disc"], 2), str(g["count_order"])]))
    return "\n".join(lines) + "\n"


def q6_text(res):
    lines = ["#"]
    if res["has_row"]:
        lines.append(fmt_decimal(res["sum"], 4))
    return "\n".join(lines) + "\n"


def q3_topk(res, limit=10):
    """ORDER BY revenue DESC (key rounded to 2 digits), o_orderdate ; LIMIT."""
    def key(g):
        w, f = decimal_sortkey(g["revenue"])
        return (-w, -f, g["o_orderdate"])
    return sorted(res["groups"], key=key)[:limit]


def wildcard_match(pattern, target):
    """wildcardMatch (function_operator_boolean.go:336-377) restated: bytes; % any run, _ any byte."""
    p = t = 0
    star_p = star_t = -1
    plen, tlen = len(pattern), len(target)
    while t < tlen:
        if p < plen and pattern[p] == 0x25:
            p += 1
            star_p = p
            if p >= plen:
                return True
            star_t = t
        elif p < plen and (pattern[p] == 0x5F or pattern[p] == target[t]):
            p += 1
            t += 1
        else:
            if star_p == -1 or star_t == -1:
                return False
            p = star_p
            star_t += 1
            t = star_t
    while p < plen and pattern[p] == 0x25:
        p += 1
    return p >= plen


def customer_filter(cust, filters, segments):
    """select count(*), sum(c_nationkey), sum(c_custkey) from customer where <string predicates>
    (likeOp / notLikeOp / equalStrOp, function_operator_boolean.go:99-104,336-392).  Returns None for an empty
    selection (the aggregate emits no row), else (count, sum, sum)."""
    n = len(cust["c_custkey"])
    keep = np.ones(n, dtype=bool)
    for cname, op, lit in filters:
        lit = lit.encode()
        if cname == "c_mktsegment":
            vals = [segments[c].encode() for c in cust[cname]]
        else:
            vals = list(cust[cname])
        if op in ("like", "not like"):
            m = np.array([wildcard_match(lit, v) for v in vals], dtype=bool)
            keep &= m if op == "like" else ~m
        else:
            m = np.array([v == lit for v in vals], dtype=bool)
            keep &= m if op == "=" else ~m
    if not keep.any():
        return None
    return (int(keep.sum()), int(cust["c_nationkey"][keep].astype(np.int64).sum()), int(cust["c_custkey"][keep].astype(np.int64).sum()))


def customer_name(custkey):
    """dbgen C_NAME: the tag "Customer" + '#' + the key zero-padded to 9 digits (TPC-H spec 4.2.3)."""
    return "Customer#%09d" % int(custkey)


def q18(cust, orders, line, qty_gt=314, limit=100):
    """TPC-H Q18 as the reference plans it (cases/tpch/query/q18.sql): the IN-subquery is a SEMI join on
    o_orderkey against `group by l_orderkey having sum(l_quantity) > k` (builder_plan.go:234-262; the
    HAVING compares HUGEINTs, function_operator_boolean.go:255-263); the outer aggregate groups by
    (c_name, c_custkey, o_orderkey, o_orderdate, o_totalprice) with sum(l_quantity) -> HUGEINT
    (function_aggr.go:620-650); ORDER BY o_totalprice DESC (DECIMAL key = Int64(2), sort_encoder.go:65-70),
    o_orderdate; LIMIT (executor_limit.go:120-137).  Returns the ordered rows as dicts."""
    lk, lq = line["l_orderkey"], line["l_quantity"].astype(np.int64)
    uk, inv = np.unique(lk, return_inverse=True)
    sums = np.zeros(len(uk), dtype=np.int64)
    np.add.at(sums, inv, lq)
    big = uk[sums > qty_gt]
    sel = np.nonzero(np.isin(orders["o_orderkey"], big))[0]
    ckeys = set(int(x) for x in cust["c_custkey"])
    qsum = dict(zip((int(k) for k in uk), (int(v) for v in sums)))
    rows = []
    for i in sel:
        ck = int(orders["o_custkey"][i])
        if ck not in ckeys:                     # INNER join customer: unmatched orders disappear
            continue
        ok = int(orders["o_orderkey"][i])
        rows.append({"c_name": customer_name(ck), "c_custkey": ck, "o_orderkey": ok,
                     "o_orderdate": int(orders["o_orderdate"][i]), "o_totalprice": int(orders["o_totalprice"][i]),
                     "sum_qty": qsum[ok]})
    rows.sort(key=lambda r: (-r["o_totalprice"], r["o_orderdate"]))
    return rows if limit is None else rows[:limit]


def q18_text(rows):
    lines = ["#" + "\t" * 5]
    for r in rows:
        tp = r["o_totalprice"]
        lines.append("\t".join([r["c_name"], str(r["c_custkey"]), str(r["o_orderkey"]), fmt_date(r["o_orderdate"]),
                                fmt_decimal((abs(tp), 2, tp < 0), 2), str(r["sum_qty"])]))
    return "\n".join(lines) + "\n"


def q9(part, supplier, partsupp, orders, line, like_word="pink", with_counts=False):
    """TPC-H Q9 (cases/tpch/query/q9.sql): six-way INNER join, `p_name like '%pink%'`, group by
    (n_name, extract(year from o_orderdate)), sum(l_extendedprice*(1-l_discount) - ps_supplycost*l_quantity).
    Typing per the reference binder: the first product is DECIMAL scale 4; l_quantity is cast
    INTEGER -> DECIMAL (value scale 0) so the second product has value scale 2; Sub -> scale 4
    (govalues: max of the scales); sum(DECIMAL) -> DECIMAL(38,4) by sequential Add, exact here (totals
    need 12 digits).  NULL-free inputs; every join is on keys that exist.  Returns [(nation, year, sum at scale 4)]
    ordered by nation, year DESC."""
    like = part["p_name_like"] if part.get("like_word") == like_word else np.array([like_word.encode() in n for n in part["p_name"]], dtype=bool)
    part_ok = np.zeros(int(part["p_partkey"].max()) + 1, dtype=bool)
    part_ok[part["p_partkey"][like]] = True
    m = part_ok[line["l_partkey"]]
    lp, ls = line["l_partkey"][m].astype(np.int64), line["l_suppkey"][m].astype(np.int64)
    ext, disc, qty, lok = line["l_extendedprice"][m], line["l_discount"][m], line["l_quantity"][m].astype(np.int64), line["l_orderkey"][m]
    # partsupp lookup on (partkey, suppkey)
    nsupp = int(supplier["s_suppkey"].max()) + 1
    pskey = partsupp["ps_partkey"].astype(np.int64) * nsupp + partsupp["ps_suppkey"].astype(np.int64)
    order_ps = np.argsort(pskey, kind="stable")
    pos = np.searchsorted(pskey[order_ps], lp * nsupp + ls)
    found = (pos < len(pskey)) & (pskey[order_ps][np.minimum(pos, len(pskey) - 1)] == lp * nsupp + ls)
    cost = partsupp["ps_supplycost"][order_ps][np.minimum(pos, len(pskey) - 1)]
    # supplier -> nation, orders -> year
    nat_of = np.full(nsupp, -1, dtype=np.int64)
    nat_of[supplier["s_suppkey"]] = supplier["s_nationkey"]
    nat = nat_of[ls]
    opos = np.searchsorted(orders["o_orderkey"], lok)
    ofound = (opos < len(orders["o_orderkey"])) & (orders["o_orderkey"][np.minimum(opos, len(orders["o_orderkey"]) - 1)] == lok)
    odate = orders["o_orderdate"][np.minimum(opos, len(orders["o_orderkey"]) - 1)]
    year = (np.datetime64("1970-01-01") + odate.astype("timedelta64[D]")).astype("datetime64[Y]").astype(np.int64) + 1970
    keep = found & ofound & (nat >= 0)
    amount = ext * (100 - disc) - cost * qty * 100          # scale 4, exact int64
    out, cnt = {}, {}
    for n_, y_, a_ in zip(nat[keep], year[keep], amount[keep]):
        k_ = (int(n_), int(y_))
        out[k_] = out.get(k_, 0) + int(a_)
        cnt[k_] = cnt.get(k_, 0) + 1
    names = nation_names()
    rows = [(names[n_], y_, v) + ((cnt[(n_, y_)],) if with_counts else ()) for (n_, y_), v in out.items()]
    rows.sort(key=lambda r: (r[0], -r[1]))
    return rows


def q9_text(rows):
    lines = ["#" + "\t" * 2]
    for nation, year, v in rows:
        lines.append("\t".join([nation, str(year), fmt_decimal((abs(v), 4, v < 0), 4)]))
    return "\n".join(lines) + "\n"


def q3_text(res, limit=10):
    lines = ["#" + "\t" * 3]
    for g in q3_topk(res, limit):
        lines.append("\t".join([str(g["l_orderkey"]), fmt_decimal(g["revenue"], 4),
                                fmt_date(g["o_orderdate"]), str(g["o_shippriority"])]))
    return "\n".join(lines) + "\n"
