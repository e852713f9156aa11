"""GPU parity: hash join build/probe + high-cardinality group-by (TPC-H Q3 shape) vs the oracle,
and the reference's own SF1 golden result files reproduced by the GPU path end to end."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run(plan, tables):
    from plan_b200 import compute as X
    ex = X.gpuPipelineExec(plan, tables)
    ex.Init()
    chunks = X.drain(ex)
    stats, explain = ex.stats, ex.Explain()
    ex.Close()
    return chunks, stats, explain


def _q3_groups(chunks):
    out = {}
    for c in chunks:
        ok, rev, od, sp = (v.Data for v in c.Data)
        for r in range(c.Card()):
            key = (int(ok[r]), int(od[r]), int(sp[r]))
            assert key not in out, "duplicate group emitted"
            x = rev[r]
            out[key] = (-1 if x["neg"] else 1) * int(x["coef"]) * 10 ** (4 - int(x["scale"]))
    return out


def check_q3(oracle, tables, host, check_counts=True, **kw):
    from plan_b200 import tpch as T
    chunks, stats, explain = _run(T.q3_plan(**kw), tables)
    ref = oracle.q3(host["customer"], host["orders"], host["lineitem"], **kw)
    assert all(c.Card() <= 2048 for c in chunks)
    got = _q3_groups(chunks)
    want = {(g["l_orderkey"], g["o_orderdate"], g["o_shippriority"]): g["x_revenue"] for g in ref["groups"]}
    assert len(got) == ref["stats"]["ngroups"] == len(want)
    assert got == want                                   # row set and exact DECIMAL sums
    if check_counts:                                     # per-rank counters when sharded
        assert stats.aux[0] == ref["stats"]["n_line_sel"]    # rows passing the lineitem filter
        assert stats.aux[1] == ref["stats"]["n_line_joined"]
        assert stats.aux[3] == ref["stats"]["n_cust_sel"]
        assert stats.aux[5] == ref["stats"]["n_orders_joined"]
    return chunks, ref


@pytest.fixture(scope="module")
def uploaded(pg, sf01_host):
    from plan_b200 import tpch as T
    t = T.upload_tables(sf01_host)
    yield t
    for x in t.values():
        x.free()


def test_q3_sf01(pg, oracle, uploaded, sf01_host):
    from plan_b200 import compute as X
    chunks, ref = check_q3(oracle, uploaded, sf01_host)
    # top-10 through the host Order/Limit stand-ins == the oracle's text
    rows = X.order_limit(chunks, [(1, True), (2, False)], 10)
    assert X.rows_text(rows, 4) == oracle.q3_text(ref)


@pytest.mark.parametrize("kw", [
    dict(segment="BUILDING"),
    dict(segment="NOSUCHSEGMENT"),                                  # empty build side -> no rows
    dict(odate_lt=8035, ship_gt=8035),                              # no order before 1992-01-01
    dict(odate_lt=8035 + 3000, ship_gt=8035 - 10),                  # every row passes the date filters
    dict(segment="MACHINERY", odate_lt=8035 + 400, ship_gt=8035 + 380),
])
def test_q3_variants(pg, oracle, uploaded, sf01_host, kw):
    if kw.get("segment") == "NOSUCHSEGMENT":
        from plan_b200 import tpch as T
        chunks, stats, _ = _run(T.q3_plan(**kw), uploaded)
        assert chunks == [] and stats.aux[1] == 0
        return
    check_q3(oracle, uploaded, sf01_host, **kw)


def check_q3_topk(oracle, tables, host, limit, **kw):
    """Limit <- Order <- Agg fused on the device == oracle ORDER BY revenue desc, o_orderdate LIMIT k."""
    from plan_b200 import compute as X, tpch as T
    chunks, stats, _ = _run(T.q3_topk_plan(limit=limit, **kw), tables)
    ref = oracle.q3(host["customer"], host["orders"], host["lineitem"], **kw)
    rows = [[v.GetValue(r) for v in c.Data] for c in chunks for r in range(c.Card())]
    assert len(rows) == min(limit, ref["stats"]["ngroups"])
    assert X.rows_text(rows, 4) == oracle.q3_text(ref, limit)     # already ordered and limited by the GPU path
    return stats


@pytest.mark.parametrize("limit", [0, 1, 10, 100, 5000])
def test_q3_topk_pushdown(pg, oracle, uploaded, sf01_host, limit):
    stats = check_q3_topk(oracle, uploaded, sf01_host, limit)
    assert stats.aux[6] > 1000          # the aggregate itself produced many groups; only `limit` came back


def test_q3_topk_other_parameters(pg, oracle, uploaded, sf01_host):
    check_q3_topk(oracle, uploaded, sf01_host, 10, segment="BUILDING", odate_lt=8035 + 3000, ship_gt=8035 - 10)
    check_q3_topk(oracle, uploaded, sf01_host, 10, segment="NOSUCHSEGMENT")


def test_join_duplicate_build_keys_emit_every_pair(pg, oracle, sf01_host):
    """INNER join semantics with a non-unique build side (join_scan.go:182-299 follows the whole
    chain): duplicating every customer doubles every revenue, duplicating orders too -> x4."""
    from plan_b200 import tpch as T
    n_o = 20000
    orders = {k: v[:n_o].copy() for k, v in sf01_host["orders"].items()}
    nl = int(np.searchsorted(sf01_host["lineitem"]["l_orderkey"], orders["o_orderkey"][-1], side="right"))
    line = {k: v[:nl].copy() for k, v in sf01_host["lineitem"].items()}
    cust = {k: np.concatenate([v, v]) for k, v in sf01_host["customer"].items()}
    orders2 = {k: np.concatenate([v, v]) for k, v in orders.items()}
    host = {"customer": cust, "orders": orders2, "lineitem": line}
    t = T.upload_tables(host)
    try:
        kw = dict(odate_lt=8035 + 3000, ship_gt=8035 - 10)
        chunks, ref = check_q3(oracle, t, host, **kw)
        base = oracle.q3(sf01_host["customer"], orders, line, **kw)
        want = {g["l_orderkey"]: 4 * g["x_revenue"] for g in base["groups"]}
        got = {k[0]: v for k, v in _q3_groups(chunks).items()}
        assert got == want
    finally:
        for x in t.values():
            x.free()


def test_reference_golden_files_sf1_on_gpu(pg):
    """Known-answer test: dbgen-equivalent SF1 data generated in HBM, Q1 / Q6 / Q3 / Q18 / Q9 through the
    C ABI, rendered with the reference's formatting rules == the reference's own result files
    (/root/reference/cases/tpch/1g/plan/q{1,6,3,18,9}.txt, committed under tests/golden/)."""
    from plan_b200 import compute as X, tpch as T
    t = T.generate_device_tables(1.0, want=T.ALL_TABLES)
    try:
        assert t["lineitem"].rows() == 6001215
        chunks, _, _ = _run(T.q6_plan(), t)
        assert X.rows_text(X.order_limit(chunks, []), 1) == open(os.path.join(GOLDEN, "ref_sf1_q6.txt")).read()
        chunks, _, _ = _run(T.q1_plan(), t)
        assert X.rows_text(X.order_limit(chunks, [(0, False), (1, False)]), 10) == open(os.path.join(GOLDEN, "ref_sf1_q1.txt")).read()
        chunks, _, _ = _run(T.q3_plan(), t)
        assert X.rows_text(X.order_limit(chunks, [(1, True), (2, False)], 10), 4) == open(os.path.join(GOLDEN, "ref_sf1_q3.txt")).read()
        chunks, _, _ = _run(T.q3_topk_plan(10), t)      # ORDER BY + LIMIT fused into the GPU pipeline
        assert X.rows_text(X.order_limit(chunks, []), 4) == open(os.path.join(GOLDEN, "ref_sf1_q3.txt")).read()
        chunks, _, _ = _run(T.q18_plan(), t)            # cases/tpch/1g/plan/q18.txt
        assert X.rows_text(X.order_limit(chunks, []), 6) == open(os.path.join(GOLDEN, "ref_sf1_q18.txt")).read()
        chunks, _, _ = _run(T.q9_plan(), t)             # cases/tpch/1g/plan/q9.txt (175 groups)
        assert X.rows_text(X.order_limit(chunks, []), 3) == open(os.path.join(GOLDEN, "ref_sf1_q9.txt")).read()
    finally:
        for x in t.values():
            x.free()


@pytest.mark.parametrize("query,golden", [(6, "ref_sf1_q6.txt"), (1, "ref_sf1_q1.txt"), (3, "ref_sf1_q3.txt"), (18, "ref_sf1_q18.txt"),
                                          (9, "ref_sf1_q9.txt")])
def test_cpp_host_shim_reproduces_golden_files(query, golden):
    """The C++ host shim (plan_b200/host: OperatorExec / PhysicalOperator / Chunk mirrors above the C
    ABI, standing in for the Go side) run like `tester tpch1g --query_id N`: its stdout is the
    reference's result file, byte for byte."""
    import subprocess
    from plan_b200 import build as B
    r = subprocess.run([B.HOST_BIN, "1", str(query)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == open(os.path.join(GOLDEN, golden)).read()


def _groupby_result(chunks):
    out = {}
    for c in chunks:
        k, s, n = (v.Data for v in c.Data)
        for r in range(c.Card()):
            sv = s[r]
            if s.dtype.names and "coef" in s.dtype.names:
                val = (-1 if sv["neg"] else 1) * int(sv["coef"])
            else:
                val = (int(sv["upper"]) << 64) + int(sv["lower"])
            assert int(k[r]) not in out
            out[int(k[r])] = (val, int(n[r]["lower"]))
    return out


def check_groupby(oracle, tables, line, **kw):
    from plan_b200 import tpch as T
    chunks, stats, explain = _run(T.groupby_plan(**kw), tables)
    assert "global open-addressing table" in explain
    want = oracle.groupby_sum(line, **kw)
    got = _groupby_result(chunks)
    assert len(got) == len(want)
    assert got == want
    return stats


@pytest.mark.parametrize("kw", [
    dict(key="l_orderkey", value="l_quantity"),                           # 150k groups at SF0.1
    dict(key="l_orderkey", value="l_quantity", having_gt=200),            # Q18's inner aggregate shape
    dict(key="l_partkey", value="l_quantity", ship_le=8035 + 1263),       # 20k groups, many rows per group
    dict(key="l_suppkey", value="l_extendedprice"),                       # sum(DECIMAL), 1000 hot groups
    dict(key="l_orderkey", value="l_quantity", having_gt=10 ** 6),        # HAVING removes everything
])
def test_high_cardinality_groupby(pg, oracle, uploaded, sf01_host, kw):
    check_groupby(oracle, uploaded, sf01_host["lineitem"], **kw)


def test_high_cardinality_groupby_topk(pg, oracle, uploaded, sf01_host):
    from plan_b200 import tpch as T
    chunks, _, _ = _run(T.groupby_plan(key="l_orderkey", value="l_quantity", having_gt=150, topk=20), uploaded)
    want = oracle.groupby_sum(sf01_host["lineitem"], having_gt=150)
    top = sorted(want.items(), key=lambda kv: (-kv[1][0], kv[0]))[:20]
    got = _groupby_result(chunks)
    rows = [(int(c.Data[0].Data[r]),) for c in chunks for r in range(c.Card())]
    assert [k for (k,) in rows] == [k for k, _ in top]          # already ordered: sum desc, key asc
    assert all(got[k] == v for k, v in top)


def check_semi(oracle, tables, host, **kw):
    from plan_b200 import tpch as T
    chunks, stats, explain = _run(T.semi_plan(**kw), tables)
    want = oracle.semi_groupby(host["orders"], host["lineitem"], **kw)
    got = _groupby_result(chunks)
    assert len(got) == len(want) and got == want
    return stats


@pytest.mark.parametrize("kw", [
    dict(anti=False), dict(anti=True),
    dict(anti=False, odate_lt=8035 + 3000, ship_gt=8035 + 2400),     # few build keys
    dict(anti=True, odate_lt=8035 + 3000, ship_gt=8035 + 5000),      # empty build side: ANTI keeps every probe row
    dict(anti=False, odate_lt=8035 + 3000, ship_gt=8035 + 5000),     # empty build side: SEMI keeps nothing
])
def test_semi_and_anti_join(pg, oracle, uploaded, sf01_host, kw):
    """`IN (subquery)` / `NOT IN`: SEMI and ANTI joins decided by the exact key bitmap of the build side."""
    check_semi(oracle, uploaded, sf01_host, **kw)


def test_semi_join_without_bitmap(pg, oracle, uploaded, sf01_host, monkeypatch):
    """Same result through the hash-table probe (generic kernel, no bitmap-only build)."""
    monkeypatch.setenv("PG_JOIN_NO_BITMAP_BUILD", "1")
    monkeypatch.setenv("PG_JOIN_GENERIC", "1")
    check_semi(oracle, uploaded, sf01_host, anti=False)
    check_semi(oracle, uploaded, sf01_host, anti=True)


@pytest.mark.parametrize("anti", [False, True])
def test_null_join_keys_never_match(pg, oracle, sf01_host, anti):
    """NULL keys on either side never match (join_table.go:152-195): dropped from the build side, and a
    probe row with a NULL key survives only an ANTI join."""
    from plan_b200 import compute as X, tpch as T
    rng = np.random.default_rng(11)
    n_o = 40000
    orders = {k: v[:n_o].copy() for k, v in sf01_host["orders"].items()}
    nl = int(np.searchsorted(sf01_host["lineitem"]["l_orderkey"], orders["o_orderkey"][-1], side="right"))
    line = {k: v[:nl].copy() for k, v in sf01_host["lineitem"].items()}
    v_ok, v_lk = rng.random(n_o) >= 0.2, rng.random(nl) >= 0.3
    pack = lambda v: np.packbits(v.astype(np.uint8), bitorder="little")   # noqa: E731
    to = X.DeviceTable.create("orders", T.ORDERS)
    to.append([orders[c[0]] for c in T.ORDERS], valid=[pack(v_ok) if c[0] == "o_orderkey" else None for c in T.ORDERS])
    to.seal()
    tl = X.DeviceTable.create("lineitem", T.LINEITEM)
    tl.append([line[c[0]] for c in T.LINEITEM], valid=[pack(v_lk) if c[0] == "l_orderkey" else None for c in T.LINEITEM])
    tl.seal()
    try:
        kw = dict(anti=anti, odate_lt=8035 + 3000, ship_gt=8035 + 900)
        chunks, _, _ = _run(T.semi_plan(**kw), {"orders": to, "lineitem": tl})
        want = oracle.semi_groupby(orders, line, valid_okey=v_ok, valid_lkey=v_lk, **kw)
        assert _groupby_result(chunks) == want
    finally:
        to.free()
        tl.free()


@pytest.mark.parametrize("empty", ["customer", "orders", "lineitem", "all"])
def test_q3_with_empty_tables(pg, oracle, sf01_host, empty):
    """Empty inputs on any side of the join tree: no rows, no crash (the reference emits nothing)."""
    from plan_b200 import tpch as T
    host = {}
    for name, cols in sf01_host.items():
        n = 0 if empty in (name, "all") else 3000
        host[name] = {k: v[:n].copy() for k, v in cols.items()}
    t = T.upload_tables(host)
    try:
        for plan in (T.q3_plan(odate_lt=8035 + 3000, ship_gt=8035 - 10), T.q3_topk_plan(10, odate_lt=8035 + 3000, ship_gt=8035 - 10)):
            chunks, stats, _ = _run(plan, t)
            if empty == "lineitem" or empty == "all" or empty == "orders" or empty == "customer":
                ref = oracle.q3(host["customer"], host["orders"], host["lineitem"], odate_lt=8035 + 3000, ship_gt=8035 - 10) \
                    if len(host["orders"]["o_orderkey"]) and len(host["customer"]["c_custkey"]) and len(host["lineitem"]["l_orderkey"]) else None
                assert ref is None
                assert chunks == []
        chunks, _, _ = _run(T.groupby_plan(key="l_orderkey", value="l_quantity"), t)
        assert (chunks == []) == (len(host["lineitem"]["l_orderkey"]) == 0)
    finally:
        for x in t.values():
            x.free()


def test_sentinel_key_values_are_refused(pg, sf01_host):
    """A key column whose value range contains the table's empty-slot sentinel cannot use the open
    addressing tables: refused at plan time (PG_EUNSUPPORTED), never a wrong group."""
    from plan_b200 import _lib as L, compute as X, tpch as T
    line = {k: v[:5000].copy() for k, v in sf01_host["lineitem"].items()}
    line["l_orderkey"][17] = np.int64(-0x7f7f7f7f7f7f7f80)        # == HT_EMPTY (0x8080808080808080)
    line["l_orderkey"][18] = np.int64(2 ** 62)
    t = T.upload_tables({"lineitem": line})
    try:
        ex = X.gpuPipelineExec(T.groupby_plan(key="l_orderkey", value="l_quantity"), t)
        with pytest.raises(L.PlanGpuError) as ei:
            ex.Init()
        assert ei.value.status == L.PG_EUNSUPPORTED
        ex.Close()
    finally:
        t["lineitem"].free()


def test_extreme_key_values_group_correctly(pg, oracle, sf01_host):
    """Keys spread over almost the whole int64 range (but not the sentinel): hashed slots, exact sums."""
    from plan_b200 import tpch as T
    n = 20000
    line = {k: v[:n].copy() for k, v in sf01_host["lineitem"].items()}
    rng = np.random.default_rng(3)
    keys = rng.integers(-2 ** 62, 2 ** 62, size=500, dtype=np.int64)
    line["l_orderkey"] = keys[rng.integers(0, 500, size=n)]
    t = T.upload_tables({"lineitem": line})
    try:
        check_groupby(oracle, t, line, key="l_orderkey", value="l_extendedprice")
    finally:
        t["lineitem"].free()


def _sorted_run_case(name, n, rng):
    """Key layouts that stress the sorted-run reduce-by-key: edges of 4-row vectors, 1024-row tiles, block chunks."""
    if name == "one_run":
        return np.full(n, 7, dtype=np.int64)
    if name == "all_distinct":
        return np.arange(n, dtype=np.int64) * 3 - 5
    if name == "tile_aligned":                      # every run is exactly one 1024-row tile
        return np.arange(n, dtype=np.int64) // 1024
    if name == "vector_aligned":                    # every run is exactly one 4-row vector
        return np.arange(n, dtype=np.int64) // 4
    if name == "long_then_short":                   # one run over many tiles, then runs of 1..7
        head = n // 2 + 3
        tail = np.repeat(np.arange(1, n), rng.integers(1, 8, size=n - 1))[: n - head]
        return np.concatenate([np.zeros(head, dtype=np.int64), tail.astype(np.int64)])[:n]
    if name == "mixed":
        lens = rng.choice([1, 2, 3, 5, 31, 257, 1023, 1025, 4099], size=n)
        return np.repeat(np.arange(n, dtype=np.int64) * 2, lens)[:n]
    raise AssertionError(name)


@pytest.mark.parametrize("name", ["one_run", "all_distinct", "tile_aligned", "vector_aligned", "long_then_short", "mixed"])
@pytest.mark.parametrize("n", [1, 5, 1024, 40000, 300001])
def test_sorted_run_groupby_edges(pg, oracle, sf01_host, monkeypatch, name, n):
    """Sorted group key => fused reduce-by-key (run_group_kernel + run_fixup_kernel): exact against the oracle,
    with and without a predicate that empties whole runs, with and without HAVING, and identical to the
    table-based path it replaces."""
    from plan_b200 import tpch as T
    rng = np.random.default_rng(n * 31 + len(name))
    line = {k: v[:n].copy() for k, v in sf01_host["lineitem"].items()}
    line["l_orderkey"] = _sorted_run_case(name, n, rng)
    assert len(line["l_orderkey"]) == n
    t = T.upload_tables({"lineitem": line})
    try:
        for extra in (dict(), dict(ship_le=8035 + 1200), dict(having_gt=100), dict(value="l_extendedprice", ship_le=8035 + 600)):
            kw = dict(key="l_orderkey", value="l_quantity")
            kw.update(extra)
            chunks, stats, explain = _run(T.groupby_plan(**kw), t)
            assert "sorted-run" in explain
            got = _groupby_result(chunks)
            assert got == oracle.groupby_sum(line, **kw)
            monkeypatch.setenv("PG_NO_SORTED_RUNS", "1")
            chunks2, _, _ = _run(T.groupby_plan(**kw), t)
            monkeypatch.delenv("PG_NO_SORTED_RUNS")
            assert _groupby_result(chunks2) == got
    finally:
        t["lineitem"].free()


def test_full_size_sf100_matches_the_oracle_fixtures(pg):
    """BASELINE's full size: SF100 (600,037,902 lineitem rows) generated in HBM, Q6 / Q1 / Q3(top 10)
    through the C ABI == the CPU oracle's SF100 results (tests/golden/oracle_sf100_*.txt, produced once by
    tests/golden/make_sf100_fixtures.py on the CPU: 10 minutes of sequential Decimal folds).  Q1's
    sum_charge for (N,O) exceeds 19 digits here, so this pins the order-dependent rounding emulation at
    the headline scale."""
    from plan_b200 import compute as X, tpch as T
    import torch
    if torch.cuda.mem_get_info(0)[0] < 60 * 2 ** 30:
        pytest.skip("needs ~60 GB of free HBM")
    t = T.generate_device_tables(100.0, want=T.ALL_TABLES)
    try:
        assert t["lineitem"].rows() == 600037902
        chunks, _, _ = _run(T.q6_plan(), t)
        assert X.rows_text(X.order_limit(chunks, []), 1) == open(os.path.join(GOLDEN, "oracle_sf100_q6.txt")).read()
        chunks, stats, _ = _run(T.q1_plan(), t)
        assert stats.aux[1] == 1                     # exactly one (group, sum) took the ordered-rounding path
        assert X.rows_text(X.order_limit(chunks, [(0, False), (1, False)]), 10) == open(os.path.join(GOLDEN, "oracle_sf100_q1.txt")).read()
        chunks, _, _ = _run(T.q3_topk_plan(10), t)
        assert X.rows_text(X.order_limit(chunks, []), 4) == open(os.path.join(GOLDEN, "oracle_sf100_q3.txt")).read()
        chunks, _, _ = _run(T.q18_plan(), t)           # tests/golden/make_sf100_q18_fixture.py
        assert X.rows_text(X.order_limit(chunks, []), 6) == open(os.path.join(GOLDEN, "oracle_sf100_q18.txt")).read()
        chunks, _, _ = _run(T.q9_plan(), t)            # tests/golden/make_sf100_q9_fixture.py
        assert X.rows_text(X.order_limit(chunks, []), 3) == open(os.path.join(GOLDEN, "oracle_sf100_q9.txt")).read()
        # size-independent property at full size: the sorted-run group-by and the table-based one agree
        gb = T.groupby_plan(key="l_orderkey", value="l_quantity", having_gt=300)
        a, _, _ = _run(gb, t)
        os.environ["PG_NO_SORTED_RUNS"] = "1"
        try:
            b, _, _ = _run(gb, t)
        finally:
            del os.environ["PG_NO_SORTED_RUNS"]
        assert _groupby_result(a) == _groupby_result(b) and len(_groupby_result(a)) > 1000
    finally:
        for x in t.values():
            x.free()


def _q18_rows(chunks):
    rows = []
    for c in chunks:
        name, ck, ok, od, tp, sq = (v.Data for v in c.Data)
        for r in range(c.Card()):
            rows.append({"c_name": name[r].decode(), "c_custkey": int(ck[r]), "o_orderkey": int(ok[r]), "o_orderdate": int(od[r]),
                         "o_totalprice": int(tp[r]), "sum_qty": (int(sq[r]["upper"]) << 64) + int(sq[r]["lower"])})
    return rows


@pytest.mark.parametrize("qty_gt,limit", [(250, 100), (200, 100), (300, 5), (100, None), (10 ** 6, 100)])
def test_q18_sf01(pg, oracle, uploaded, sf01_host, qty_gt, limit):
    """TPC-H Q18's shape: SEMI join against a HAVING-filtered sub-aggregate pushed down to the orders scan,
    INNER joins on unique keys, five group keys functionally dependent on the order row (one of them a
    VARCHAR), ORDER BY a DECIMAL key DESC + date, LIMIT -- against the oracle at several thresholds."""
    from plan_b200 import tpch as T
    chunks, stats, explain = _run(T.q18_plan(qty_gt=qty_gt, limit=limit), uploaded)
    assert "aggregate over lineitem" in explain and "dependent keys" in explain
    want = oracle.q18(sf01_host["customer"], sf01_host["orders"], sf01_host["lineitem"], qty_gt=qty_gt, limit=limit)
    got = _q18_rows(chunks)
    if limit is None:       # unordered plan root: compare as sets
        key = lambda r: r["o_orderkey"]   # noqa: E731
        assert sorted(got, key=key) == sorted(want, key=key)
    else:
        assert got == want


def _q9_rows(chunks):
    rows = []
    for c in chunks:
        for r in range(c.Card()):
            nat, yr, sm = c.Data[0], c.Data[1], c.Data[2].Data[r]
            rows.append((nat.Dict[int(nat.Data[r])], int(yr.Data[r]), (-1 if sm["neg"] else 1) * int(sm["coef"])))
    return rows


def test_device_generator_q9_tables(pg, oracle):
    """part / supplier / partsupp / nation generated in HBM == the dbgen-exact CPU generator; p_name is checked
    through GPU LIKE counts (VARCHAR columns are not readable as flat arrays)."""
    from plan_b200 import tpch as T
    sf = 0.05
    t = T.generate_device_tables(sf, want=("part", "supplier", "partsupp", "nation"))
    try:
        sup, ps = oracle.gen_supplier(sf), oracle.gen_partsupp(sf)
        for k, v in sup.items():
            assert np.array_equal(t["supplier"].read_column(k), v), k
        for k, v in ps.items():
            assert np.array_equal(t["partsupp"].read_column(k), v), k
        assert list(t["nation"].read_column("n_nationkey")) == list(range(25))
        for word in ("pink", "green", "almond", "yellow", "zzz"):
            part = oracle.gen_part(sf, word)
            chunks, _, _ = _run(T.part_like_plan("%" + word + "%"), t)
            want = part["p_partkey"][part["p_name_like"]]
            if len(want) == 0:
                assert chunks == []
            else:
                h = lambda v: (int(v["upper"]) << 64) + int(v["lower"])   # noqa: E731
                assert (h(chunks[0].Data[0].Data[0]), h(chunks[0].Data[1].Data[0])) == (len(want), int(want.astype(np.int64).sum()))
    finally:
        for x in t.values():
            x.free()


@pytest.mark.parametrize("word", ["pink", "green", "lace", "nosuchcolour"])
def test_q9_sf005(pg, oracle, word):
    """TPC-H Q9's shape (star join: five INNER joins on the lineitem spine, a two-column key, LIKE on the build
    side, EXTRACT(year) group key, difference of products) against the oracle."""
    from plan_b200 import tpch as T
    sf = 0.05
    t = T.generate_device_tables(sf, want=T.ALL_TABLES)
    try:
        chunks, stats, explain = _run(T.q9_plan(word), t)
        assert "StarJoin" in explain
        orders, line = oracle.gen_orders_lineitem(sf)
        want = oracle.q9(oracle.gen_part(sf, word), oracle.gen_supplier(sf), oracle.gen_partsupp(sf), orders, line, like_word=word)
        assert _q9_rows(chunks) == want
    finally:
        for x in t.values():
            x.free()


@pytest.mark.parametrize("word", ["pink", "lace", "nosuchcolour"])
def test_q9_through_the_row_exchange(pg, oracle, word, monkeypatch):
    """PG_FORCE_EXCHANGE=1: the lineitem x partsupp join runs through the all-to-all row exchange (exchange.cuh: both
    sides hash-partitioned on (partkey, suppkey), records packed, scattered into destination order, joined from the
    receive buffers) even on one rank, where the exchange degenerates to a device copy.  Same answer as the local join."""
    from plan_b200 import tpch as T
    sf = 0.05
    t = T.generate_device_tables(sf, want=T.ALL_TABLES)
    try:
        monkeypatch.setenv("PG_FORCE_EXCHANGE", "1")
        chunks, stats, explain = _run(T.q9_plan(word), t)
        assert "partsupp(2-col key, payload, ROW EXCHANGE" in explain
        orders, line = oracle.gen_orders_lineitem(sf)
        want = oracle.q9(oracle.gen_part(sf, word), oracle.gen_supplier(sf), oracle.gen_partsupp(sf), orders, line, like_word=word)
        assert _q9_rows(chunks) == want
        monkeypatch.setenv("PG_FORCE_EXCHANGE", "0")
        chunks2, _, explain2 = _run(T.q9_plan(word), t)
        assert "ROW EXCHANGE" not in explain2 and _q9_rows(chunks2) == want
    finally:
        for x in t.values():
            x.free()


def test_q9_row_exchange_reproduces_the_reference_golden_file(pg, monkeypatch):
    """cases/tpch/1g/plan/q9.txt through the exchanged join (SF1, device-generated tables)."""
    from plan_b200 import compute as X, tpch as T
    t = T.generate_device_tables(1.0, want=T.ALL_TABLES)
    try:
        monkeypatch.setenv("PG_FORCE_EXCHANGE", "1")
        chunks, _, explain = _run(T.q9_plan(), t)
        assert "ROW EXCHANGE" in explain
        assert X.rows_text(X.order_limit(chunks, []), 3) == open(os.path.join(GOLDEN, "ref_sf1_q9.txt")).read()
    finally:
        for x in t.values():
            x.free()


def test_row_exchange_refuses_duplicate_build_keys(pg, oracle, monkeypatch):
    """A probe row that matches two received build rows needs row multiplication in the sink: refused, never wrong."""
    from plan_b200 import _lib as L, compute as X, tpch as T
    host, part = _q9_host(oracle, 0.01)
    host["partsupp"] = {k: np.concatenate([v, v[:50]]) for k, v in host["partsupp"].items()}
    t = T.upload_tables(host)
    try:
        monkeypatch.setenv("PG_FORCE_EXCHANGE", "1")
        ex = X.gpuPipelineExec(T.q9_plan("e"), t)       # '%e%' keeps nearly every part: the doubled keys are probed
        ex.Init()
        assert "ROW EXCHANGE" in ex.Explain()
        with pytest.raises(L.PlanGpuError) as ei:
            X.drain(ex)
        assert ei.value.status == L.PG_EUNSUPPORTED and "more than one build row" in str(ei.value)
        ex.Close()
    finally:
        for x in t.values():
            x.free()


def _q9_host(oracle, sf, word="pink"):
    orders, line = oracle.gen_orders_lineitem(sf)
    part = oracle.gen_part(sf, word)
    host = {"lineitem": line, "orders": orders, "part": {"p_partkey": part["p_partkey"], "p_name": part["p_name"]},
            "supplier": oracle.gen_supplier(sf), "partsupp": oracle.gen_partsupp(sf),
            "nation": {"n_nationkey": np.arange(25, dtype=np.int32), "n_name": np.arange(25, dtype=np.uint8)}}
    return host, part


@pytest.mark.parametrize("empty", ["lineitem", "part", "partsupp", "orders", "nation"])
def test_q9_with_an_empty_table(pg, oracle, empty):
    """A star join with an empty fact table or an empty build side yields no group (INNER joins), not an error."""
    from plan_b200 import tpch as T
    host, _ = _q9_host(oracle, 0.01)
    host[empty] = {k: v[:0] for k, v in host[empty].items()}
    t = T.upload_tables(host)
    try:
        chunks, _, explain = _run(T.q9_plan(), t)
        assert "StarJoin" in explain and chunks == []
    finally:
        for x in t.values():
            x.free()


def test_q9_uploaded_tables_and_duplicate_build_keys(pg, oracle):
    """Host-uploaded (not device-generated) tables give the oracle's result; a build side with duplicate keys would need
    row multiplication in the star sink and is refused at execution time with PG_EUNSUPPORTED -- never answered wrongly."""
    from plan_b200 import _lib as L, compute as X, tpch as T
    sf = 0.01
    host, part = _q9_host(oracle, sf)
    t = T.upload_tables(host)
    try:
        chunks, _, _ = _run(T.q9_plan(), t)
        assert _q9_rows(chunks) == oracle.q9(part, host["supplier"], host["partsupp"], host["orders"], host["lineitem"])
    finally:
        for x in t.values():
            x.free()
    host["supplier"] = {k: np.concatenate([v, v[:10]]) for k, v in host["supplier"].items()}     # ten supplier keys twice
    t = T.upload_tables(host)
    try:
        ex = X.gpuPipelineExec(T.q9_plan(), t)
        ex.Init()
        with pytest.raises(L.PlanGpuError) as ei:
            X.drain(ex)
        assert ei.value.status == L.PG_EUNSUPPORTED and "more than one build row" in str(ei.value)
        ex.Close()
    finally:
        for x in t.values():
            x.free()


def test_q18_with_empty_tables_and_no_qualifying_order(pg, oracle, sf01_host):
    from plan_b200 import tpch as T
    for empty in ("customer", "orders", "lineitem"):
        host = {k: dict(v) for k, v in sf01_host.items()}
        host[empty] = {k: v[:0] for k, v in host[empty].items()}
        t = T.upload_tables(host)
        try:
            chunks, _, _ = _run(T.q18_plan(qty_gt=200), t)
            assert chunks == []
        finally:
            for x in t.values():
                x.free()


@pytest.mark.parametrize("negated", [False, True])
@pytest.mark.parametrize("kw", [dict(), dict(odate_lt=8035 + 3000, ship_gt=8035 + 2400)])
def test_exists_as_mark_join(pg, oracle, uploaded, sf01_host, negated, kw):
    """EXISTS / NOT EXISTS: Filter(mark = true|false) <- MARK / AntiMARK join (builder_plan.go:380-429,
    join_scan.go:124-165) gives the rows of the SEMI / ANTI join."""
    from plan_b200 import tpch as T
    chunks, _, _ = _run(T.exists_plan(negated=negated, **kw), uploaded)
    want = oracle.semi_groupby(sf01_host["orders"], sf01_host["lineitem"], anti=negated, **kw)
    got = _groupby_result(chunks)
    assert len(got) == len(want) and got == want


def test_mark_join_without_its_filter_is_refused(pg, uploaded):
    from plan_b200 import _lib as L, compute as X, tpch as T
    plan = T.exists_plan()
    plan.Children[0] = plan.Children[0].Children[0]          # drop Filter(mark = true): the mark column would have to be produced
    ex = X.gpuPipelineExec(plan, uploaded)
    with pytest.raises(L.PlanGpuError) as ei:
        ex.Init()
    assert ei.value.status == L.PG_EUNSUPPORTED
    ex.Close()


@pytest.mark.parametrize("key,value", [("l_suppkey", "l_extendedprice"), ("l_partkey", "l_quantity"), ("l_orderkey", "l_discount")])
def test_high_cardinality_avg(pg, oracle, uploaded, sf01_host, key, value):
    """avg in the global-table aggregate: avg(DECIMAL) = sum.Quo(count) bit for bit (govalues Quo restated in
    hostdec.hpp vs the oracle's decimal.h), avg(INTEGER) = float64(sum) / float64(count) exactly."""
    from plan_b200 import tpch as T
    import ctypes as C
    chunks, _, _ = _run(T.groupby_avg_plan(key=key, value=value), uploaded)
    want = oracle.groupby_sum(sf01_host["lineitem"], key=key, value=value)
    L = oracle.lib()
    n = 0
    for c in chunks:
        k, a, cnt = (v.Data for v in c.Data)
        for r in range(c.Card()):
            s_, c_ = want[int(k[r])]
            assert int(cnt[r]["lower"]) == c_
            if value == "l_quantity":
                assert float(a[r]) == float(s_) / float(c_)
            else:
                oc, os_, on = C.c_uint64(), C.c_int(), C.c_int()
                assert L.orc_dec_quo(C.c_uint64(abs(s_)), 2, int(s_ < 0), C.c_uint64(c_), 0, 0, C.byref(oc), C.byref(os_), C.byref(on)) == 0
                assert (int(a[r]["coef"]), int(a[r]["scale"]), int(a[r]["neg"])) == (oc.value, os_.value, on.value)
            n += 1
    assert n == len(want)


@pytest.mark.parametrize("kw", [dict(segment_in=["BUILDING", "MACHINERY"]), dict(segment_ne="HOUSEHOLD"), dict(segment_in=["NOPE"])])
def test_q3_with_code_set_filter_on_the_build_side(pg, oracle, uploaded, sf01_host, kw):
    """IN / <> on a dictionary column of a join's build-side scan (a code set evaluated by pipeline_kernel).
    The oracle's Q3 filters on ONE segment code, so the expected result is Q3 over a customer table whose
    segment codes are rewritten to 0 = selected / 1 = not selected."""
    from plan_b200 import tpch as T
    chunks, _, _ = _run(T.q3_plan(**kw), uploaded)
    seg = sf01_host["customer"]["c_mktsegment"]
    if "segment_in" in kw:
        sel = np.isin(seg, [T.SEGMENTS.index(s) for s in kw["segment_in"] if s in T.SEGMENTS])
    else:
        sel = seg != T.SEGMENTS.index(kw["segment_ne"])
    cust = dict(sf01_host["customer"])
    cust["c_mktsegment"] = np.where(sel, 0, 1).astype(np.uint8)
    ref = oracle.q3(cust, sf01_host["orders"], sf01_host["lineitem"], segment=0)
    got = _q3_groups(chunks)
    want = {(g["l_orderkey"], g["o_orderdate"], g["o_shippriority"]): g["x_revenue"] for g in ref["groups"]}
    assert got == want and len(got) == ref["stats"]["ngroups"]


def test_baseline_sf10_configs_match_the_oracle_fixtures(pg):
    """BASELINE.json configs[1] and configs[2]: Q6 and Q3 (top 10) at SF10 (59,986,052 lineitem rows) on one B200,
    plus Q1, against the CPU oracle's SF10 results (tests/golden/oracle_sf10_*.txt, `make_sf100_fixtures.py 10`)."""
    from plan_b200 import compute as X, tpch as T
    t = T.generate_device_tables(10.0)
    try:
        assert t["lineitem"].rows() == 59986052
        chunks, _, _ = _run(T.q6_plan(), t)
        assert X.rows_text(X.order_limit(chunks, []), 1) == open(os.path.join(GOLDEN, "oracle_sf10_q6.txt")).read()
        chunks, _, _ = _run(T.q3_topk_plan(10), t)
        assert X.rows_text(X.order_limit(chunks, []), 4) == open(os.path.join(GOLDEN, "oracle_sf10_q3.txt")).read()
        chunks, _, _ = _run(T.q1_plan(), t)
        assert X.rows_text(X.order_limit(chunks, [(0, False), (1, False)]), 10) == open(os.path.join(GOLDEN, "oracle_sf10_q1.txt")).read()
    finally:
        for x in t.values():
            x.free()


def test_q9_with_an_average(pg, oracle):
    """avg(DECIMAL) in the star sink = sum.Quo(count), bit for bit against the oracle's decimal restatement."""
    from plan_b200 import tpch as T
    import ctypes as C
    sf = 0.05
    t = T.generate_device_tables(sf, want=T.ALL_TABLES)
    try:
        chunks, _, _ = _run(T.q9_plan(agg="avg"), t)
        orders, line = oracle.gen_orders_lineitem(sf)
        want = oracle.q9(oracle.gen_part(sf, "pink"), oracle.gen_supplier(sf), oracle.gen_partsupp(sf), orders, line, with_counts=True)
        got = []
        for c in chunks:
            for r in range(c.Card()):
                a = c.Data[2].Data[r]
                got.append((c.Data[0].Dict[int(c.Data[0].Data[r])], int(c.Data[1].Data[r]), (int(a["coef"]), int(a["scale"]), int(a["neg"])),
                            int(c.Data[3].Data[r]["lower"])))
        L = oracle.lib()
        exp = []
        for nation, year, s_, n_ in want:
            oc, os_, on = C.c_uint64(), C.c_int(), C.c_int()
            assert L.orc_dec_quo(C.c_uint64(abs(s_)), 4, int(s_ < 0), C.c_uint64(n_), 0, 0, C.byref(oc), C.byref(os_), C.byref(on)) == 0
            exp.append((nation, year, (oc.value, os_.value, on.value), n_))
        assert got == exp
    finally:
        for x in t.values():
            x.free()
