"""GPU parity of the ROW-EMITTING operators (plan_b200/csrc/rows.cu + rowvm.cuh) against the tree-walking oracle
(oracle/rowexec.py), through the C ABI.  Reference: executor_filter.go:12-118, executor_project.go:24-82,
executor_join.go:62-123, join_scan.go:67-299, executeCase expr_exec.go:144-246, binDecimalDivOp
function_operator_binary.go:195-210.  Rows are compared as sorted text (the reference's own Value.String rendering);
row ORDER is not part of the contract of these operators."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    import __graft_entry__ as G
    G.build()
    from plan_b200 import _lib as L
    L.check(L.lib().pg_init(0))
    return L


@pytest.fixture(scope="module")
def data(pg):
    """small TPC-H tables on the host (oracle generator) with NULLs injected, uploaded through pg_table_append"""
    from oracle import oracle as O, rowexec as R
    from plan_b200 import compute as X, tpch as T
    sf = 0.002
    orders, line = O.gen_orders_lineitem(sf)
    cust = O.gen_customer(sf)
    rng = np.random.default_rng(11)
    host = {"lineitem": line, "orders": orders, "customer": cust}
    valid = {"lineitem": {"l_quantity": rng.random(len(line["l_orderkey"])) >= 0.1, "l_discount": rng.random(len(line["l_orderkey"])) >= 0.05,
                          "l_orderkey": rng.random(len(line["l_orderkey"])) >= 0.02},
             "orders": {"o_custkey": rng.random(len(orders["o_orderkey"])) >= 0.05},
             "customer": {}}
    # some orders and customers removed so that LEFT / ANTI / MARK joins have unmatched rows on both sides
    keep_o = rng.random(len(orders["o_orderkey"])) >= 0.3
    host["orders"] = {k: v[keep_o] for k, v in orders.items()}
    valid["orders"] = {k: v[keep_o] for k, v in valid["orders"].items()}
    keep_c = rng.random(len(cust["c_custkey"])) >= 0.4
    host["customer"] = {k: v[keep_c] for k, v in cust.items()}
    schemas = {"lineitem": T.LINEITEM, "orders": T.ORDERS, "customer": T.CUSTOMER}
    tables, rows = {}, {}
    for name, cols in host.items():
        sch = [c for c in schemas[name] if c[0] in cols]
        t = X.DeviceTable.create(name, sch)
        vl = [(np.packbits(valid[name][c[0]].astype(np.uint8), bitorder="little") if c[0] in valid[name] else None) for c in sch]
        t.append([cols[c[0]] for c in sch], valid=vl)
        t.seal(0)
        tables[name] = t
        rows[name] = R.table_rows(cols, sch, valid[name])
    sub = T.Schema(**{n: t.columns for n, t in tables.items()})
    yield tables, rows, sub
    for t in tables.values():
        t.free()


def _check(op, tables, rows, expect_explain="Rows["):
    from oracle import rowexec as R
    from plan_b200 import compute as X
    ex = X.gpuPipelineExec(op, tables)
    ex.Init()
    assert expect_explain in ex.Explain(), ex.Explain()
    chunks = X.drain(ex)
    assert all(c.Card() <= 2048 for c in chunks)
    got = sorted("\t".join(v.GetValue(r).String() for v in c.Data) for c in chunks for r in range(c.Card()))
    ex.Close()
    types = [o.DataTyp for o in ex._agg_op().Outputs]
    want = R.format_rows(R.execute(op, rows), types)
    assert len(got) == len(want), (len(got), len(want))
    assert got == want
    return got


def _ops():
    from plan_b200 import chunk as K, compute as X
    return K, X, K.LType(K.LTID_BOOLEAN)


def test_filter_project_over_a_scan(pg, data):
    """Project <- Filter <- Scan: general boolean filters (OR, IN and <> on INTEGER, NOT, NULL operands) and projected
    arithmetic: DECIMAL * (1 - DECIMAL), DECIMAL / DECIMAL (govalues Quo), INTEGER + INTEGER, EXTRACT(year)."""
    tables, rows, S = data
    K, X, B = _ops()
    lc = lambda n: S.col("lineitem", n)   # noqa: E731
    I, D152 = K.IntegerType(), K.DecimalType(15, 2)
    scan = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("lineitem"),
                              Filters=[X.func("or", B, X.func("in", B, lc("l_quantity"), X.const(3, I), X.const(17, I), X.const(44, I)),
                                              X.func("and", B, X.func("<>", B, lc("l_linenumber"), X.const(1, I)),
                                                     X.func("not", B, X.func(">=", B, lc("l_discount"), X.const(3, D152)))))])
    flt = X.PhysicalOperator(X.POT_Filter, Children=[scan],
                             Filters=[X.func("<", B, lc("l_shipdate"), X.const(9500, K.DateType()))])
    one = X.cast(X.const(1, I), D152)
    outs = [lc("l_orderkey"), lc("l_returnflag"),
            X.func("*", K.DecimalType(18, 4), X.cast(lc("l_extendedprice"), K.DecimalType(16, 2)), X.func("-", K.DecimalType(16, 2), one, lc("l_discount"))),
            X.func("/", K.DecimalType(38, 6), lc("l_extendedprice"), X.func("+", K.DecimalType(16, 2), one, lc("l_tax"))),
            X.func("+", I, lc("l_quantity"), lc("l_linenumber")),
            X.func("extract", I, X.const("year", K.VarcharType()), lc("l_shipdate")),
            X.func("<=", B, lc("l_commitdate"), lc("l_receiptdate"))]
    op = X.PhysicalOperator(X.POT_Project, Outputs=outs, Children=[flt])
    got = _check(op, tables, rows)
    assert len(got) > 100 and any("NULL" in r for r in got)


def test_case_expressions(pg, data):
    """CASE WHEN ... THEN ... ELSE (Q12 / Q14 style), a CASE without ELSE (NULL), and a division by zero sitting in a
    branch that is never taken for the rows that would trip it."""
    tables, rows, S = data
    K, X, B = _ops()
    lc = lambda n: S.col("lineitem", n)   # noqa: E731
    I, D152, V = K.IntegerType(), K.DecimalType(15, 2), K.VarcharType()
    scan = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("lineitem"), Filters=[X.func("<", B, lc("l_linenumber"), X.const(4, I))])
    zero = X.const(0, D152)
    promo = X.func("case", K.DecimalType(18, 4), X.cast(zero, K.DecimalType(18, 4)),
                   X.func("=", B, lc("l_returnflag"), X.const("R", V)),
                   X.func("*", K.DecimalType(18, 4), lc("l_extendedprice"), lc("l_discount")))
    bucket = X.func("case", I, X.const(None, I),
                    X.func("<", B, lc("l_quantity"), X.const(10, I)), X.const(1, I),
                    X.func("<", B, lc("l_quantity"), X.const(30, I)), X.const(2, I))
    safe_div = X.func("case", K.DecimalType(38, 6), X.cast(X.const(-1, I), K.DecimalType(38, 6)),
                      X.func(">", B, lc("l_discount"), zero), X.func("/", K.DecimalType(38, 6), lc("l_tax"), lc("l_discount")))
    op = X.PhysicalOperator(X.POT_Project, Outputs=[lc("l_orderkey"), lc("l_linenumber"), promo, bucket, safe_div], Children=[scan])
    _check(op, tables, rows)


def test_division_by_zero_is_an_error(pg, data):
    """the reference panics in binDecimalDivOp and the query fails: an error status, never a made-up value"""
    tables, rows, S = data
    K, X, B = _ops()
    lc = lambda n: S.col("lineitem", n)   # noqa: E731
    scan = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("lineitem"))
    op = X.PhysicalOperator(X.POT_Project, Outputs=[X.func("/", K.DecimalType(38, 6), lc("l_tax"), lc("l_discount"))], Children=[scan])
    ex = X.gpuPipelineExec(op, tables)
    ex.Init()
    with pytest.raises(pg.PlanGpuError) as ei:
        X.drain(ex)
    assert "division by zero" in str(ei.value)
    ex.Close()


def _join(X, K, B, S, jointype, outs, probe="orders", build="customer", pk="o_custkey", bk="c_custkey", pfilters=None, bfilters=None):
    p = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo(probe), Filters=pfilters or [])
    b = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo(build), Filters=bfilters or [])
    return X.PhysicalOperator(X.POT_Join, Children=[p, b], Outputs=outs,
                              Info=X.JoinOpInfo(jointype, [X.func("=", B, S.col(probe, pk, 0), S.col(build, bk, 1))]))


@pytest.mark.parametrize("jt", ["inner", "left", "semi", "anti", "mark"])
def test_row_emitting_joins(pg, data, jt):
    """orders x customer on a nullable probe key, build side filtered, unmatched rows on both sides; a VARCHAR column
    (c_name) and a dictionary column (c_mktsegment) carried from the build side; the MARK column projected."""
    tables, rows, S = data
    K, X, B = _ops()
    OI, CI = S.idx["orders"], S.idx["customer"]
    V = K.VarcharType()
    jtype = {"inner": X.JOIN_INNER, "left": X.JOIN_LEFT, "semi": X.JOIN_SEMI, "anti": X.JOIN_ANTI, "mark": X.JOIN_MARK}[jt]
    outs = [X.col(0, OI["o_orderkey"], K.BigintType()), X.col(0, OI["o_custkey"], K.IntegerType()), X.col(0, OI["o_orderdate"], K.DateType()),
            X.col(0, OI["o_totalprice"], K.DecimalType(15, 2))]
    if jt in ("inner", "left"):
        outs += [X.col(1, CI["c_name"], V), X.col(1, CI["c_mktsegment"], V), X.col(1, CI["c_nationkey"], K.IntegerType())]
    if jt == "mark":
        outs += [X.col(2, 0, B)]
    bf = [X.func("<>", B, S.col("customer", "c_mktsegment"), X.const("BUILDING", V))]
    pf = [X.func(">=", B, S.col("orders", "o_orderdate"), X.const(8500, K.DateType()))]
    op = _join(X, K, B, S, jtype, outs, pfilters=pf, bfilters=bf)
    got = _check(op, tables, rows)
    assert len(got) > 50
    if jt == "left":
        assert any(r.endswith("NULL\tNULL\tNULL") for r in got) and any(not r.endswith("NULL") for r in got)
    if jt == "mark":
        assert {r.split("\t")[-1] for r in got} == {"true", "false", "NULL"}


def test_inner_join_with_duplicate_build_keys_filter_and_project_above(pg, data):
    """lineitem x orders probing from ORDERS (build = lineitem: several build rows per key -> every pair is emitted),
    a Filter above the join that mixes both sides, and a Project with a CASE over the joined row."""
    tables, rows, S = data
    K, X, B = _ops()
    OI, LI = S.idx["orders"], S.idx["lineitem"]
    D152, I = K.DecimalType(15, 2), K.IntegerType()
    jouts = [X.col(0, OI["o_orderkey"], K.BigintType()), X.col(0, OI["o_orderdate"], K.DateType()), X.col(1, LI["l_shipdate"], K.DateType()),
             X.col(1, LI["l_extendedprice"], D152), X.col(1, LI["l_quantity"], I), X.col(0, OI["o_totalprice"], D152)]
    j = _join(X, K, B, S, X.JOIN_INNER, jouts, probe="orders", build="lineitem", pk="o_orderkey", bk="l_orderkey",
              bfilters=[X.func("<", B, S.col("lineitem", "l_linenumber"), X.const(5, I))])
    flt = X.PhysicalOperator(X.POT_Filter, Children=[j],
                             Filters=[X.func(">", B, X.col(0, 2, K.DateType()), X.col(0, 1, K.DateType())),
                                      X.func("<", B, X.func("*", K.DecimalType(18, 2), X.col(0, 3, D152), X.cast(X.const(3, I), D152)), X.col(0, 5, D152))])
    share = X.func("case", K.DecimalType(38, 6), X.const(None, K.DecimalType(38, 6)),
                   X.func(">", B, X.col(0, 4, I), X.const(20, I)), X.func("/", K.DecimalType(38, 6), X.col(0, 3, D152), X.col(0, 5, D152)))
    op = X.PhysicalOperator(X.POT_Project, Outputs=[X.col(0, 0, K.BigintType()), X.col(0, 2, K.DateType()), share], Children=[flt])
    got = _check(op, tables, rows)
    assert len(got) > 100


def test_left_join_filter_sees_the_padded_rows(pg, data):
    """Filter(c_nationkey IS-not-comparable) above a LEFT join: the NULL-padded rows reach the filter as NULLs (and fail it)."""
    tables, rows, S = data
    K, X, B = _ops()
    OI, CI = S.idx["orders"], S.idx["customer"]
    I = K.IntegerType()
    jouts = [X.col(0, OI["o_orderkey"], K.BigintType()), X.col(1, CI["c_nationkey"], I), X.col(1, CI["c_custkey"], I)]
    j = _join(X, K, B, S, X.JOIN_LEFT, jouts)
    for filt, expect_null in ((X.func("<", B, X.col(0, 1, I), X.const(12, I)), False),
                              (X.func("or", B, X.func("<", B, X.col(0, 1, I), X.const(3, I)), X.func(">", B, X.col(0, 0, K.BigintType()), X.const(5000, K.BigintType()))), True)):
        op = X.PhysicalOperator(X.POT_Filter, Children=[j], Filters=[filt], Outputs=jouts)
        got = _check(op, tables, rows)
        assert any("NULL" in r for r in got) == expect_null


def test_bare_scan_and_bare_join_roots(pg, data):
    """no Project: a filtered Scan returns every table column, a Join its output list"""
    tables, rows, S = data
    K, X, B = _ops()
    from plan_b200 import tpch as T
    sch = tables["customer"].columns
    scan = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("customer"), Outputs=[X.col(0, i, T._ltype_of(c)) for i, c in enumerate(sch)],
                              Filters=[X.func("like", B, S.col("customer", "c_mktsegment"), X.const("%U%", K.VarcharType()))])
    got = _check(scan, tables, rows)
    assert got and all("Customer#" in r for r in got)


def test_unsupported_row_shapes_are_refused(pg, data):
    """no CPU fallback: what the row pipeline cannot run is PG_EUNSUPPORTED at prepare time"""
    tables, rows, S = data
    K, X, B = _ops()
    OI, CI = S.idx["orders"], S.idx["customer"]
    V = K.VarcharType()
    # a predicate on a host-resident VARCHAR column
    scan = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("customer"), Outputs=[X.col(0, 0, K.IntegerType())],
                              Filters=[X.func("=", B, S.col("customer", "c_name"), X.const("Customer#000000001", V))])
    ex = X.gpuPipelineExec(scan, tables)
    with pytest.raises(pg.PlanGpuError) as ei:
        ex.Init()
    assert ei.value.status == pg.PG_EUNSUPPORTED
    ex.Close()
    # an expression that needs a deeper evaluation stack than the interpreter has (right-nested: every operand is pushed first)
    I = K.IntegerType()
    e = S.col("orders", "o_shippriority")
    for _ in range(10):
        e = X.func("+", I, S.col("orders", "o_custkey"), e)
    deep = X.PhysicalOperator(X.POT_Project, Outputs=[e], Children=[X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("orders"))])
    ex = X.gpuPipelineExec(deep, tables)
    with pytest.raises(pg.PlanGpuError) as ei:
        ex.Init()
    assert ei.value.status == pg.PG_EUNSUPPORTED and "evaluation stack" in str(ei.value)
    ex.Close()
    # ... while the same sum written left-nested (what the parser produces for a + b + c ...) needs two slots and runs
    e = S.col("orders", "o_shippriority")
    for _ in range(10):
        e = X.func("+", I, e, S.col("orders", "o_custkey"))
    flat = X.PhysicalOperator(X.POT_Project, Outputs=[e], Children=[X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("orders"))])
    assert len(_check(flat, tables, rows)) == len(rows["orders"])
    # build columns above a SEMI join
    j = _join(X, K, B, S, X.JOIN_SEMI, [X.col(0, OI["o_orderkey"], K.BigintType()), X.col(1, CI["c_nationkey"], K.IntegerType())])
    ex = X.gpuPipelineExec(j, tables)
    with pytest.raises(pg.PlanGpuError) as ei:
        ex.Init()
    assert ei.value.status == pg.PG_EUNSUPPORTED
    ex.Close()


# ---- expression-driven scan aggregate (scanagg_vm.cuh): the same row programs under sum / avg / min / max / count ----

def _agg_check(op, tables, rows, expect="expression programs"):
    return _check(op, tables, rows, expect_explain=expect)


def test_aggregates_over_case_and_general_predicates(pg, data):
    """TPC-H Q12 / Q14 style: sum(CASE WHEN ... THEN x ELSE 0 END), avg / min / max of a CASE, count(expr) and count(*),
    under a filter with OR, IN and <> on INTEGER and a column-to-column comparison -- none of which lowers to ranges or
    affine products.  NULLs in l_quantity / l_discount: NULL arguments are ignored, NULL predicates are not TRUE."""
    tables, rows, S = data
    K, X, B = _ops()
    lc = lambda n: S.col("lineitem", n)   # noqa: E731
    I, D152, V, H = K.IntegerType(), K.DecimalType(15, 2), K.VarcharType(), K.HugeintType()
    filt = [X.func("or", B, X.func("in", B, lc("l_quantity"), *[X.const(q, I) for q in (1, 2, 3, 5, 8, 13, 21, 34)]),
                   X.func("and", B, X.func("<>", B, lc("l_linenumber"), X.const(2, I)), X.func("<", B, lc("l_commitdate"), lc("l_receiptdate")))),
            X.func("<", B, lc("l_shipdate"), X.const(10400, K.DateType()))]
    scan = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("lineitem"), Filters=filt)
    one = X.cast(X.const(1, I), D152)
    disc_price = X.func("*", K.DecimalType(18, 4), X.cast(lc("l_extendedprice"), K.DecimalType(16, 2)), X.func("-", K.DecimalType(16, 2), one, lc("l_discount")))
    late = X.func("case", I, X.const(0, I), X.func("<", B, lc("l_commitdate"), lc("l_receiptdate")), X.const(1, I))
    promo = X.func("case", K.DecimalType(18, 4), X.cast(X.const(0, I), K.DecimalType(18, 4)), X.func("=", B, lc("l_linestatus"), X.const("F", V)), disc_price)
    some = X.func("case", D152, X.const(None, D152), X.func(">", B, lc("l_quantity"), X.const(25, I)), lc("l_extendedprice"))
    aggs = [X.func("sum", H, late), X.func("sum", K.DecimalType(38, 4), promo), X.func("sum", K.DecimalType(38, 4), disc_price),
            X.func("avg", K.DecimalType(38, 2), some), X.func("min", D152, some), X.func("max", D152, some),
            X.func("count", H, some), X.func("count", H), X.func("avg", K.DoubleType(), lc("l_quantity"))]
    outs = [X.col(0, 0, V)] + [X.col(1, i, a.DataTyp) for i, a in enumerate(aggs)]
    op = X.PhysicalOperator(X.POT_Agg, Outputs=outs, Children=[scan], Info=X.AggOpInfo(aggs, [lc("l_returnflag")]))
    got = _agg_check(op, tables, rows)
    assert len(got) == 3
    # ungrouped, and a predicate nothing passes
    op = X.PhysicalOperator(X.POT_Agg, Outputs=outs[1:], Children=[scan], Info=X.AggOpInfo(aggs, []))
    op.Outputs = [X.col(1, i, a.DataTyp) for i, a in enumerate(aggs)]
    assert len(_agg_check(op, tables, rows)) == 1
    none = X.PhysicalOperator(X.POT_Scan, Info=X.ScanOpInfo("lineitem"),
                              Filters=[X.func("or", B, X.func("<", B, lc("l_quantity"), X.const(0, I)), X.func(">", B, lc("l_linenumber"), X.const(9, I)))])
    op = X.PhysicalOperator(X.POT_Agg, Outputs=[X.col(0, 0, V), X.col(1, 0, H)], Children=[none], Info=X.AggOpInfo([X.func("count", H)], [lc("l_returnflag")]))
    assert _agg_check(op, tables, rows) == []


def test_q1_and_q6_through_the_expression_kernel(pg, monkeypatch):
    """PG_FORCE_VM=1: TPC-H Q6 (float32 BETWEEN via cast(DECIMAL AS FLOAT)) and Q1 (8 aggregates, 3-factor products) run as
    row programs and must give exactly what the specialised kernels and the C oracle give."""
    import test_gpu_scanagg as SA
    from oracle import oracle as O
    from plan_b200 import tpch as T
    sf = 0.02
    t = T.generate_device_tables(sf, want=("lineitem",))
    try:
        _, line = O.gen_orders_lineitem(sf)
        monkeypatch.setenv("PG_FORCE_VM", "1")
        for kw in ({}, {"disc_lit": 0.05, "disc_eps": 0.02}, {"qty_lt": 1}):
            chunks, stats, explain = SA._run(T.q6_plan(**kw), t)
            assert "expression programs" in explain
            ref = O.q6(line, **kw)
            assert stats.aux[0] == ref["rows_selected"]
            if ref["has_row"]:
                got = SA._dec(chunks[0].Data[0], 0)
                assert SA._dec_value(got)[0] * 10 ** (4 - got[1]) == ref["exact"]
            else:
                assert chunks == []
        chunks, stats, explain = SA._run(T.q1_plan(), t)
        assert "expression programs" in explain
        ref = O.q1(line)
        rows = sorted("\t".join(v.GetValue(r).String() for v in chunks[0].Data) for r in range(chunks[0].Card()))
        assert rows == sorted(O.q1_text(ref).strip("\n").split("\n")[1:])
    finally:
        for x in t.values():
            x.free()


def test_aggregates_with_case_over_a_join(pg, data):
    """TPC-H Q12 / Q14 shapes: Agg <- [Filter] <- Join(lineitem, orders) with CASE / IN inside the aggregates, grouped by a
    char column of the probe side (Q12: l_shipmode; here l_returnflag) or ungrouped (Q14), INNER and LEFT.  build_join_agg
    refuses these (not affine products); the pairs of the row pipeline feed vm_scanagg_kernel instead."""
    tables, rows, S = data
    K, X, B = _ops()
    OI, LI = S.idx["orders"], S.idx["lineitem"]
    I, D152, V, H = K.IntegerType(), K.DecimalType(15, 2), K.VarcharType(), K.HugeintType()
    jouts = [X.col(0, LI["l_returnflag"], V), X.col(0, LI["l_extendedprice"], D152), X.col(0, LI["l_discount"], D152),
             X.col(1, OI["o_orderstatus"], V), X.col(1, OI["o_totalprice"], D152), X.col(0, LI["l_quantity"], I), X.col(0, LI["l_shipdate"], K.DateType()),
             X.col(1, OI["o_orderdate"], K.DateType())]
    for jt in (X.JOIN_INNER, X.JOIN_LEFT):
        j = _join(X, K, B, S, jt, jouts, probe="lineitem", build="orders", pk="l_orderkey", bk="o_orderkey",
                  pfilters=[X.func("<", B, S.col("lineitem", "l_linenumber"), X.const(6, I))],
                  bfilters=[X.func(">", B, S.col("orders", "o_totalprice"), X.const(2000000, D152))])
        one = X.cast(X.const(1, I), D152)
        rev = X.func("*", K.DecimalType(18, 4), X.cast(X.col(0, 1, D152), K.DecimalType(16, 2)), X.func("-", K.DecimalType(16, 2), one, X.col(0, 2, D152)))
        high = X.func("case", I, X.const(0, I), X.func("in", B, X.col(0, 3, V), X.const("F", V), X.const("P", V)), X.const(1, I))
        promo = X.func("case", K.DecimalType(18, 4), X.cast(X.const(0, I), K.DecimalType(18, 4)), X.func(">", B, X.col(0, 6, K.DateType()), X.col(0, 7, K.DateType())), rev)
        aggs = [X.func("sum", H, high), X.func("sum", K.DecimalType(38, 4), promo), X.func("sum", K.DecimalType(38, 4), rev),
                X.func("avg", K.DecimalType(38, 2), X.col(0, 4, D152)), X.func("count", H), X.func("max", D152, X.col(0, 4, D152)),
                X.func("avg", K.DoubleType(), X.col(0, 5, I))]
        outs = [X.col(0, 0, V)] + [X.col(1, i, a.DataTyp) for i, a in enumerate(aggs)]
        op = X.PhysicalOperator(X.POT_Agg, Outputs=outs, Children=[j], Info=X.AggOpInfo(aggs, [X.col(0, 0, V)]))
        got = _check(op, tables, rows, expect_explain="JoinAgg[expression programs]")
        assert len(got) == 3
        # ungrouped, with a filter between the aggregate and the join (it sees the joined, NULL-padded rows)
        flt = X.PhysicalOperator(X.POT_Filter, Children=[j], Filters=[X.func("or", B, X.func("<", B, X.col(0, 5, I), X.const(10, I)), X.func(">", B, X.col(0, 4, D152), X.const(20000000, D152)))])
        op = X.PhysicalOperator(X.POT_Agg, Outputs=[X.col(1, i, a.DataTyp) for i, a in enumerate(aggs)], Children=[flt], Info=X.AggOpInfo(aggs, []))
        assert len(_check(op, tables, rows, expect_explain="JoinAgg[expression programs]")) == 1


def test_q4_q12_q14_and_q19_reproduce_the_reference_golden_files(pg):
    """cases/tpch/1g/plan/q4.txt (EXISTS as a MARK join whose build side carries a column-to-column filter), q12.txt, q14.txt and q19.txt
    (an OR of three AND groups over columns of both join sides, evaluated per joined row), byte for byte, from dbgen-exact SF1 columns uploaded through the C ABI:
    the reference's OWN known answers for CASE / OR / IN / LIKE inside aggregates over a join (the expression-driven join
    aggregate of rows.cu), Q14's final FLOAT projection `100.00 * a / b` done by the host parent in float32 as the reference does."""
    import os
    from oracle import oracle as O
    from plan_b200 import compute as X, tpch as T
    golden = os.path.join(os.path.dirname(__file__), "golden")
    sf = 1.0
    orders, line = O.gen_orders_lineitem(sf)
    extra = O.gen_q12_q14_columns(sf)
    x19 = O.gen_q19_columns(sf)
    npart = len(extra["p_type"])
    host = {"q12": {"lineitem": (T.Q12_LINEITEM, dict(line, l_shipmode=extra["l_shipmode"])),
                    "orders": (T.Q12_ORDERS, dict(orders, o_orderpriority=extra["o_orderpriority"]))},
            "q4": {"orders": (T.Q4_ORDERS, dict(orders, o_orderpriority=extra["o_orderpriority"])), "lineitem": (T.Q4_LINEITEM, line)},
            "q19": {"lineitem": (T.Q19_LINEITEM, dict(line, l_shipmode=extra["l_shipmode"], l_shipinstruct=x19["l_shipinstruct"])),
                    "part": (T.Q19_PART, {"p_partkey": np.arange(1, npart + 1, dtype=np.int32), "p_brand": x19["p_brand"], "p_size": x19["p_size"],
                                          "p_container": x19["p_container"]})},
            "q14": {"lineitem": (T.Q14_LINEITEM, line),
                    "part": (T.Q14_PART, {"p_partkey": np.arange(1, npart + 1, dtype=np.int32), "p_type": extra["p_type"]})}}
    for q, plan in (("q4", T.q4_plan()), ("q12", T.q12_plan()), ("q14", T.q14_plan()), ("q19", T.q19_plan())):
        tables = {}
        for name, (sch, cols) in host[q].items():
            t = X.DeviceTable.create(name, sch)
            t.append([cols[c[0]] for c in sch])
            t.seal(0)
            tables[name] = t
        try:
            ex = X.gpuPipelineExec(plan, tables)
            ex.Init()
            assert "JoinAgg[expression programs]" in ex.Explain(), ex.Explain()
            chunks = X.drain(ex)
            ex.Close()
            if q in ("q4", "q12"):
                rows = X.order_limit(chunks, [(0, False)])
                assert X.rows_text(rows, 2 if q == "q4" else 3) == open(os.path.join(golden, "ref_sf1_%s.txt" % q)).read()
            elif q == "q19":      # a general OR over both sides of the join, above the join
                assert X.rows_text(X.order_limit(chunks, []), 1) == open(os.path.join(golden, "ref_sf1_q19.txt")).read()
            else:
                a, b = chunks[0].Data[0].Data[0], chunks[0].Data[1].Data[0]
                ref = O.q14(line, extra)
                assert int(a["coef"]) * 10 ** (4 - int(a["scale"])) == ref["promo"] and int(b["coef"]) * 10 ** (4 - int(b["scale"])) == ref["total"]
                got = O.q14_promo_revenue(int(a["coef"]) * 10 ** (4 - int(a["scale"])), int(b["coef"]) * 10 ** (4 - int(b["scale"])))
                assert "#\n" + got + "\n" == open(os.path.join(golden, "ref_sf1_q14.txt")).read()
        finally:
            for t in tables.values():
                t.free()
