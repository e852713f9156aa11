// tpch_plans.hpp -- the physical plans the reference's planner emits for TPC-H Q6 / Q1 / Q3
// (typing per SURVEY.md 8c-1), built by hand because the planner stays in Go.
#pragma once
#include <math.h>

#include "../../include/plangpu_tpch.h"
#include "gpu_exec.hpp"

namespace planhost {

inline int32_t days_from_civil(int y, unsigned m, unsigned d)
{
    y -= m <= 2;
    const int era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int)doe - 719468;
}

inline LType lineitem_type(int c)
{
    switch (c) {
    case PG_L_ORDERKEY: return LType::Bigint();
    case PG_L_EXTENDEDPRICE: case PG_L_DISCOUNT: case PG_L_TAX: return LType::Decimal(15, 2);
    case PG_L_RETURNFLAG: case PG_L_LINESTATUS: return LType::Varchar();
    case PG_L_SHIPDATE: case PG_L_COMMITDATE: case PG_L_RECEIPTDATE: return LType::Date();
    default: return LType::Integer();
    }
}
// Column positions in the BOUND tables.  By default the full generated schemas (PG_L_* / PG_O_* / PG_C_*); a host that
// uploads only the referenced columns (ScanOpInfo.Columns in the reference) installs the pruned layout here.
struct ColumnLayout {
    std::vector<int> line, orders, customer;      // full index -> bound index (-1: column not uploaded); empty = identity
    int at(const std::vector<int> &m, int c) const { return m.empty() ? c : m[(size_t)c]; }
};
inline ColumnLayout &layout() { static ColumnLayout l; return l; }
inline Expr lcol(int c, int side = 0) { return col(side, layout().at(layout().line, c), lineitem_type(c)); }
inline int ocol(int c) { return layout().at(layout().orders, c); }
inline int ccol(int c) { return layout().at(layout().customer, c); }

inline Op make(POT t) { auto p = std::make_shared<PhysicalOperator>(); p->Typ = t; return p; }

inline Expr disc_price(Expr ext, Expr disc)
{
    Expr one = cast(constI(1, LType::Integer()), LType::Decimal(15, 2));
    return func("*", LType::Decimal(18, 4), {cast(std::move(ext), LType::Decimal(16, 2)),
                                             func("-", LType::Decimal(16, 2), {one, std::move(disc)})});
}

inline Op q6_plan()
{
    LType B = LType::Boolean();
    float flo = (float)0.03 - (float)0.01, fhi = (float)0.03 + (float)0.01;   // subFloat32 / addFloat32 folding
    Expr discf = cast(lcol(PG_L_DISCOUNT), LType::Float());
    Op scan = make(POT_Scan);
    scan->Table = "lineitem";
    scan->Filters = {
        func(">=", B, {lcol(PG_L_SHIPDATE), constI(days_from_civil(1994, 1, 1), LType::Date())}),
        func("<", B, {lcol(PG_L_SHIPDATE), constI(days_from_civil(1995, 1, 1), LType::Date())}),
        func("and", B, {func(">=", B, {discf, constF(flo, LType::Float())}), func("<=", B, {discf, constF(fhi, LType::Float())})}),
        func("<", B, {lcol(PG_L_QUANTITY), constI(24, LType::Integer())})};
    Op agg = make(POT_Agg);
    agg->Aggs = {func("sum", LType::Decimal(38, 4), {func("*", LType::Decimal(18, 4), {lcol(PG_L_EXTENDEDPRICE), lcol(PG_L_DISCOUNT)})})};
    agg->Outputs = {col(1, 0, LType::Decimal(38, 4))};
    agg->Children = {scan};
    return agg;
}

inline Op q1_plan()     // Order(l_returnflag, l_linestatus) <- Agg <- Scan
{
    LType B = LType::Boolean();
    Op scan = make(POT_Scan);
    scan->Table = "lineitem";
    scan->Filters = {func("<=", B, {lcol(PG_L_SHIPDATE), constI(days_from_civil(1998, 8, 11), LType::Date())})};
    Expr one_plus_tax = func("+", LType::Decimal(16, 2), {cast(constI(1, LType::Integer()), LType::Decimal(15, 2)), lcol(PG_L_TAX)});
    Expr charge = func("*", LType::Decimal(18, 8), {disc_price(lcol(PG_L_EXTENDEDPRICE), lcol(PG_L_DISCOUNT)), cast(one_plus_tax, LType::Decimal(18, 4))});
    Op agg = make(POT_Agg);
    agg->GroupBys = {lcol(PG_L_RETURNFLAG), lcol(PG_L_LINESTATUS)};
    agg->Aggs = {func("sum", LType::Hugeint(), {lcol(PG_L_QUANTITY)}),
                 func("sum", LType::Decimal(38, 2), {lcol(PG_L_EXTENDEDPRICE)}),
                 func("sum", LType::Decimal(38, 4), {disc_price(lcol(PG_L_EXTENDEDPRICE), lcol(PG_L_DISCOUNT))}),
                 func("sum", LType::Decimal(38, 8), {charge}),
                 func("avg", LType::Double(), {lcol(PG_L_QUANTITY)}),
                 func("avg", LType::Decimal(38, 2), {lcol(PG_L_EXTENDEDPRICE)}),
                 func("avg", LType::Decimal(38, 2), {lcol(PG_L_DISCOUNT)}),
                 func("count", LType::Hugeint(), {lcol(PG_L_ORDERKEY)})};
    agg->Outputs = {col(0, 0, LType::Varchar()), col(0, 1, LType::Varchar())};
    for (size_t i = 0; i < agg->Aggs.size(); i++) agg->Outputs.push_back(col(1, (int)i, agg->Aggs[i].DataTyp));
    agg->Children = {scan};
    Op order = make(POT_Order);
    order->OrderBys = {{col(0, 0, LType::Varchar()), false}, {col(0, 1, LType::Varchar()), false}};
    order->Outputs = agg->Outputs;
    order->Children = {agg};
    return order;
}

inline Op q3_plan(int64_t limit = 10)   // Limit <- Order(revenue desc, o_orderdate) <- Agg <- Join <- {lineitem, Join <- {orders, customer}}
{
    LType B = LType::Boolean();
    int32_t d = days_from_civil(1995, 3, 29);
    Op cust = make(POT_Scan);
    cust->Table = "customer";
    cust->Filters = {func("=", B, {col(0, ccol(PG_C_MKTSEGMENT), LType::Varchar()), constS("HOUSEHOLD")})};
    Op ord = make(POT_Scan);
    ord->Table = "orders";
    ord->Filters = {func("<", B, {col(0, ocol(PG_O_ORDERDATE), LType::Date()), constI(d, LType::Date())})};
    Op line = make(POT_Scan);
    line->Table = "lineitem";
    line->Filters = {func(">", B, {lcol(PG_L_SHIPDATE), constI(d, LType::Date())})};
    Op j1 = make(POT_Join);
    j1->Children = {ord, cust};
    j1->OnConds = {func("=", B, {col(0, ocol(PG_O_CUSTKEY), LType::Integer()), col(1, ccol(PG_C_CUSTKEY), LType::Integer())})};
    j1->Outputs = {col(0, ocol(PG_O_ORDERKEY), LType::Bigint()), col(0, ocol(PG_O_ORDERDATE), LType::Date()), col(0, ocol(PG_O_SHIPPRIORITY), LType::Integer())};
    Op j2 = make(POT_Join);
    j2->Children = {line, j1};
    j2->OnConds = {func("=", B, {lcol(PG_L_ORDERKEY), col(1, 0, LType::Bigint())})};
    j2->Outputs = {lcol(PG_L_ORDERKEY), lcol(PG_L_EXTENDEDPRICE), lcol(PG_L_DISCOUNT), col(1, 1, LType::Date()), col(1, 2, LType::Integer())};
    Op agg = make(POT_Agg);
    agg->GroupBys = {col(0, 0, LType::Bigint()), col(0, 3, LType::Date()), col(0, 4, LType::Integer())};
    agg->Aggs = {func("sum", LType::Decimal(38, 4), {disc_price(col(0, 1, LType::Decimal(15, 2)), col(0, 2, LType::Decimal(15, 2)))})};
    agg->Outputs = {col(0, 0, LType::Bigint()), col(1, 0, LType::Decimal(38, 4)), col(0, 1, LType::Date()), col(0, 2, LType::Integer())};
    agg->Children = {j2};
    Op order = make(POT_Order);
    order->OrderBys = {{col(0, 1, LType::Decimal(38, 4)), true}, {col(0, 2, LType::Date()), false}};
    order->Outputs = agg->Outputs;
    order->Children = {agg};
    Op lim = make(POT_Limit);
    lim->Limit = limit;
    lim->Outputs = agg->Outputs;
    lim->Children = {order};
    return lim;
}

// Limit <- Order(o_totalprice desc, o_orderdate) <- Agg(5 keys; sum(l_quantity))
//   <- SEMI Join(o_orderkey = l_orderkey) <- { Join(lineitem, Join(orders, customer)), Agg(lineitem by l_orderkey having sum > k) }
inline Op q18_plan(int64_t qty_gt = 314, int64_t limit = 100)
{
    LType B = LType::Boolean(), I = LType::Integer(), BI = LType::Bigint(), D = LType::Date(), H = LType::Hugeint(), V = LType::Varchar(),
          P = LType::Decimal(15, 2);
    Op cust = make(POT_Scan), ord = make(POT_Scan), line = make(POT_Scan), sub_scan = make(POT_Scan);
    cust->Table = "customer";
    ord->Table = "orders";
    line->Table = "lineitem";
    sub_scan->Table = "lineitem";
    Op j1 = make(POT_Join);
    j1->Children = {ord, cust};
    j1->OnConds = {func("=", B, {col(0, PG_O_CUSTKEY, I), col(1, PG_C_CUSTKEY, I)})};
    j1->Outputs = {col(0, PG_O_ORDERKEY, BI), col(0, PG_O_ORDERDATE, D), col(0, PG_O_TOTALPRICE, P), col(1, PG_C_NAME, V), col(1, PG_C_CUSTKEY, I)};
    Op j2 = make(POT_Join);
    j2->Children = {line, j1};
    j2->OnConds = {func("=", B, {lcol(PG_L_ORDERKEY), col(1, 0, BI)})};
    j2->Outputs = {lcol(PG_L_QUANTITY), col(1, 0, BI), col(1, 1, D), col(1, 2, P), col(1, 3, V), col(1, 4, I)};
    Op sub = make(POT_Agg);
    sub->GroupBys = {lcol(PG_L_ORDERKEY)};
    sub->Aggs = {func("sum", H, {lcol(PG_L_QUANTITY)})};
    sub->Filters = {func(">", B, {col(1, 0, H), cast(constI(qty_gt, I), H)})};
    sub->Outputs = {col(0, 0, BI)};
    sub->Children = {sub_scan};
    Op semi = make(POT_Join);
    semi->JoinTyp = PG_JOIN_SEMI;
    semi->Children = {j2, sub};
    semi->OnConds = {func("=", B, {col(0, 1, BI), col(1, 0, BI)})};
    semi->Outputs = {col(0, 0, I), col(0, 1, BI), col(0, 2, D), col(0, 3, P), col(0, 4, V), col(0, 5, I)};
    Op agg = make(POT_Agg);
    agg->GroupBys = {col(0, 4, V), col(0, 5, I), col(0, 1, BI), col(0, 2, D), col(0, 3, P)};
    agg->Aggs = {func("sum", H, {col(0, 0, I)})};
    agg->Outputs = {col(0, 0, V), col(0, 1, I), col(0, 2, BI), col(0, 3, D), col(0, 4, P), col(1, 0, H)};
    agg->Children = {semi};
    Op order = make(POT_Order);
    order->OrderBys = {{col(0, 4, P), true}, {col(0, 3, D), false}};
    order->Outputs = agg->Outputs;
    order->Children = {agg};
    Op lim = make(POT_Limit);
    lim->Limit = limit;
    lim->Outputs = agg->Outputs;
    lim->Children = {order};
    return lim;
}

// Order(nation, o_year desc) <- Agg(n_name, extract(year from o_orderdate); sum(ext*(1-disc) - ps_supplycost*l_quantity))
//   <- five INNER joins stacked on the lineitem scan: part (p_name like '%pink%'), supplier, partsupp (2-column key), orders, nation
inline Op q9_plan(const std::string &word = "pink")
{
    LType B = LType::Boolean(), I = LType::Integer(), BI = LType::Bigint(), D = LType::Date(), V = LType::Varchar(), P = LType::Decimal(15, 2);
    auto scan = [](const char *name) { Op s = make(POT_Scan); s->Table = name; return s; };
    Op line = scan("lineitem"), part = scan("part"), supp = scan("supplier"), ps = scan("partsupp"), ord = scan("orders"), nat = scan("nation");
    part->Filters = {func("like", B, {col(0, PG_P_NAME, V), constS("%" + word + "%")})};
    std::vector<LType> types = {BI, I, I, I, P, P};
    auto up = [&]() { std::vector<Expr> o; for (size_t i = 0; i < types.size(); i++) o.push_back(col(0, (int)i, types[i])); return o; };
    Op j1 = make(POT_Join);
    j1->Children = {line, part};
    j1->OnConds = {func("=", B, {lcol(PG_L_PARTKEY), col(1, PG_P_PARTKEY, I)})};
    j1->Outputs = {lcol(PG_L_ORDERKEY), lcol(PG_L_PARTKEY), lcol(PG_L_SUPPKEY), lcol(PG_L_QUANTITY), lcol(PG_L_EXTENDEDPRICE), lcol(PG_L_DISCOUNT)};
    Op j2 = make(POT_Join);
    j2->Children = {j1, supp};
    j2->OnConds = {func("=", B, {col(0, 2, I), col(1, PG_S_SUPPKEY, I)})};
    j2->Outputs = up();
    j2->Outputs.push_back(col(1, PG_S_NATIONKEY, I));
    types.push_back(I);
    Op j3 = make(POT_Join);
    j3->Children = {j2, ps};
    j3->OnConds = {func("=", B, {col(0, 2, I), col(1, PG_PS_SUPPKEY, I)}), func("=", B, {col(0, 1, I), col(1, PG_PS_PARTKEY, I)})};
    j3->Outputs = up();
    j3->Outputs.push_back(col(1, PG_PS_SUPPLYCOST, P));
    types.push_back(P);
    Op j4 = make(POT_Join);
    j4->Children = {j3, ord};
    j4->OnConds = {func("=", B, {col(0, 0, BI), col(1, PG_O_ORDERKEY, BI)})};
    j4->Outputs = up();
    j4->Outputs.push_back(col(1, PG_O_ORDERDATE, D));
    types.push_back(D);
    Op j5 = make(POT_Join);
    j5->Children = {j4, nat};
    j5->OnConds = {func("=", B, {col(0, 6, I), col(1, PG_N_NATIONKEY, I)})};
    j5->Outputs = up();
    j5->Outputs.push_back(col(1, PG_N_NAME, V));
    Expr amount = func("-", LType::Decimal(18, 4), {disc_price(col(0, 4, P), col(0, 5, P)),
                                                    func("*", LType::Decimal(18, 4), {col(0, 7, P), cast(col(0, 3, I), P)})});
    Op agg = make(POT_Agg);
    agg->GroupBys = {col(0, 9, V), func("extract", I, {constS("year"), col(0, 8, D)})};
    agg->Aggs = {func("sum", LType::Decimal(38, 4), {amount})};
    agg->Outputs = {col(0, 0, V), col(0, 1, I), col(1, 0, LType::Decimal(38, 4))};
    agg->Children = {j5};
    Op order = make(POT_Order);
    order->OrderBys = {{col(0, 0, V), false}, {col(0, 1, I), true}};
    order->Outputs = agg->Outputs;
    order->Children = {agg};
    return order;
}

}  // namespace planhost
