// append_bench.cc -- the ingest path as the reference's executor would drive it, measured end to end.
//
// The Go shim receives what scanExecutor produces: chunks of util.DefaultVectorSize = 2048 rows
// (/root/reference/pkg/util/util.go:124, pkg/compute/executor_scan.go:158-223) whose vectors live on the
// pageable Go heap (util.GAlloc = make([]byte), pkg/util/mem.go:29-39), flattens each needed vector into a
// native column buffer and calls pg_table_append once per chunk.  This harness does exactly that from C++:
// PAGEABLE host columns, one pg_table_append[_cols] call per <chunk_rows> rows, then seal, compile, execute
// Q6 / Q1 / Q3(top 10) and fetch the results -- all inside the timed region.
//
//   planhost_append <sf> <chunk_rows> <native|narrow> <steps>
//     native: buffers in the column's native encoding (int64 DECIMAL / BIGINT, int32 INTEGER / DATE)
//     narrow: pg_colbuf buffers at the narrowest width of each column's value range (what a shim that
//             tracks min/max while flattening hands over)
// One JSON line on stdout; the Q6 / Q1 / Q3 result text (golden-file format) on stderr for the caller to check.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <sstream>

#include "tpch_plans.hpp"

using namespace planhost;

namespace {

struct HostCol {
    std::string name;
    int type = 0, width = 0, scale = 0;
    std::vector<const char *> dict;
    int native = 0;                    // bytes per value in the native encoding
    std::vector<uint8_t> wide;         // pageable: native encoding
    std::vector<uint8_t> narrow;       // pageable: narrow encoding
    int nwidth = 0;
    int64_t nbase = 0;
};
struct HostTable {
    std::string name;
    std::vector<HostCol> cols;
    int64_t nrows = 0;
};

int native_size(int t) { return (t == PG_T_INT32 || t == PG_T_DATE32) ? 4 : (t == PG_T_CHAR1 || t == PG_T_DICT8) ? 1 : 8; }

template <typename T> void minmax(const std::vector<uint8_t> &b, int64_t n, int64_t *lo, int64_t *hi)
{
    const T *v = (const T *)b.data();
    int64_t a = INT64_MAX, z = INT64_MIN;
    for (int64_t i = 0; i < n; i++) { a = v[i] < a ? v[i] : a; z = v[i] > z ? v[i] : z; }
    *lo = a; *hi = z;
}
template <typename S, typename D> void narrow_copy(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst, int64_t n, int64_t base)
{
    dst.resize((size_t)n * sizeof(D));
    const S *s = (const S *)src.data();
    D *d = (D *)dst.data();
    for (int64_t i = 0; i < n; i++) d[i] = (D)((int64_t)s[i] - base);
}

// what a shim that knows the value range of a column hands over: 1, 2 or 4 bytes per value + a base
void make_narrow(HostCol &c, int64_t n)
{
    c.nwidth = c.native;
    c.nbase = 0;
    if (c.type == PG_T_CHAR1 || c.type == PG_T_DICT8 || n == 0) return;
    int64_t lo, hi;
    if (c.native == 8) minmax<int64_t>(c.wide, n, &lo, &hi); else minmax<int32_t>(c.wide, n, &lo, &hi);
    const uint64_t span = (uint64_t)hi - (uint64_t)lo;
    int w = c.native;
    if (span <= 0xff) w = 1; else if (span <= 0xffff) w = 2; else if (span <= 0x7fffffff && c.native == 8) w = 4;
    if (w >= c.native) return;
    c.nwidth = w;
    c.nbase = lo;
    if (c.native == 8) {
        if (w == 1) narrow_copy<int64_t, uint8_t>(c.wide, c.narrow, n, lo);
        else if (w == 2) narrow_copy<int64_t, uint16_t>(c.wide, c.narrow, n, lo);
        else narrow_copy<int64_t, int32_t>(c.wide, c.narrow, n, lo);
    } else {
        if (w == 1) narrow_copy<int32_t, uint8_t>(c.wide, c.narrow, n, lo);
        else narrow_copy<int32_t, uint16_t>(c.wide, c.narrow, n, lo);
    }
}

HostTable pull(pg_table *t, const std::string &name, const std::vector<std::pair<int, HostCol>> &cols)
{
    HostTable h;
    h.name = name;
    check(pg_table_rows(t, &h.nrows));
    for (auto &kv : cols) {
        HostCol c = kv.second;
        c.native = native_size(c.type);
        c.wide.resize((size_t)h.nrows * (size_t)c.native);
        check(pg_table_read_column(t, kv.first, 0, h.nrows, c.wide.data()));
        make_narrow(c, h.nrows);
        h.cols.push_back(std::move(c));
    }
    return h;
}

pg_table *ingest(const HostTable &h, int64_t chunk_rows, bool narrow, int64_t *calls, int64_t *bytes)
{
    std::vector<pg_coldesc> cd(h.cols.size());
    for (size_t i = 0; i < h.cols.size(); i++) {
        const HostCol &c = h.cols[i];
        cd[i] = pg_coldesc{c.name.c_str(), c.type, c.width, c.scale, (int32_t)c.dict.size(), c.dict.empty() ? nullptr : c.dict.data()};
    }
    pg_table *t = nullptr;
    check(pg_table_create(h.name.c_str(), (int)cd.size(), cd.data(), &t));
    std::vector<pg_colbuf> bufs(h.cols.size());
    std::vector<const void *> ptrs(h.cols.size());
    for (int64_t r = 0; r < h.nrows; r += chunk_rows) {
        const int64_t n = std::min(chunk_rows, h.nrows - r);
        for (size_t i = 0; i < h.cols.size(); i++) {
            const HostCol &c = h.cols[i];
            if (narrow) {
                const bool nn = c.nwidth < c.native;
                bufs[i].data = nn ? c.narrow.data() + (size_t)r * (size_t)c.nwidth : c.wide.data() + (size_t)r * (size_t)c.native;
                bufs[i].width = nn ? c.nwidth : 0;
                bufs[i].reserved = 0;
                bufs[i].base = nn ? c.nbase : 0;
                bufs[i].valid = nullptr;
                *bytes += n * (nn ? c.nwidth : c.native);
            } else {
                ptrs[i] = c.wide.data() + (size_t)r * (size_t)c.native;
                *bytes += n * c.native;
            }
        }
        if (narrow) check(pg_table_append_cols(t, n, bufs.data()));
        else check(pg_table_append(t, n, ptrs.data(), nullptr));
        ++*calls;
    }
    check(pg_table_seal(t, 0));
    return t;
}

std::string run_query(Op plan, const std::map<std::string, pg_table *> &tables, size_t ncols, int64_t *d2h)
{
    GpuPipelineExec ex(plan, tables);
    ex.Init();
    std::string text = "#" + std::string(ncols - 1, '\t') + "\n";
    for (;;) {
        Chunk out;
        if (ex.Execute(nullptr, &out) == Done) break;
        for (auto &v : out.Data) *d2h += (int64_t)v.Data.size();
        for (int64_t r = 0; r < out.Count; r++) {
            for (size_t c = 0; c < out.Data.size(); c++) { if (c) text += '\t'; text += out.Data[c].ValueString(r); }
            text += '\n';
        }
    }
    ex.Close();
    return text;
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s <sf> <chunk_rows> <native|narrow> <steps>\n", argv[0]); return 2; }
    const double sf = atof(argv[1]);
    const int64_t chunk_rows = atoll(argv[2]);
    const bool narrow = strcmp(argv[3], "narrow") == 0;
    const int steps = atoi(argv[4]);
    static const char *segs[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"};
    try {
        check(pg_init(0));
        HostTable line, orders, customer;
        {
            pg_table *o = nullptr, *l = nullptr, *c = nullptr;
            check(pg_tpch_orders_lineitem(sf, 0, pg_tpch_num_orders(sf), &o, &l));
            check(pg_tpch_customer(sf, 0, pg_tpch_num_customers(sf), &c));
            auto hc = [](const char *n, int t, int w = 0, int s = 0) { HostCol c; c.name = n; c.type = t; c.width = w; c.scale = s; return c; };
            line = pull(l, "lineitem", {{PG_L_ORDERKEY, hc("l_orderkey", PG_T_INT64)}, {PG_L_QUANTITY, hc("l_quantity", PG_T_INT32)},
                                        {PG_L_EXTENDEDPRICE, hc("l_extendedprice", PG_T_DECIMAL64, 15, 2)}, {PG_L_DISCOUNT, hc("l_discount", PG_T_DECIMAL64, 15, 2)},
                                        {PG_L_TAX, hc("l_tax", PG_T_DECIMAL64, 15, 2)}, {PG_L_RETURNFLAG, hc("l_returnflag", PG_T_CHAR1)},
                                        {PG_L_LINESTATUS, hc("l_linestatus", PG_T_CHAR1)}, {PG_L_SHIPDATE, hc("l_shipdate", PG_T_DATE32)}});
            orders = pull(o, "orders", {{PG_O_ORDERKEY, hc("o_orderkey", PG_T_INT64)}, {PG_O_CUSTKEY, hc("o_custkey", PG_T_INT32)},
                                        {PG_O_ORDERDATE, hc("o_orderdate", PG_T_DATE32)}, {PG_O_SHIPPRIORITY, hc("o_shippriority", PG_T_INT32)}});
            HostCol seg = hc("c_mktsegment", PG_T_DICT8);
            seg.dict.assign(segs, segs + 5);
            customer = pull(c, "customer", {{PG_C_CUSTKEY, hc("c_custkey", PG_T_INT32)}, {PG_C_MKTSEGMENT, seg}});
            pg_table_free(o);
            pg_table_free(l);
            pg_table_free(c);
        }
        // the uploaded tables hold only the referenced columns: install the pruned layout for the plan builders
        ColumnLayout &lay = layout();
        lay.line.assign(PG_L_NCOLS, -1);
        lay.line[PG_L_ORDERKEY] = 0; lay.line[PG_L_QUANTITY] = 1; lay.line[PG_L_EXTENDEDPRICE] = 2; lay.line[PG_L_DISCOUNT] = 3;
        lay.line[PG_L_TAX] = 4; lay.line[PG_L_RETURNFLAG] = 5; lay.line[PG_L_LINESTATUS] = 6; lay.line[PG_L_SHIPDATE] = 7;
        lay.orders.assign(PG_O_NCOLS, -1);
        lay.orders[PG_O_ORDERKEY] = 0; lay.orders[PG_O_CUSTKEY] = 1; lay.orders[PG_O_ORDERDATE] = 2; lay.orders[PG_O_SHIPPRIORITY] = 3;
        lay.customer.assign(PG_C_NCOLS, -1);
        lay.customer[PG_C_CUSTKEY] = 0; lay.customer[PG_C_MKTSEGMENT] = 1;

        std::string q6, q1, q3;
        int64_t calls = 0, bytes = 0, d2h = 0;
        double total_ms = 0;
        for (int s = -1; s < steps; s++) {          // step -1 is the warm-up (allocator, pinned staging)
            calls = bytes = d2h = 0;
            auto t0 = std::chrono::steady_clock::now();
            std::map<std::string, pg_table *> tabs;
            tabs["lineitem"] = ingest(line, chunk_rows, narrow, &calls, &bytes);
            tabs["orders"] = ingest(orders, chunk_rows, narrow, &calls, &bytes);
            tabs["customer"] = ingest(customer, chunk_rows, narrow, &calls, &bytes);
            q6 = run_query(q6_plan(), tabs, 1, &d2h);
            q1 = run_query(q1_plan(), tabs, 10, &d2h);
            q3 = run_query(q3_plan(10), tabs, 4, &d2h);
            for (auto &kv : tabs) pg_table_free(kv.second);
            auto t1 = std::chrono::steady_clock::now();
            if (s >= 0) total_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
        }
        fprintf(stderr, "== q6\n%s== q1\n%s== q3\n%s", q6.c_str(), q1.c_str(), q3.c_str());
        printf("{\"ms_per_step\": %.3f, \"steps\": %d, \"lineitem_rows\": %lld, \"chunk_rows\": %lld, \"append_calls_per_step\": %lld, "
               "\"h2d_bytes_per_step\": %lld, \"d2h_bytes_per_step\": %lld, \"buffers\": \"%s\", \"host_memory\": \"pageable (malloc)\"}\n",
               total_ms / steps, steps, (long long)line.nrows, (long long)chunk_rows, (long long)calls, (long long)bytes, (long long)d2h,
               narrow ? "narrow (pg_table_append_cols)" : "native (pg_table_append)");
    } catch (const PlanError &e) {
        fprintf(stderr, "plangpu error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
