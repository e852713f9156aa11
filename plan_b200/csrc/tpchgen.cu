// tpchgen.cu -- dbgen-equivalent TPC-H generator running on the GPU (bench/test data).
//
// Independent CUDA restatement of dbgen's published algorithm (see plangpu_tpch.h);
// cross-checked row for row against the CPU generator of the oracle in tests.
// One thread per order: the j-th lineitem of order i takes draw 7*i+j+1 of each L_*
// stream, so the stream state is seed * 16807^(7i) mod (2^31-1) -- two modular
// exponentiations per thread, no sequential dependency between orders.
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/plangpu_tpch.h"
#include "common.cuh"
#include "pipeline.hpp"

namespace pg {

__host__ __device__ inline i64 tg_mulmod(i64 a, i64 b) { return (i64)(((u64)a * (u64)b) % 2147483647ULL); }
__host__ __device__ inline i64 tg_pow(i64 base, i64 n)
{
    i64 r = 1;
    while (n > 0) {
        if (n & 1) r = tg_mulmod(r, base);
        base = tg_mulmod(base, base);
        n >>= 1;
    }
    return r;
}
__device__ inline i64 tg_draw(i64 &s, i64 lo, i64 hi)
{
    s = tg_mulmod(s, 16807);
    double r = (double)(hi - lo + 1);
    return lo + (i64)(((double)s / 2147483647.0) * r);
}

constexpr i64 SD_L_QTY = 209208115, SD_L_DCNT = 554590007, SD_L_TAX = 721958466, SD_L_PKEY = 1808217256,
              SD_L_SKEY = 2095021727, SD_L_SDTE = 1769349045, SD_L_CDTE = 904914315, SD_L_RDTE = 373135028,
              SD_L_RFLG = 717419739, SD_O_ODATE = 1066728069, SD_O_CKEY = 851767375, SD_O_LCNT = 1434868289,
              SD_C_MSEG = 1140279430, SD_C_NTRG = 1489529863;
constexpr int EPOCH_1992 = 8035, ODATE_SPAN = 2406, CURRENT_OFF = 1263;

__global__ void tg_count_kernel(i64 o_lo, i64 n, int *__restrict__ counts)
{
    i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    i64 s = tg_mulmod(SD_O_LCNT, tg_pow(16807, o_lo + k));
    counts[k] = (int)tg_draw(s, 1, 7);
}

struct TgOut {
    i64 *o_orderkey; int *o_custkey; int *o_orderdate; int *o_shippriority; i64 *o_totalprice; uint8_t *o_orderstatus;
    i64 *l_orderkey; int *l_partkey; int *l_suppkey; int *l_linenumber; int *l_quantity; i64 *l_extendedprice;
    i64 *l_discount; i64 *l_tax; uint8_t *l_returnflag; uint8_t *l_linestatus; int *l_shipdate; int *l_commitdate;
    int *l_receiptdate;
};

__global__ void tg_fill_kernel(i64 o_lo, i64 n, const i64 *__restrict__ line_off, i64 ncust, i64 npart, i64 nsupp,
                               bool want_orders, bool want_lines, TgOut o)
{
    i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    i64 i = o_lo + k;
    i64 p1 = tg_pow(16807, i), p7 = tg_pow(16807, 7 * i);
    i64 s_ckey = tg_mulmod(SD_O_CKEY, p1), s_odate = tg_mulmod(SD_O_ODATE, p1), s_lcnt = tg_mulmod(SD_O_LCNT, p1);
    i64 okey = (((i + 1) >> 3) << 5) | ((i + 1) & 7);
    i64 ck = tg_draw(s_ckey, 1, ncust);
    i64 delta = 1;
    while (ck % 3 == 0) { ck += delta; if (ck > ncust) ck = ncust; delta = -delta; }
    i64 od = tg_draw(s_odate, 0, ODATE_SPAN - 1);
    int lines = (int)tg_draw(s_lcnt, 1, 7);
    i64 s_qty = tg_mulmod(SD_L_QTY, p7), s_dc = tg_mulmod(SD_L_DCNT, p7), s_tax = tg_mulmod(SD_L_TAX, p7),
        s_pk = tg_mulmod(SD_L_PKEY, p7), s_sk = tg_mulmod(SD_L_SKEY, p7), s_sd = tg_mulmod(SD_L_SDTE, p7),
        s_cd = tg_mulmod(SD_L_CDTE, p7), s_rd = tg_mulmod(SD_L_RDTE, p7), s_rf = tg_mulmod(SD_L_RFLG, p7);
    i64 row = line_off[k];
    i64 total = 0;
    int shipped = 0;
    for (int j = 0; j < lines; j++, row++) {
        i64 qty = tg_draw(s_qty, 1, 50);
        i64 dc = tg_draw(s_dc, 0, 10);
        i64 tax = tg_draw(s_tax, 0, 8);
        i64 pk = tg_draw(s_pk, 1, npart);
        i64 sn = tg_draw(s_sk, 0, 3);
        i64 sd = od + tg_draw(s_sd, 1, 121);
        i64 cd = od + tg_draw(s_cd, 30, 90);
        i64 rd = sd + tg_draw(s_rd, 1, 30);
        i64 price = 90000 + (pk / 10) % 20001 + (pk % 1000) * 100;
        i64 ep = price * qty;
        i64 sk = (pk + sn * (nsupp / 4 + (pk - 1) / nsupp)) % nsupp + 1;
        uint8_t rf = 'N';
        if (rd <= CURRENT_OFF) rf = tg_draw(s_rf, 1, 2) == 1 ? 'R' : 'A';
        uint8_t ls = 'O';
        if (sd <= CURRENT_OFF) { ls = 'F'; shipped++; }
        total += ((ep * (100 - dc)) / 100) * (100 + tax) / 100;
        if (want_lines) {
            o.l_orderkey[row] = okey; o.l_partkey[row] = (int)pk; o.l_suppkey[row] = (int)sk;
            o.l_linenumber[row] = j + 1; o.l_quantity[row] = (int)qty; o.l_extendedprice[row] = ep;
            o.l_discount[row] = dc; o.l_tax[row] = tax; o.l_returnflag[row] = rf; o.l_linestatus[row] = ls;
            o.l_shipdate[row] = (int)(EPOCH_1992 + sd); o.l_commitdate[row] = (int)(EPOCH_1992 + cd);
            o.l_receiptdate[row] = (int)(EPOCH_1992 + rd);
        }
    }
    if (want_orders) {
        o.o_orderkey[k] = okey; o.o_custkey[k] = (int)ck; o.o_orderdate[k] = (int)(EPOCH_1992 + od);
        o.o_shippriority[k] = 0; o.o_totalprice[k] = total;
        o.o_orderstatus[k] = shipped == 0 ? 'O' : (shipped == lines ? 'F' : 'P');
    }
}

__global__ void tg_customer_kernel(i64 c_lo, i64 n, int *__restrict__ custkey, uint8_t *__restrict__ seg, int *__restrict__ nat)
{
    i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    i64 p1 = tg_pow(16807, c_lo + k);
    i64 s_seg = tg_mulmod(SD_C_MSEG, p1), s_nat = tg_mulmod(SD_C_NTRG, p1);
    custkey[k] = (int)(c_lo + k + 1);
    seg[k] = (uint8_t)(tg_draw(s_seg, 1, 5) - 1);
    nat[k] = (int)tg_draw(s_nat, 0, 24);
}

static i64 num_orders(double sf) { return (i64)(1500000.0 * sf + 0.5); }
static i64 num_customers(double sf) { return (i64)(150000.0 * sf + 0.5); }
static i64 num_parts(double sf) { return (i64)(200000.0 * sf + 0.5); }
static i64 num_supp(double sf) { i64 n = (i64)(10000.0 * sf + 0.5); return n < 4 ? 4 : n; }

// ---- part / supplier / partsupp (TPC-H Q9's dimension tables) --------------------------------
constexpr i64 SD_P_NAME = 709314158, SD_PS_SCST = 1051288424, SD_S_NTRG = 110356601;
constexpr int TG_NAME_SLOT = 56;
__device__ const char tg_colors[92][12] = {
    "almond", "antique", "aquamarine", "azure", "beige", "bisque", "black", "blanched", "blue", "blush", "brown", "burlywood",
    "burnished", "chartreuse", "chiffon", "chocolate", "coral", "cornflower", "cornsilk", "cream", "cyan", "dark", "deep", "dim",
    "dodger", "drab", "firebrick", "floral", "forest", "frosted", "gainsboro", "ghost", "goldenrod", "green", "grey", "honeydew",
    "hot", "indian", "ivory", "khaki", "lace", "lavender", "lawn", "lemon", "light", "lime", "linen", "magenta", "maroon", "medium",
    "metallic", "midnight", "mint", "misty", "moccasin", "navajo", "navy", "olive", "orange", "orchid", "pale", "papaya", "peach",
    "peru", "pink", "plum", "powder", "puff", "purple", "red", "rose", "rosy", "royal", "saddle", "salmon", "sandy", "seashell",
    "sienna", "sky", "slate", "smoke", "snow", "spring", "steel", "tan", "thistle", "tomato", "turquoise", "violet", "wheat", "white",
    "yellow"};

// P_NAME: the identity permutation of the 92 colours shuffled with 92 draws of the part's own slice of
// stream P_NAME_SD (swap a[k], a[UnifInt(k, 91)]), first five joined with blanks
__global__ void tg_part_kernel(i64 n, int *__restrict__ p_partkey, char *__restrict__ names, uint8_t *__restrict__ name_len)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i64 s = tg_mulmod(SD_P_NAME, tg_pow(16807, 92 * i));
    uint8_t perm[92];
    for (int k = 0; k < 92; k++) perm[k] = (uint8_t)k;
    for (int k = 0; k < 92; k++) {
        int src = (int)tg_draw(s, k, 91);
        uint8_t t = perm[src]; perm[src] = perm[k]; perm[k] = t;
    }
    char *out = names + i * TG_NAME_SLOT;
    int at = 0;
    for (int w = 0; w < 5; w++) {
        const char *c = tg_colors[perm[w]];
        for (int j = 0; c[j]; j++) out[at++] = c[j];
        if (w < 4) out[at++] = ' ';
    }
    name_len[i] = (uint8_t)at;
    p_partkey[i] = (int)(i + 1);
}

__global__ void tg_supplier_kernel(i64 n, int *__restrict__ s_suppkey, int *__restrict__ s_nationkey)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i64 s = tg_mulmod(SD_S_NTRG, tg_pow(16807, i));
    s_suppkey[i] = (int)(i + 1);
    s_nationkey[i] = (int)tg_draw(s, 0, 24);
}

__global__ void tg_partsupp_kernel(i64 row_lo, i64 nrows, i64 nsupp, int *__restrict__ ps_partkey, int *__restrict__ ps_suppkey, i64 *__restrict__ ps_supplycost)
{
    i64 o = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nrows) return;
    const i64 r = row_lo + o;                     // global row: any row range is generated independently
    const i64 pk = r / 4 + 1, j = r % 4;
    i64 s = tg_mulmod(SD_PS_SCST, tg_pow(16807, r));
    ps_partkey[o] = (int)pk;
    ps_suppkey[o] = (int)((pk + j * (nsupp / 4 + (pk - 1) / nsupp)) % nsupp + 1);
    ps_supplycost[o] = tg_draw(s, 100, 100000);
}

}  // namespace pg

using namespace pg;

extern "C" {

int64_t pg_tpch_num_orders(double sf) { return num_orders(sf); }
int64_t pg_tpch_num_customers(double sf) { return num_customers(sf); }
int64_t pg_tpch_num_parts(double sf) { return num_parts(sf); }

int pg_tpch_orders_lineitem(double sf, int64_t order_lo, int64_t order_hi, pg_table **orders, pg_table **lineitem)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_tpch_orders_lineitem: call pg_init first");
    if (order_lo < 0 || order_hi < order_lo || order_hi > num_orders(sf)) PG_FAIL(PG_EINVAL, "pg_tpch_orders_lineitem: bad order range");
    PG_CUDA(cudaSetDevice(c.device));
    cudaStream_t st = c.stream;
    i64 n = order_hi - order_lo;
    DevBuf d_counts, d_off, d_tmp, d_before;
    PG_TRY(d_counts.alloc(sizeof(int) * (size_t)(n + 1)));
    PG_TRY(d_off.alloc(sizeof(i64) * (size_t)(n + 1)));
    PG_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * (size_t)(n + 1), st));
    int threads = 256;
    if (n > 0) tg_count_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(order_lo, n, d_counts.as<int>());
    PG_CUDA(cudaGetLastError());
    // exclusive prefix sum of the line counts -> first lineitem row of each order
    size_t tmp_bytes = 0;
    PG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts.as<int>(), d_off.as<i64>(), (int)(n + 1), st));
    PG_TRY(d_tmp.alloc(tmp_bytes));
    PG_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_counts.as<int>(), d_off.as<i64>(), (int)(n + 1), st));
    i64 nlines = 0;
    PG_CUDA(cudaMemcpyAsync(&nlines, d_off.as<i64>() + n, sizeof(i64), cudaMemcpyDeviceToHost, st));
    // global lineitem row offset of this shard = lineitems of orders [0, order_lo)
    i64 before = 0;
    if (order_lo > 0) {
        // count in bounded pieces to keep the scratch small
        i64 done = 0;
        const i64 piece = (i64)1 << 24;
        DevBuf d_c2, d_sum, d_t2;
        PG_TRY(d_c2.alloc(sizeof(int) * (size_t)piece));
        PG_TRY(d_sum.alloc(sizeof(i64)));
        size_t t2 = 0;
        PG_CUDA(cub::DeviceReduce::Sum(nullptr, t2, d_c2.as<int>(), d_sum.as<int>(), (int)piece, st));
        PG_TRY(d_t2.alloc(t2));
        while (done < order_lo) {
            i64 m = std::min(piece, order_lo - done);
            tg_count_kernel<<<(unsigned)((m + threads - 1) / threads), threads, 0, st>>>(done, m, d_c2.as<int>());
            int *d_isum = d_sum.as<int>();   // <= 7 * 2^24 fits in int32
            PG_CUDA(cub::DeviceReduce::Sum(d_t2.p, t2, d_c2.as<int>(), d_isum, (int)m, st));
            int part = 0;
            PG_CUDA(cudaMemcpyAsync(&part, d_isum, sizeof(int), cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            before += part;
            done += m;
        }
    }
    PG_CUDA(cudaStreamSynchronize(st));

    pg_table *to = nullptr, *tl = nullptr;
    TgOut o{};
    if (orders) {
        pg_coldesc cd[PG_O_NCOLS] = {
            {"o_orderkey", PG_T_INT64, 0, 0, 0, nullptr}, {"o_custkey", PG_T_INT32, 0, 0, 0, nullptr},
            {"o_orderdate", PG_T_DATE32, 0, 0, 0, nullptr}, {"o_shippriority", PG_T_INT32, 0, 0, 0, nullptr},
            {"o_totalprice", PG_T_DECIMAL64, 15, 2, 0, nullptr}, {"o_orderstatus", PG_T_CHAR1, 0, 0, 0, nullptr}};
        PG_TRY(pg_table_create("orders", PG_O_NCOLS, cd, &to));
        PG_TRY(pg_table_reserve(to, n));
        o.o_orderkey = (i64 *)to->cols[PG_O_ORDERKEY].d_data; o.o_custkey = (int *)to->cols[PG_O_CUSTKEY].d_data;
        o.o_orderdate = (int *)to->cols[PG_O_ORDERDATE].d_data; o.o_shippriority = (int *)to->cols[PG_O_SHIPPRIORITY].d_data;
        o.o_totalprice = (i64 *)to->cols[PG_O_TOTALPRICE].d_data; o.o_orderstatus = (uint8_t *)to->cols[PG_O_ORDERSTATUS].d_data;
    }
    if (lineitem) {
        pg_coldesc cd[PG_L_NCOLS] = {
            {"l_orderkey", PG_T_INT64, 0, 0, 0, nullptr}, {"l_partkey", PG_T_INT32, 0, 0, 0, nullptr},
            {"l_suppkey", PG_T_INT32, 0, 0, 0, nullptr}, {"l_linenumber", PG_T_INT32, 0, 0, 0, nullptr},
            {"l_quantity", PG_T_INT32, 0, 0, 0, nullptr}, {"l_extendedprice", PG_T_DECIMAL64, 15, 2, 0, nullptr},
            {"l_discount", PG_T_DECIMAL64, 15, 2, 0, nullptr}, {"l_tax", PG_T_DECIMAL64, 15, 2, 0, nullptr},
            {"l_returnflag", PG_T_CHAR1, 0, 0, 0, nullptr}, {"l_linestatus", PG_T_CHAR1, 0, 0, 0, nullptr},
            {"l_shipdate", PG_T_DATE32, 0, 0, 0, nullptr}, {"l_commitdate", PG_T_DATE32, 0, 0, 0, nullptr},
            {"l_receiptdate", PG_T_DATE32, 0, 0, 0, nullptr}};
        PG_TRY(pg_table_create("lineitem", PG_L_NCOLS, cd, &tl));
        PG_TRY(pg_table_reserve(tl, nlines));
        o.l_orderkey = (i64 *)tl->cols[PG_L_ORDERKEY].d_data; o.l_partkey = (int *)tl->cols[PG_L_PARTKEY].d_data;
        o.l_suppkey = (int *)tl->cols[PG_L_SUPPKEY].d_data; o.l_linenumber = (int *)tl->cols[PG_L_LINENUMBER].d_data;
        o.l_quantity = (int *)tl->cols[PG_L_QUANTITY].d_data; o.l_extendedprice = (i64 *)tl->cols[PG_L_EXTENDEDPRICE].d_data;
        o.l_discount = (i64 *)tl->cols[PG_L_DISCOUNT].d_data; o.l_tax = (i64 *)tl->cols[PG_L_TAX].d_data;
        o.l_returnflag = (uint8_t *)tl->cols[PG_L_RETURNFLAG].d_data; o.l_linestatus = (uint8_t *)tl->cols[PG_L_LINESTATUS].d_data;
        o.l_shipdate = (int *)tl->cols[PG_L_SHIPDATE].d_data; o.l_commitdate = (int *)tl->cols[PG_L_COMMITDATE].d_data;
        o.l_receiptdate = (int *)tl->cols[PG_L_RECEIPTDATE].d_data;
    }
    if (n > 0 && (orders || lineitem)) {
        tg_fill_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(
            order_lo, n, d_off.as<i64>(), num_customers(sf), num_parts(sf), num_supp(sf), orders != nullptr, lineitem != nullptr, o);
        PG_CUDA(cudaGetLastError());
    }
    PG_CUDA(cudaStreamSynchronize(st));
    if (orders) {
        PG_TRY(pg_table_set_rows(to, n));
        PG_TRY(pg_table_seal(to, order_lo));
        *orders = to;
    }
    if (lineitem) {
        PG_TRY(pg_table_set_rows(tl, nlines));
        PG_TRY(pg_table_seal(tl, before));
        *lineitem = tl;
    }
    return PG_OK;
}

int pg_tpch_customer(double sf, int64_t cust_lo, int64_t cust_hi, pg_table **customer)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_tpch_customer: call pg_init first");
    if (!customer || cust_lo < 0 || cust_hi < cust_lo || cust_hi > num_customers(sf)) PG_FAIL(PG_EINVAL, "pg_tpch_customer: bad arguments");
    PG_CUDA(cudaSetDevice(c.device));
    static const char *segs[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"};
    pg_coldesc cd[PG_C_NCOLS] = {{"c_custkey", PG_T_INT32, 0, 0, 0, nullptr},
                                 {"c_mktsegment", PG_T_DICT8, 0, 0, 5, segs},
                                 {"c_nationkey", PG_T_INT32, 0, 0, 0, nullptr},
                                 {"c_name", PG_T_VARCHAR, 25, 0, 0, nullptr}};
    pg_table *t = nullptr;
    PG_TRY(pg_table_create("customer", PG_C_NCOLS, cd, &t));
    i64 n = cust_hi - cust_lo;
    PG_TRY(pg_table_reserve(t, n));
    if (n > 0) {
        tg_customer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(
            cust_lo, n, (int *)t->cols[PG_C_CUSTKEY].d_data, (uint8_t *)t->cols[PG_C_MKTSEGMENT].d_data,
            (int *)t->cols[PG_C_NATIONKEY].d_data);
        PG_CUDA(cudaGetLastError());
    }
    PG_CUDA(cudaStreamSynchronize(c.stream));
    PG_TRY(pg_table_set_rows(t, n));
    {   // C_NAME = "Customer#" + key zero-padded to 9 digits (TPC-H 4.2.3); VARCHAR columns live on the host
        Column &nm = t->cols[PG_C_NAME];
        nm.h_off.resize((size_t)n + 1);
        nm.h_bytes.resize((size_t)n * 18);
        char *b = &nm.h_bytes[0];
        for (i64 i = 0; i < n; i++) {
            char *s = b + (size_t)i * 18;
            memcpy(s, "Customer#", 9);
            i64 key = cust_lo + i + 1;
            for (int d = 17; d >= 9; d--) { s[d] = (char)('0' + key % 10); key /= 10; }
            nm.h_off[(size_t)i] = i * 18;
        }
        nm.h_off[(size_t)n] = n * 18;
    }
    PG_TRY(pg_table_seal(t, cust_lo));
    *customer = t;
    return PG_OK;
}

int pg_tpch_part(double sf, pg_table **part)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_tpch_part: call pg_init first");
    if (!part) PG_FAIL(PG_EINVAL, "pg_tpch_part: bad arguments");
    PG_CUDA(cudaSetDevice(c.device));
    pg_coldesc cd[PG_P_NCOLS] = {{"p_partkey", PG_T_INT32, 0, 0, 0, nullptr}, {"p_name", PG_T_VARCHAR, 55, 0, 0, nullptr}};
    pg_table *t = nullptr;
    PG_TRY(pg_table_create("part", PG_P_NCOLS, cd, &t));
    const i64 n = num_parts(sf);
    PG_TRY(pg_table_reserve(t, n));
    DevBuf d_names, d_len;
    PG_TRY(d_names.alloc((size_t)n * TG_NAME_SLOT));
    PG_TRY(d_len.alloc((size_t)n));
    tg_part_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c.stream>>>(n, (int *)t->cols[PG_P_PARTKEY].d_data, d_names.as<char>(), d_len.as<uint8_t>());
    PG_CUDA(cudaGetLastError());
    std::vector<char> h_names((size_t)n * TG_NAME_SLOT);
    std::vector<uint8_t> h_len((size_t)n);
    PG_CUDA(cudaMemcpyAsync(h_names.data(), d_names.p, h_names.size(), cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaMemcpyAsync(h_len.data(), d_len.p, h_len.size(), cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaStreamSynchronize(c.stream));
    PG_TRY(pg_table_set_rows(t, n));
    Column &nm = t->cols[PG_P_NAME];          // VARCHAR columns live on the host (plus a device copy made at seal)
    nm.h_off.resize((size_t)n + 1);
    size_t total = 0;
    for (i64 i = 0; i < n; i++) { nm.h_off[(size_t)i] = (int64_t)total; total += h_len[(size_t)i]; }
    nm.h_off[(size_t)n] = (int64_t)total;
    nm.h_bytes.resize(total);
    for (i64 i = 0; i < n; i++) memcpy(&nm.h_bytes[(size_t)nm.h_off[(size_t)i]], h_names.data() + (size_t)i * TG_NAME_SLOT, h_len[(size_t)i]);
    PG_TRY(pg_table_seal(t, 0));
    *part = t;
    return PG_OK;
}

int pg_tpch_supplier(double sf, pg_table **supplier)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_tpch_supplier: call pg_init first");
    if (!supplier) PG_FAIL(PG_EINVAL, "pg_tpch_supplier: bad arguments");
    PG_CUDA(cudaSetDevice(c.device));
    pg_coldesc cd[PG_S_NCOLS] = {{"s_suppkey", PG_T_INT32, 0, 0, 0, nullptr}, {"s_nationkey", PG_T_INT32, 0, 0, 0, nullptr}};
    pg_table *t = nullptr;
    PG_TRY(pg_table_create("supplier", PG_S_NCOLS, cd, &t));
    const i64 n = num_supp(sf);
    PG_TRY(pg_table_reserve(t, n));
    tg_supplier_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(n, (int *)t->cols[PG_S_SUPPKEY].d_data, (int *)t->cols[PG_S_NATIONKEY].d_data);
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaStreamSynchronize(c.stream));
    PG_TRY(pg_table_set_rows(t, n));
    PG_TRY(pg_table_seal(t, 0));
    *supplier = t;
    return PG_OK;
}

int pg_tpch_partsupp_range(double sf, int64_t row_lo, int64_t row_hi, pg_table **partsupp)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_tpch_partsupp: call pg_init first");
    if (!partsupp) PG_FAIL(PG_EINVAL, "pg_tpch_partsupp: bad arguments");
    const i64 total = 4 * num_parts(sf);
    if (row_hi < 0) row_hi = total;
    if (row_lo < 0 || row_hi < row_lo || row_hi > total) PG_FAIL(PG_EINVAL, "pg_tpch_partsupp_range: bad row range");
    PG_CUDA(cudaSetDevice(c.device));
    pg_coldesc cd[PG_PS_NCOLS] = {{"ps_partkey", PG_T_INT32, 0, 0, 0, nullptr}, {"ps_suppkey", PG_T_INT32, 0, 0, 0, nullptr},
                                  {"ps_supplycost", PG_T_DECIMAL64, 15, 2, 0, nullptr}};
    pg_table *t = nullptr;
    PG_TRY(pg_table_create("partsupp", PG_PS_NCOLS, cd, &t));
    const i64 n = row_hi - row_lo;
    PG_TRY(pg_table_reserve(t, n));
    if (n > 0)
        tg_partsupp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(row_lo, n, num_supp(sf), (int *)t->cols[PG_PS_PARTKEY].d_data,
                                                                            (int *)t->cols[PG_PS_SUPPKEY].d_data, (i64 *)t->cols[PG_PS_SUPPLYCOST].d_data);
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaStreamSynchronize(c.stream));
    PG_TRY(pg_table_set_rows(t, n));
    PG_TRY(pg_table_seal(t, row_lo));
    *partsupp = t;
    return PG_OK;
}

int pg_tpch_partsupp(double sf, pg_table **partsupp) { return pg_tpch_partsupp_range(sf, 0, -1, partsupp); }

int pg_tpch_nation(pg_table **nation)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_tpch_nation: call pg_init first");
    if (!nation) PG_FAIL(PG_EINVAL, "pg_tpch_nation: bad arguments");
    static const char *names[25] = {"ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY", "INDIA", "INDONESIA",
                                    "IRAN", "IRAQ", "JAPAN", "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA", "ROMANIA",
                                    "SAUDI ARABIA", "VIETNAM", "RUSSIA", "UNITED KINGDOM", "UNITED STATES"};
    pg_coldesc cd[PG_N_NCOLS] = {{"n_nationkey", PG_T_INT32, 0, 0, 0, nullptr}, {"n_name", PG_T_DICT8, 0, 0, 25, names}};
    pg_table *t = nullptr;
    PG_TRY(pg_table_create("nation", PG_N_NCOLS, cd, &t));
    int32_t keys[25];
    uint8_t codes[25];
    for (int i = 0; i < 25; i++) { keys[i] = i; codes[i] = (uint8_t)i; }
    const void *cols[2] = {keys, codes};
    PG_TRY(pg_table_append(t, 25, cols, nullptr));
    PG_TRY(pg_table_seal(t, 0));
    *nation = t;
    return PG_OK;
}

}  // extern "C"
