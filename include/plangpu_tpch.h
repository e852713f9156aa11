/*
 * plangpu_tpch.h -- in-box synthetic TPC-H data for benchmarks and tests (NOT part of
 * the reference-facing operator ABI).  There is no network on the GPU box, so lineitem /
 * orders / customer are generated directly in HBM by a dbgen-equivalent generator: the
 * same per-column Park-Miller streams, 7 draws per order per lineitem stream, sparse
 * order keys -- bit-identical to the official dbgen for the generated columns (the
 * reference's golden SF1 results, /root/reference/cases/tpch/1g/plan/q{1,3,6}.txt, are
 * reproduced on it).  Any order range [order_lo, order_hi) is generated independently,
 * which is how row-range shards are produced in place on each GPU.
 */
#ifndef PLANGPU_TPCH_H
#define PLANGPU_TPCH_H
#include "plangpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* column order of the generated tables */
enum { PG_L_ORDERKEY = 0, PG_L_PARTKEY, PG_L_SUPPKEY, PG_L_LINENUMBER, PG_L_QUANTITY, PG_L_EXTENDEDPRICE,
       PG_L_DISCOUNT, PG_L_TAX, PG_L_RETURNFLAG, PG_L_LINESTATUS, PG_L_SHIPDATE, PG_L_COMMITDATE,
       PG_L_RECEIPTDATE, PG_L_NCOLS };
enum { PG_O_ORDERKEY = 0, PG_O_CUSTKEY, PG_O_ORDERDATE, PG_O_SHIPPRIORITY, PG_O_TOTALPRICE, PG_O_ORDERSTATUS,
       PG_O_NCOLS };
enum { PG_C_CUSTKEY = 0, PG_C_MKTSEGMENT, PG_C_NATIONKEY, PG_C_NAME, PG_C_NCOLS };

enum { PG_P_PARTKEY = 0, PG_P_NAME, PG_P_NCOLS };
enum { PG_S_SUPPKEY = 0, PG_S_NATIONKEY, PG_S_NCOLS };
enum { PG_PS_PARTKEY = 0, PG_PS_SUPPKEY, PG_PS_SUPPLYCOST, PG_PS_NCOLS };
enum { PG_N_NATIONKEY = 0, PG_N_NAME, PG_N_NCOLS };

int64_t pg_tpch_num_orders(double sf);
int64_t pg_tpch_num_customers(double sf);

/* orders [order_lo, order_hi) (0-based order index) and their lineitems, generated on the
 * device into two new SEALED tables.  global row offsets are recorded in the tables so
 * shards merge in row order.  Either out pointer may be NULL to skip that table
 * (lineitem generation still needs the line counts). */
int pg_tpch_orders_lineitem(double sf, int64_t order_lo, int64_t order_hi, pg_table **orders, pg_table **lineitem);
int pg_tpch_customer(double sf, int64_t cust_lo, int64_t cust_hi, pg_table **customer);

/* the dimension tables TPC-H Q9 joins (whole tables, SEALED): part(p_partkey, p_name VARCHAR),
 * supplier(s_suppkey, s_nationkey), partsupp(ps_partkey, ps_suppkey, ps_supplycost DECIMAL(15,2)),
 * nation(n_nationkey, n_name as a 25-entry dictionary column) */
int pg_tpch_part(double sf, pg_table **part);
int pg_tpch_supplier(double sf, pg_table **supplier);
int pg_tpch_partsupp(double sf, pg_table **partsupp);
/* rows [row_lo, row_hi) of partsupp (4 rows per part, in ps_partkey order; row_hi < 0 = to the end): a row-range
 * shard of the build side of Q9's lineitem x partsupp join, which is then keyed differently from the fact table's
 * shards -- the case the all-to-all row exchange exists for (BASELINE config 5) */
int pg_tpch_partsupp_range(double sf, int64_t row_lo, int64_t row_hi, pg_table **partsupp);
int64_t pg_tpch_num_parts(double sf);
int pg_tpch_nation(pg_table **nation);

#ifdef __cplusplus
}
#endif
#endif
