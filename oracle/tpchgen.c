/*
 * oracle/tpchgen.c -- TEST INFRASTRUCTURE (CPU). Not part of the product path.
 *
 * dbgen-equivalent TPC-H generator for the columns the reference's TPC-H cases
 * read.  The reference loads official dbgen SF1 data
 * (/root/reference/Makefile:47,57-67) and its golden results
 * (cases/tpch/1g/plan/q*.txt) are computed on it, so this generator
 * restates dbgen's published algorithm bit for bit for those columns:
 *   - Park-Miller LCG  seed' = seed*16807 mod (2^31-1), one stream per column,
 *     UnifInt = lo + (int64)((double)seed/2147483647.0 * (hi-lo+1));
 *   - every lineitem stream is consumed exactly 7 times per order (row_stop
 *     advances the unused draws), so stream position = 7*order_index + line,
 *     which makes any order range independently generable (jump-ahead by
 *     modular exponentiation);
 *   - sparse order keys (8 of every 32), customer mortality (custkey%3 != 0),
 *     retail price from partkey, R/A return flag drawn only when
 *     receiptdate <= 1995-06-17.
 * Pinned by tests/test_oracle_golden.py against the first rows of the official SF1
 * tables and, end to end, by ALL 22 of the reference's golden result files
 * (cases/tpch/1g/plan/q1.txt ... q22.txt): the sections further down add, query by
 * query, the other columns those files read -- part / supplier / partsupp / nation,
 * ship modes, priorities, part types / brands / containers, balances, generated
 * addresses and phone numbers, the Customer-Complaints suppliers and dbgen's 300 MiB
 * comment text pool.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TG_M 2147483647LL
#define TG_A 16807LL

/* dbgen stream seeds (Seed[] table of dbgen's rnd.h, by column) */
enum {
    SD_L_QTY = 209208115, SD_L_DCNT = 554590007, SD_L_TAX = 721958466,
    SD_L_PKEY = 1808217256, SD_L_SKEY = 2095021727, SD_L_SDTE = 1769349045,
    SD_L_CDTE = 904914315, SD_L_RDTE = 373135028, SD_L_RFLG = 717419739,
    SD_O_ODATE = 1066728069, SD_O_CKEY = 851767375, SD_O_LCNT = 1434868289,
    SD_C_MSEG = 1140279430, SD_C_NTRG = 1489529863
};

#define EPOCH_1992_01_01 8035   /* days from 1970-01-01 to 1992-01-01 */
#define ODATE_SPAN 2406         /* 1992-01-01 .. 1998-08-02 inclusive  */
#define CURRENT_OFF 1263        /* 1995-06-17 as offset from 1992-01-01 */

static inline int64_t tg_mulmod(int64_t a, int64_t b) { return (int64_t)(((__int128)a * b) % TG_M); }

static int64_t tg_powmod(int64_t n)
{
    int64_t r = 1, b = TG_A;
    while (n > 0) {
        if (n & 1) r = tg_mulmod(r, b);
        b = tg_mulmod(b, b);
        n >>= 1;
    }
    return r;
}

/* state of a stream after n draws from its initial seed */
static inline int64_t tg_jump(int64_t seed, int64_t n) { return tg_mulmod(seed, tg_powmod(n)); }

static inline int64_t tg_draw(int64_t *s, int64_t lo, int64_t hi)
{
    *s = (*s * TG_A) % TG_M;
    double r = (double)(hi - lo + 1);
    return lo + (int64_t)(((double)*s / 2147483647.0) * r);
}

int64_t tg_num_orders(double sf) { return (int64_t)(1500000.0 * sf + 0.5); }
int64_t tg_num_customers(double sf) { return (int64_t)(150000.0 * sf + 0.5); }
static int64_t tg_num_parts(double sf) { return (int64_t)(200000.0 * sf + 0.5); }
static int64_t tg_num_supp(double sf) { int64_t n = (int64_t)(10000.0 * sf + 0.5); return n < 4 ? 4 : n; }

/* number of lineitem rows belonging to orders [o_lo, o_hi) (0-based order index) */
int64_t tg_count_lineitems(double sf, int64_t o_lo, int64_t o_hi)
{
    (void)sf;
    int64_t s = tg_jump(SD_O_LCNT, o_lo), n = 0;
    for (int64_t i = o_lo; i < o_hi; i++) n += tg_draw(&s, 1, 7);
    return n;
}

static inline int64_t tg_orderkey(int64_t idx1)   /* idx1 is 1-based */
{
    return ((idx1 >> 3) << 5) | (idx1 & 7);
}

/*
 * Generate orders [o_lo,o_hi) and their lineitems.  Any output pointer may be
 * NULL.  Lineitem arrays must hold tg_count_lineitems(sf,o_lo,o_hi) rows.
 * DECIMAL(15,2) columns are int64 unscaled (cents); dates are int32 days since
 * 1970-01-01; flags are raw bytes.
 * Returns the number of lineitem rows written.
 */
int64_t tg_gen_orders_lineitem(double sf, int64_t o_lo, int64_t o_hi,
    /* orders */
    int64_t *o_orderkey, int32_t *o_custkey, int32_t *o_orderdate, int32_t *o_shippriority,
    int64_t *o_totalprice, uint8_t *o_orderstatus,
    /* lineitem */
    int64_t *l_orderkey, int32_t *l_partkey, int32_t *l_suppkey, int32_t *l_linenumber,
    int32_t *l_quantity, int64_t *l_extendedprice, int64_t *l_discount, int64_t *l_tax,
    uint8_t *l_returnflag, uint8_t *l_linestatus,
    int32_t *l_shipdate, int32_t *l_commitdate, int32_t *l_receiptdate)
{
    const int64_t ncust = tg_num_customers(sf), npart = tg_num_parts(sf), nsupp = tg_num_supp(sf);
    int64_t s_ckey = tg_jump(SD_O_CKEY, o_lo), s_odate = tg_jump(SD_O_ODATE, o_lo),
            s_lcnt = tg_jump(SD_O_LCNT, o_lo);
    int64_t row = 0;
    for (int64_t i = o_lo; i < o_hi; i++) {
        int64_t okey = tg_orderkey(i + 1);
        int64_t ck = tg_draw(&s_ckey, 1, ncust);
        int64_t delta = 1;
        while (ck % 3 == 0) { ck += delta; if (ck > ncust) ck = ncust; delta = -delta; }
        int64_t od = tg_draw(&s_odate, 0, ODATE_SPAN - 1);
        int64_t lines = tg_draw(&s_lcnt, 1, 7);
        int64_t base = 7 * i;
        int64_t s_qty = tg_jump(SD_L_QTY, base), s_dc = tg_jump(SD_L_DCNT, base),
                s_tax = tg_jump(SD_L_TAX, base), s_pk = tg_jump(SD_L_PKEY, base),
                s_sk = tg_jump(SD_L_SKEY, base), s_sd = tg_jump(SD_L_SDTE, base),
                s_cd = tg_jump(SD_L_CDTE, base), s_rd = tg_jump(SD_L_RDTE, base),
                s_rf = tg_jump(SD_L_RFLG, base);
        int64_t total = 0; int shipped = 0;
        for (int64_t j = 0; j < lines; j++, row++) {
            int64_t qty = tg_draw(&s_qty, 1, 50);
            int64_t dc = tg_draw(&s_dc, 0, 10);
            int64_t tax = tg_draw(&s_tax, 0, 8);
            int64_t pk = tg_draw(&s_pk, 1, npart);
            int64_t sn = tg_draw(&s_sk, 0, 3);
            int64_t sd = od + tg_draw(&s_sd, 1, 121);
            int64_t cd = od + tg_draw(&s_cd, 30, 90);
            int64_t rd = sd + tg_draw(&s_rd, 1, 30);
            int64_t price = 90000 + (pk / 10) % 20001 + (pk % 1000) * 100;
            int64_t ep = price * qty;
            int64_t sk = (pk + sn * (nsupp / 4 + (pk - 1) / nsupp)) % nsupp + 1;
            uint8_t rf = 'N';
            if (rd <= CURRENT_OFF) rf = (tg_draw(&s_rf, 1, 2) == 1) ? 'R' : 'A';
            uint8_t ls = 'O';
            if (sd <= CURRENT_OFF) { ls = 'F'; shipped++; }
            total += ((ep * (100 - dc)) / 100) * (100 + tax) / 100;
            if (l_orderkey) l_orderkey[row] = okey;
            if (l_partkey) l_partkey[row] = (int32_t)pk;
            if (l_suppkey) l_suppkey[row] = (int32_t)sk;
            if (l_linenumber) l_linenumber[row] = (int32_t)(j + 1);
            if (l_quantity) l_quantity[row] = (int32_t)qty;
            if (l_extendedprice) l_extendedprice[row] = ep;
            if (l_discount) l_discount[row] = dc;
            if (l_tax) l_tax[row] = tax;
            if (l_returnflag) l_returnflag[row] = rf;
            if (l_linestatus) l_linestatus[row] = ls;
            if (l_shipdate) l_shipdate[row] = (int32_t)(EPOCH_1992_01_01 + sd);
            if (l_commitdate) l_commitdate[row] = (int32_t)(EPOCH_1992_01_01 + cd);
            if (l_receiptdate) l_receiptdate[row] = (int32_t)(EPOCH_1992_01_01 + rd);
        }
        int64_t k = i - o_lo;
        if (o_orderkey) o_orderkey[k] = okey;
        if (o_custkey) o_custkey[k] = (int32_t)ck;
        if (o_orderdate) o_orderdate[k] = (int32_t)(EPOCH_1992_01_01 + od);
        if (o_shippriority) o_shippriority[k] = 0;
        if (o_totalprice) o_totalprice[k] = total;
        if (o_orderstatus) o_orderstatus[k] = shipped == 0 ? 'O' : (shipped == lines ? 'F' : 'P');
    }
    return row;
}

/* customer segment dictionary, dists.dss "msegmnt" order */
static const char *TG_SEGMENTS[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"};
const char *tg_segment_name(int code) { return (code >= 0 && code < 5) ? TG_SEGMENTS[code] : ""; }

/* customers [c_lo,c_hi) (0-based); c_mktsegment is the dictionary code 0..4 */
void tg_gen_customer(double sf, int64_t c_lo, int64_t c_hi,
                     int32_t *c_custkey, uint8_t *c_mktsegment, int32_t *c_nationkey)
{
    (void)sf;
    int64_t s_seg = tg_jump(SD_C_MSEG, c_lo), s_nat = tg_jump(SD_C_NTRG, c_lo);
    for (int64_t i = c_lo; i < c_hi; i++) {
        int64_t seg = tg_draw(&s_seg, 1, 5);
        int64_t nat = tg_draw(&s_nat, 0, 24);
        if (c_custkey) c_custkey[i - c_lo] = (int32_t)(i + 1);
        if (c_mktsegment) c_mktsegment[i - c_lo] = (uint8_t)(seg - 1);
        if (c_nationkey) c_nationkey[i - c_lo] = (int32_t)nat;
    }
}

/* ------------------------------------------------------------------------------------------------
 * part / supplier / partsupp / nation -- the columns TPC-H Q9 touches (cases/tpch/query/q9.sql).
 * dbgen: P_NAME = 5 words of the 92-word "colors" distribution: the identity permutation of the 92
 * indices is shuffled with 92 draws of stream P_NAME_SD (for i in 0..91: swap a[i], a[UnifInt(i, 91)])
 * and the first five members are joined with blanks; PS_SUPPLYCOST = UnifInt(100, 100000) cents from
 * PS_SCST_SD, 4 partsupp rows per part with the same supplier bridge as l_suppkey; S_NATIONKEY =
 * UnifInt(0, 24) from S_NTRG_SD; the 25 nations are fixed.
 */
enum { SD_P_NAME = 709314158, SD_PS_SCST = 1051288424, SD_S_NTRG = 110356601 };

static const char *TG_COLORS[92] = {
    "almond", "antique", "aquamarine", "azure", "beige", "bisque", "black", "blanched", "blue", "blush", "brown", "burlywood",
    "burnished", "chartreuse", "chiffon", "chocolate", "coral", "cornflower", "cornsilk", "cream", "cyan", "dark", "deep", "dim",
    "dodger", "drab", "firebrick", "floral", "forest", "frosted", "gainsboro", "ghost", "goldenrod", "green", "grey", "honeydew",
    "hot", "indian", "ivory", "khaki", "lace", "lavender", "lawn", "lemon", "light", "lime", "linen", "magenta", "maroon", "medium",
    "metallic", "midnight", "mint", "misty", "moccasin", "navajo", "navy", "olive", "orange", "orchid", "pale", "papaya", "peach",
    "peru", "pink", "plum", "powder", "puff", "purple", "red", "rose", "rosy", "royal", "saddle", "salmon", "sandy", "seashell",
    "sienna", "sky", "slate", "smoke", "snow", "spring", "steel", "tan", "thistle", "tomato", "turquoise", "violet", "wheat", "white",
    "yellow"};

static const char *TG_NATIONS[25] = {
    "ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY", "INDIA", "INDONESIA", "IRAN", "IRAQ", "JAPAN",
    "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA", "ROMANIA", "SAUDI ARABIA", "VIETNAM", "RUSSIA", "UNITED KINGDOM",
    "UNITED STATES"};

const char *tg_nation_name(int key) { return (key >= 0 && key < 25) ? TG_NATIONS[key] : ""; }
int64_t tg_num_parts_pub(double sf) { return tg_num_parts(sf); }
int64_t tg_num_supp_pub(double sf) { return tg_num_supp(sf); }

/* parts [p_lo, p_hi) (0-based).  name_buf receives the names back to back, name_off[i]..name_off[i+1];
 * name_off must hold (p_hi - p_lo + 1) entries, name_buf 56 bytes per part.  contains_word (optional):
 * 1 when the name contains `word` as a substring (the `p_name like '%word%'` predicate). */
void tg_gen_part(double sf, int64_t p_lo, int64_t p_hi, int32_t *p_partkey, char *name_buf, int64_t *name_off,
                 const char *word, uint8_t *contains_word)
{
    (void)sf;
    int64_t s = tg_jump(SD_P_NAME, 92 * p_lo), at = 0;
    for (int64_t i = p_lo; i < p_hi; i++) {
        int perm[92];
        for (int k = 0; k < 92; k++) perm[k] = k;
        for (int k = 0; k < 92; k++) {
            int64_t src = tg_draw(&s, k, 91);
            int t = perm[src]; perm[src] = perm[k]; perm[k] = t;
        }
        char nm[64];
        size_t n = 0;
        for (int w = 0; w < 5; w++) {
            const char *c = TG_COLORS[perm[w]];
            size_t l = strlen(c);
            memcpy(nm + n, c, l);
            n += l;
            if (w < 4) nm[n++] = ' ';
        }
        nm[n] = 0;
        if (p_partkey) p_partkey[i - p_lo] = (int32_t)(i + 1);
        if (name_off) name_off[i - p_lo] = at;
        if (name_buf) memcpy(name_buf + at, nm, n);
        if (contains_word) contains_word[i - p_lo] = (word && strstr(nm, word)) ? 1 : 0;
        at += (int64_t)n;
    }
    if (name_off) name_off[p_hi - p_lo] = at;
}

void tg_gen_supplier(double sf, int64_t s_lo, int64_t s_hi, int32_t *s_suppkey, int32_t *s_nationkey)
{
    (void)sf;
    int64_t s = tg_jump(SD_S_NTRG, s_lo);
    for (int64_t i = s_lo; i < s_hi; i++) {
        int64_t nat = tg_draw(&s, 0, 24);
        if (s_suppkey) s_suppkey[i - s_lo] = (int32_t)(i + 1);
        if (s_nationkey) s_nationkey[i - s_lo] = (int32_t)nat;
    }
}

/* the 4 partsupp rows of every part in [p_lo, p_hi): output arrays hold 4*(p_hi-p_lo) rows */
void tg_gen_partsupp(double sf, int64_t p_lo, int64_t p_hi, int32_t *ps_partkey, int32_t *ps_suppkey, int64_t *ps_supplycost)
{
    const int64_t nsupp = tg_num_supp(sf);
    int64_t s = tg_jump(SD_PS_SCST, 4 * p_lo), row = 0;
    for (int64_t i = p_lo; i < p_hi; i++) {
        const int64_t pk = i + 1;
        for (int64_t j = 0; j < 4; j++, row++) {
            int64_t cost = tg_draw(&s, 100, 100000);
            if (ps_partkey) ps_partkey[row] = (int32_t)pk;
            if (ps_suppkey) ps_suppkey[row] = (int32_t)((pk + j * (nsupp / 4 + (pk - 1) / nsupp)) % nsupp + 1);
            if (ps_supplycost) ps_supplycost[row] = cost;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Columns of TPC-H Q12 / Q14 beyond the ones above (cases/tpch/query/q12.sql, q14.sql): the raw draws of
 *   l_shipmode      pick_str(smode,  L_SMODE_SD): UnifInt(1, 7)   per line, 7 draws per order like every lineitem stream
 *   o_orderpriority pick_str(o_oprio, O_PRIO_SD): UnifInt(1, 5)   per order
 *   p_type          pick_str(p_types, P_TYPE_SD): UnifInt(1, 150) per part
 * as 0-based indices into dbgen's distributions (dists.dss); the callers hold the member lists.  Pinned by the
 * reference's golden cases/tpch/1g/plan/q12.txt and q14.txt (tests/test_oracle_golden.py).
 */
enum { SD_L_SMODE = 675466456, SD_O_PRIO = 591449447, SD_P_TYPE = 1841581359 };

void tg_gen_q12_q14_draws(double sf, int64_t o_lo, int64_t o_hi, uint8_t *o_prio /* [orders] */, uint8_t *l_smode /* [lines] */,
                          int64_t p_lo, int64_t p_hi, uint8_t *p_type /* [parts] */)
{
    (void)sf;
    if (o_prio || l_smode) {
        int64_t s_lcnt = tg_jump(SD_O_LCNT, o_lo), s_prio = tg_jump(SD_O_PRIO, o_lo);
        int64_t row = 0;
        for (int64_t i = o_lo; i < o_hi; i++) {
            const int64_t lines = tg_draw(&s_lcnt, 1, 7);
            const int64_t pr = tg_draw(&s_prio, 1, 5);
            if (o_prio) o_prio[i - o_lo] = (uint8_t)(pr - 1);
            int64_t s_sm = tg_jump(SD_L_SMODE, 7 * i);
            for (int64_t j = 0; j < lines; j++, row++) {
                const int64_t m = tg_draw(&s_sm, 1, 7);
                if (l_smode) l_smode[row] = (uint8_t)(m - 1);
            }
        }
    }
    if (p_type) {
        int64_t s = tg_jump(SD_P_TYPE, p_lo);
        for (int64_t i = p_lo; i < p_hi; i++) p_type[i - p_lo] = (uint8_t)(tg_draw(&s, 1, 150) - 1);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Columns of TPC-H Q19 (cases/tpch/query/q19.sql): raw draws as 0-based indices / values
 *   p_brand      "Brand#MN": M = UnifInt(1,5) from P_MFG_SD, N = UnifInt(1,5) from P_BRND_SD  -> code (M-1)*5 + (N-1)
 *   p_size       UnifInt(1,50) from P_SIZE_SD
 *   p_container  pick_str(p_cntr, P_CNTR_SD): UnifInt(1,40) -> index into {SM,LG,MED,JUMBO,WRAP} x {CASE,BOX,BAG,JAR,PKG,PACK,CAN,DRUM}
 *   l_shipinstruct pick_str(instruct, L_SHIP_SD): UnifInt(1,4) per line (7 draws per order); index 0 = DELIVER IN PERSON
 * Seeds and member orders pinned by the reference's golden cases/tpch/1g/plan/q19.txt (tests/test_oracle_golden.py).
 */
enum { SD_P_MFG = 1, SD_P_BRND = 46831694, SD_P_SIZE = 1193163244, SD_P_CNTR = 727633698, SD_L_SHIP = 1371272478 };

void tg_gen_q19_draws(double sf, int64_t o_lo, int64_t o_hi, uint8_t *l_shipinstruct /* [lines] */,
                      int64_t p_lo, int64_t p_hi, uint8_t *p_brand, int32_t *p_size, uint8_t *p_container /* [parts] */)
{
    (void)sf;
    if (l_shipinstruct) {
        int64_t s_lcnt = tg_jump(SD_O_LCNT, o_lo);
        int64_t row = 0;
        for (int64_t i = o_lo; i < o_hi; i++) {
            const int64_t lines = tg_draw(&s_lcnt, 1, 7);
            int64_t s = tg_jump(SD_L_SHIP, 7 * i);
            for (int64_t j = 0; j < lines; j++, row++) l_shipinstruct[row] = (uint8_t)(tg_draw(&s, 1, 4) - 1);
        }
    }
    if (p_brand || p_size || p_container) {
        int64_t s_m = tg_jump(SD_P_MFG, p_lo), s_b = tg_jump(SD_P_BRND, p_lo), s_s = tg_jump(SD_P_SIZE, p_lo), s_c = tg_jump(SD_P_CNTR, p_lo);
        for (int64_t i = p_lo; i < p_hi; i++) {
            const int64_t m = tg_draw(&s_m, 1, 5), b = tg_draw(&s_b, 1, 5), sz = tg_draw(&s_s, 1, 50), c = tg_draw(&s_c, 1, 40);
            if (p_brand) p_brand[i - p_lo] = (uint8_t)((m - 1) * 5 + (b - 1));
            if (p_size) p_size[i - p_lo] = (int32_t)sz;
            if (p_container) p_container[i - p_lo] = (uint8_t)(c - 1);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Columns of TPC-H Q11 / Q22 (cases/tpch/query/q11.sql, q22.sql):
 *   c_acctbal    UnifInt(-99999, 999999) cents from C_ABAL_SD, one draw per customer
 *   ps_availqty  UnifInt(1, 9999) from PS_QTY_SD, one draw per partsupp row (4 per part)
 * (substring(c_phone, 1, 2) needs no stream: dbgen's phone number starts with the two digits of 10 + c_nationkey.)
 * Pinned by the reference's golden cases/tpch/1g/plan/q11.txt and q22.txt (tests/test_oracle_golden.py).
 */
enum { SD_C_ABAL = 298370230, SD_PS_QTY = 1671059989 };

void tg_gen_q11_q22_draws(double sf, int64_t c_lo, int64_t c_hi, int64_t *c_acctbal /* [customers] */,
                          int64_t p_lo, int64_t p_hi, int32_t *ps_availqty /* [4 * parts] */)
{
    (void)sf;
    if (c_acctbal) {
        int64_t s = tg_jump(SD_C_ABAL, c_lo);
        for (int64_t i = c_lo; i < c_hi; i++) c_acctbal[i - c_lo] = tg_draw(&s, -99999, 999999);
    }
    if (ps_availqty) {
        int64_t s = tg_jump(SD_PS_QTY, 4 * p_lo);
        for (int64_t r = 0; r < 4 * (p_hi - p_lo); r++) ps_availqty[r] = (int32_t)tg_draw(&s, 1, 9999);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Supplier text columns of TPC-H Q15 / Q16 / Q20 (cases/tpch/query/q15.sql, q16.sql, q20.sql):
 *   s_address  dbgen a_rnd(10, 40, S_ADDR_SD): a length draw, then one UnifInt(0, MAX_LONG) draw per 5 characters, 6 bits per character
 *              indexing the 64-character alphabet below (the draw is NEGATIVE: dbgen's int32 range wraps; `& 077` and `>>= 6`
 *              then act on the two's complement, which the golden addresses confirm); the stream advances 9 draws per supplier
 *   s_phone    "CC-LLL-LLL-LLLL": CC = 10 + s_nationkey, then UnifInt(100,999), UnifInt(100,999), UnifInt(1000,9999) from S_PHNE_SD
 *              (3 draws per supplier)
 *   complaint  dbgen plants "Customer ... Complaints" into s_comment when UnifInt(1,10000, BBB_CMNT_SD) <= 10 and
 *              UnifInt(0,100, BBB_TYPE_SD) < 50 (else "Customer ... Recommends"); the grammar-generated comment text itself never
 *              holds the word "Customer", so the flag IS `s_comment like '%Customer%Complaints%'`
 * Pinned by the reference's golden cases/tpch/1g/plan/q15.txt (one address + phone), q20.txt (177 addresses) and q16.txt (whose counts
 * imply exactly the complaint suppliers 358, 2820, 3804, 9504 at SF1: the two BBB seeds below are the ones that select them).
 */
enum { SD_S_ADDR = 706178559, SD_S_PHNE = 884434366, SD_BBB_CMNT = 202794285, SD_BBB_TYPE = 753643799 };
static const char TG_ALNUM[65] = "0123456789abcdefghijklmnopqrstuvwxyz ABCDEFGHIJKLMNOPQRSTUVWXYZ,";

/* dbgen a_rnd(10, 40, stream) for row i of a 9-draws-per-row stream; dst holds 41 bytes */
static void tg_address(int64_t seed0, int64_t i, char *dst)
{
    int64_t s = tg_jump(seed0, 9 * i);
    const int64_t len = tg_draw(&s, 10, 40);
    int64_t bits = 0;
    for (int64_t k = 0; k < len; k++) {
        if (k % 5 == 0) {       /* UnifInt(0, MAX_LONG): dbgen computes the range in int32, 2^31-1 - 0 + 1 wraps to -2^31 */
            s = (s * TG_A) % TG_M;
            bits = (int64_t)(((double)s / 2147483647.0) * -2147483648.0);
        }
        dst[k] = TG_ALNUM[bits & 63];
        bits >>= 6;
    }
    dst[len] = 0;
}

/* addr_buf: 41 bytes per supplier (NUL-terminated); phone: 3 ints per supplier; complaint: 1 byte per supplier;
 * acctbal: UnifInt(-99999, 999999) cents from S_ABAL_SD */
enum { SD_S_ABAL = 962338209, SD_C_ADDR = 881155353, SD_C_PHNE = 1521138112 };

void tg_gen_supplier_text(double sf, int64_t s_lo, int64_t s_hi, char *addr_buf, int32_t *phone, uint8_t *complaint, int64_t *acctbal)
{
    (void)sf;
    int64_t s_ph = tg_jump(SD_S_PHNE, 3 * s_lo), s_bc = tg_jump(SD_BBB_CMNT, s_lo), s_bt = tg_jump(SD_BBB_TYPE, s_lo), s_ab = tg_jump(SD_S_ABAL, s_lo);
    for (int64_t i = s_lo; i < s_hi; i++) {
        if (addr_buf) tg_address(SD_S_ADDR, i, addr_buf + 41 * (i - s_lo));
        const int64_t p1 = tg_draw(&s_ph, 100, 999), p2 = tg_draw(&s_ph, 100, 999), p3 = tg_draw(&s_ph, 1000, 9999);
        if (phone) { phone[3 * (i - s_lo)] = (int32_t)p1; phone[3 * (i - s_lo) + 1] = (int32_t)p2; phone[3 * (i - s_lo) + 2] = (int32_t)p3; }
        const int64_t bad = tg_draw(&s_bc, 1, 10000), type = tg_draw(&s_bt, 0, 100), bal = tg_draw(&s_ab, -99999, 999999);
        if (complaint) complaint[i - s_lo] = (uint8_t)(bad <= 10 && type < 50);
        if (acctbal) acctbal[i - s_lo] = bal;
    }
}

/* c_address / c_phone of customers [c_lo, c_hi) (Q10), same construction from C_ADDR_SD / C_PHNE_SD */
void tg_gen_customer_text(double sf, int64_t c_lo, int64_t c_hi, char *addr_buf, int32_t *phone)
{
    (void)sf;
    int64_t s_ph = tg_jump(SD_C_PHNE, 3 * c_lo);
    for (int64_t i = c_lo; i < c_hi; i++) {
        if (addr_buf) tg_address(SD_C_ADDR, i, addr_buf + 41 * (i - c_lo));
        const int64_t p1 = tg_draw(&s_ph, 100, 999), p2 = tg_draw(&s_ph, 100, 999), p3 = tg_draw(&s_ph, 1000, 9999);
        if (phone) { phone[3 * (i - c_lo)] = (int32_t)p1; phone[3 * (i - c_lo) + 1] = (int32_t)p2; phone[3 * (i - c_lo) + 2] = (int32_t)p3; }
    }
}

/* ------------------------------------------------------------------------------------------------
 * dbgen's comment text (every *_comment column; TPC-H Q2 / Q10 print one, Q13 filters on o_comment).
 * dbgen pre-generates ONE 300 MiB pool of pseudo-English from stream 5 and cuts every comment out of it:
 *   pool      sentences appended with one blank between them until 300 * 2^20 bytes are filled (the last one truncated).
 *             sentence = a form of `grammar` (N V T | N V P T | N V N T | N P V N T | N P V P T): N = noun phrase (a form of
 *             `np`: N | J N | J, J N | D J N), V = verb phrase (a form of `vp`: V | X V | V D | X V D), P = preposition + " the "
 *             + noun phrase, T = terminator glued to the previous word.  Every choice is pick_str: UnifInt(1, total weight)
 *             against the cumulative weights of dists.dss, one draw each, all from the same stream.
 *   comment   offset = UnifInt(0, pool size - max), length = UnifInt(min, max) from the column's own stream (2 draws per row),
 *             min / max = int(0.4 * avg) / int(1.6 * avg); avg = 73 (c_comment), 63 (s_comment), 49 (o_comment).
 * Word lists and weights are dists.dss's.  Pinned by the reference's golden q2.txt (100 s_comment values) and q10.txt (20
 * c_comment values): all 120 are reproduced byte for byte, at offsets spread over the whole pool, which fixes every list member's
 * LENGTH and weight; the spelling of one weight-1 preposition ("whithout", dists.dss's own typo) shows in none of the 120 and is
 * fixed by its length only.  Q13's `o_comment not like '%pending%accounts%'` is pinned by q13.txt.
 * Not restated: dbgen overwrites part of s_comment with "Customer ... Complaints / Recommends" for the ~10 in 10000 suppliers
 * tg_gen_supplier_text flags (none of q2.txt's rows); tg_gen_supplier_text's flag is what Q16 needs.
 */
typedef struct { const char *text; int w; } tg_ent;
typedef struct { const tg_ent *e; int n; int cum[48]; } tg_dist;
static const tg_ent TGE_GRAMMAR[] = {{"N V T",3},{"N V P T",3},{"N V N T",3},{"N P V N T",1},{"N P V P T",1}};
static const tg_ent TGE_NP[] = {{"N",10},{"J N",20},{"J, J N",10},{"D J N",50}};
static const tg_ent TGE_VP[] = {{"V",30},{"X V",1},{"V D",40},{"X V D",1}};
static const tg_ent TGE_NOUNS[] = {{"packages",40},{"requests",40},{"accounts",40},{"deposits",40},{"foxes",20},{"ideas",20},{"theodolites",20},
    {"pinto beans",20},{"instructions",20},{"dependencies",10},{"excuses",10},{"platelets",10},{"asymptotes",10},{"courts",5},{"dolphins",5},
    {"multipliers",1},{"sauternes",1},{"warthogs",1},{"frets",1},{"dinos",1},{"attainments",1},{"somas",1},{"Tiresias",1},{"patterns",1},{"forges",1},
    {"braids",1},{"frays",1},{"warhorses",1},{"dugouts",1},{"notornis",1},{"epitaphs",1},{"pearls",1},{"tithes",1},{"waters",1},{"orbits",1},{"gifts",1},
    {"sheaves",1},{"depths",1},{"sentiments",1},{"decoys",1},{"realms",1},{"pains",1},{"grouches",1},{"escapades",1},{"hockey players",1}};
static const tg_ent TGE_VERBS[] = {{"sleep",20},{"wake",20},{"are",20},{"cajole",20},{"haggle",20},{"nag",10},{"use",10},{"boost",10},{"affix",5},
    {"detect",5},{"integrate",5},{"maintain",1},{"nod",1},{"was",1},{"lose",1},{"sublate",1},{"solve",1},{"thrash",1},{"promise",1},{"engage",1},
    {"hinder",1},{"print",1},{"x-ray",1},{"breach",1},{"eat",1},{"grow",1},{"impress",1},{"mold",1},{"poach",1},{"serve",1},{"run",1},{"dazzle",1},
    {"snooze",1},{"doze",1},{"unwind",1},{"kindle",1},{"play",1},{"hang",1},{"believe",1},{"doubt",1}};
static const tg_ent TGE_ADJ[] = {{"special",20},{"pending",20},{"unusual",20},{"express",20},{"furious",1},{"sly",1},{"careful",1},{"blithe",1},
    {"quick",1},{"fluffy",1},{"slow",1},{"quiet",1},{"ruthless",1},{"thin",1},{"close",1},{"dogged",1},{"daring",1},{"brave",1},{"stealthy",1},
    {"permanent",1},{"enticing",1},{"idle",1},{"busy",1},{"regular",50},{"final",40},{"ironic",40},{"even",30},{"bold",20},{"silent",10}};
static const tg_ent TGE_ADV[] = {{"sometimes",1},{"always",1},{"never",1},{"furiously",50},{"slyly",50},{"carefully",50},{"blithely",40},
    {"quickly",30},{"fluffily",20},{"slowly",1},{"quietly",1},{"ruthlessly",1},{"thinly",1},{"closely",1},{"doggedly",1},{"daringly",1},{"bravely",1},
    {"stealthily",1},{"permanently",1},{"enticingly",1},{"idly",1},{"busily",1},{"regularly",1},{"finally",1},{"ironically",1},{"evenly",1},
    {"boldly",1},{"silently",1}};
static const tg_ent TGE_PREP[] = {{"about",50},{"above",50},{"according to",50},{"across",50},{"after",50},{"against",40},{"along",40},
    {"alongside of",30},{"among",30},{"around",20},{"at",10},{"atop",1},{"before",1},{"behind",1},{"beneath",1},{"beside",1},{"besides",1},
    {"between",1},{"beyond",1},{"by",1},{"despite",1},{"during",1},{"except",1},{"for",1},{"from",1},{"in place of",1},{"inside",1},{"instead of",1},
    {"into",1},{"near",1},{"of",1},{"on",1},{"outside",1},{"over",1},{"past",1},{"since",1},{"through",1},{"throughout",1},{"to",1},{"toward",1},
    {"under",1},{"until",1},{"up",1},{"upon",1},{"whithout",1},{"with",1},{"within",1}};
static const tg_ent TGE_AUX[] = {{"do",1},{"may",1},{"might",1},{"shall",1},{"will",1},{"would",1},{"can",1},{"could",1},{"should",1},{"ought to",1},
    {"must",1},{"will have to",1},{"shall have to",1},{"could have to",1},{"should have to",1},{"must have to",1},{"need to",1},{"try to",1}};
static const tg_ent TGE_TERM[] = {{".",50},{";",1},{":",1},{"?",1},{"!",1},{"--",1}};
#define TG_DIST(name, arr) static tg_dist name = { arr, (int)(sizeof(arr) / sizeof(arr[0])), {0} }
TG_DIST(TGD_GRAMMAR, TGE_GRAMMAR); TG_DIST(TGD_NP, TGE_NP); TG_DIST(TGD_VP, TGE_VP); TG_DIST(TGD_NOUNS, TGE_NOUNS); TG_DIST(TGD_VERBS, TGE_VERBS);
TG_DIST(TGD_ADJ, TGE_ADJ); TG_DIST(TGD_ADV, TGE_ADV); TG_DIST(TGD_PREP, TGE_PREP); TG_DIST(TGD_AUX, TGE_AUX); TG_DIST(TGD_TERM, TGE_TERM);

enum { SD_TEXT_POOL = 933588178, SD_C_CMNT = 1335826707, SD_S_CMNT = 1341315363, SD_O_CMNT = 276090261 };
#define TG_POOL_SIZE (300LL * 1024 * 1024)
static char *tg_pool;
static int64_t tg_pool_seed;

static int tg_pick(tg_dist *d, char *target)
{
    const int64_t j = tg_draw(&tg_pool_seed, 1, d->cum[d->n - 1]);
    int i = 0;
    while (d->cum[i] < j) i++;
    strcpy(target, d->e[i].text);
    return i;
}

/* a noun or verb phrase: every word followed by the punctuation glued to its form letter (the comma of "J,") and one blank */
static int tg_phrase(char *dest, tg_dist *forms)
{
    char syntax[16];
    int res = 0;
    tg_pick(forms, syntax);
    for (const char *c = syntax; *c; ) {
        while (*c == ' ') c++;
        if (!*c) break;
        tg_dist *src = *c == 'D' ? &TGD_ADV : *c == 'V' ? &TGD_VERBS : *c == 'X' ? &TGD_AUX : *c == 'J' ? &TGD_ADJ : &TGD_NOUNS;
        const int len = (int)strlen(src->e[tg_pick(src, dest)].text);
        dest += len; res += len;
        c++;
        if (*c && *c != ' ') { *dest++ = *c++; res++; }
        *dest++ = ' '; res++;
    }
    return res;
}

static int tg_sentence(char *dest)
{
    char syntax[16];
    int res = 0;
    tg_pick(&TGD_GRAMMAR, syntax);
    for (const char *c = syntax; *c; c++) {
        if (*c == ' ') continue;
        int len = 0;
        if (*c == 'V') len = tg_phrase(dest, &TGD_VP);
        else if (*c == 'N') len = tg_phrase(dest, &TGD_NP);
        else if (*c == 'P') {
            len = (int)strlen(TGD_PREP.e[tg_pick(&TGD_PREP, dest)].text);
            memcpy(dest + len, " the ", 5);
            len += 5;
            len += tg_phrase(dest + len, &TGD_NP);
        } else {                                        /* 'T': the terminator replaces the blank after the last word */
            dest--;
            len = (int)strlen(TGD_TERM.e[tg_pick(&TGD_TERM, dest)].text);
            res--;
        }
        dest += len; res += len;
    }
    *dest = 0;
    return res;
}

/* the 300 MiB pool (built once, about 3 s); NULL when the allocation fails */
const char *tg_text_pool(void)
{
    if (tg_pool) return tg_pool;
    tg_dist *all[] = {&TGD_GRAMMAR, &TGD_NP, &TGD_VP, &TGD_NOUNS, &TGD_VERBS, &TGD_ADJ, &TGD_ADV, &TGD_PREP, &TGD_AUX, &TGD_TERM};
    for (int k = 0; k < 10; k++) { int c = 0; for (int i = 0; i < all[k]->n; i++) { c += all[k]->e[i].w; all[k]->cum[i] = c; } }
    char *pool = (char *)malloc((size_t)TG_POOL_SIZE + 1), *cp = pool, sentence[512];
    if (!pool) return NULL;
    tg_pool_seed = SD_TEXT_POOL;
    int64_t filled = 0;
    while (filled < TG_POOL_SIZE) {
        const int len = tg_sentence(sentence);
        const int64_t room = TG_POOL_SIZE - filled;
        if (room >= len + 1) { memcpy(cp, sentence, (size_t)len); cp += len; *cp++ = ' '; filled += len + 1; }
        else { memcpy(cp, sentence, (size_t)room); cp += room; filled += room; }
    }
    *cp = 0;
    tg_pool = pool;
    return tg_pool;
}

/* comment spans of rows [lo, hi) of the column whose stream starts at `seed` (2 draws per row) and whose average length is avg */
void tg_comment_spans(int64_t seed, int avg, int64_t lo, int64_t hi, int64_t *off, int32_t *len)
{
    const int mn = (int)(avg * 0.4), mx = (int)(avg * 1.6);
    int64_t s = tg_jump(seed, 2 * lo);
    for (int64_t i = lo; i < hi; i++) {
        off[i - lo] = tg_draw(&s, 0, TG_POOL_SIZE - mx);
        len[i - lo] = (int32_t)tg_draw(&s, mn, mx);
    }
}

/* flags[i] = comment i matches '%w1%w2%' (w2 after the end of the first w1; wildcardMatch's greedy-with-backtracking result for
 * two literal words is the same as: some occurrence of w1 is followed by an occurrence of w2, i.e. the FIRST w1 is) */
int tg_comments_like2(int64_t seed, int avg, int64_t lo, int64_t hi, const char *w1, const char *w2, uint8_t *flags)
{
    const char *pool = tg_text_pool();
    if (!pool) return -1;
    const int mn = (int)(avg * 0.4), mx = (int)(avg * 1.6);
    const size_t n1 = strlen(w1), n2 = strlen(w2);
    int64_t s = tg_jump(seed, 2 * lo);
    char buf[256];
    for (int64_t i = lo; i < hi; i++) {
        const int64_t off = tg_draw(&s, 0, TG_POOL_SIZE - mx);
        const int len = (int)tg_draw(&s, mn, mx);
        memcpy(buf, pool + off, (size_t)len);
        buf[len] = 0;
        const char *p = strstr(buf, w1);
        flags[i - lo] = (uint8_t)(p != NULL && strstr(p + n1, w2) != NULL);
        (void)n2;
    }
    return 0;
}
