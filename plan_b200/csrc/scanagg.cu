// scanagg.cu -- `Agg <- Scan[filters]` pipelines: shape matching, launch, exact finalise.
//
// Reference operators replaced: aggExecutor over scanExecutor with pushed-down filters
// (/root/reference/pkg/compute/executor_aggr.go:106-262, executor_scan.go:225-241);
// plan shapes per SURVEY.md 3.4 (Q6: const group + sum(DEC*DEC); Q1: 2 x VARCHAR(1) keys,
// 8 aggregates).
#include <algorithm>

#include "hostdec.hpp"
#include "pipeline.hpp"
#include "scanagg.cuh"
#include "rowvm_compile.hpp"
#include "scanagg_vm.cuh"

namespace pg {

static i128 maxabs(i64 lo, i64 hi)
{
    i128 a = lo < 0 ? -(i128)lo : (i128)lo, b = hi < 0 ? -(i128)hi : (i128)hi;
    return a > b ? a : b;
}

static Range find_range(const std::vector<Range> &rs, int col)
{
    for (auto &r : rs) if (r.col == col) return r;
    Range r;
    r.col = col;
    return r;
}

// A logical inclusive range on column c -> bounds on the STORED values (logical = base + stored), clamped to
// what the stored type can hold; an empty range becomes [1, 0].
static void stored_range(const Column &c, i64 lo, i64 hi, i64 *slo, i64 *shi)
{
    const int pw = c.phys_width();
    const i128 tmin = pw == 8 ? (i128)INT64_MIN : pw == 4 ? (i128)INT32_MIN : 0;
    const i128 tmax = pw == 8 ? (i128)INT64_MAX : pw == 4 ? (i128)INT32_MAX : pw == 2 ? 0xffff : 0xff;
    i128 a = (i128)lo - c.base, b = (i128)hi - c.base;
    if (a < tmin) a = tmin;
    if (b > tmax) b = tmax;
    if (lo > hi || a > b) { *slo = 1; *shi = 0; return; }
    *slo = (i64)a;
    *shi = (i64)b;
}

static bool fits_i32(i128 v) { return v >= (i128)INT32_MIN && v <= (i128)INT32_MAX; }
// can the column's LOGICAL values be handled as 32-bit ints by a narrow kernel?
static bool col_is_narrow(const Column &c)
{
    return c.phys_width() <= 4 && fits_i32(c.vmin) && fits_i32(c.vmax) && fits_i32(c.base);
}

static int sms_times(const void *kernel, int threads, size_t smem)
{
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    return ctx().prop.multiProcessorCount * per_sm;
}
static int grid_for(const void *kernel, int threads, size_t smem, i64 ntiles)
{
    i64 g = sms_times(kernel, threads, smem);
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    return (int)g;
}

static pg_decimal to_pg_decimal(const HDec &d)
{
    pg_decimal o;
    o.coef = d.coef;
    o.scale = d.scale;
    o.neg = d.neg ? 1u : 0u;
    return o;
}

static i128 make_i128(u64 lo, u64 hi) { return (i128)(((u128)hi << 64) | (u128)lo); }

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// ------------------------------------------------------------------- sumprod --

typedef void (*SumProdKernel)(const SumProdParams, i64 *);
static SumProdKernel sumprod_variant(bool wide, bool has_a, bool has_b)
{
#define PG_SP(W, U) (has_a && has_b ? sumprod_kernel<W, true, true, U> : has_a ? sumprod_kernel<W, true, false, U> \
                     : has_b ? sumprod_kernel<W, false, true, U> : sumprod_kernel<W, false, false, U>)
    return wide ? PG_SP(true, 2) : PG_SP(false, 4);      // 32-byte raw vectors: two tiles in flight keep the wide kernel under 128 registers
#undef PG_SP
}

// staged (bulk-copy fed) variants over narrow columns, scan_staged.cuh
typedef void (*SumProdSKernel)(const SumProdSParams, i64 *);
// Width signatures with their own instantiation (everything else runs the run-time-width variant, W = -1):
// {fa, fb, pa, pb} = {4, 1, 2, 1}: DECIMAL price x byte-wide factor under a 16-bit date and a byte predicate (TPC-H Q6 at any SF)
static SumProdSKernel sumprod_staged_variant(const int *w /* [4] role widths */, bool xr, bool yr, int qpt, bool *specialised)
{
    const bool has_a = w[2] != 0, has_b = w[3] != 0;
#define PG_SPSQ(FA, FB, PA, PB, X, Y) (qpt >= 4 ? sumprod_staged_kernel<FA, FB, PA, PB, X, Y, 4> : qpt >= 2 ? sumprod_staged_kernel<FA, FB, PA, PB, X, Y, 2> : sumprod_staged_kernel<FA, FB, PA, PB, X, Y, 1>)
#define PG_SPSR(FA, FB, PA, PB) (xr ? (yr ? PG_SPSQ(FA, FB, PA, PB, true, true) : PG_SPSQ(FA, FB, PA, PB, true, false)) : (yr ? PG_SPSQ(FA, FB, PA, PB, false, true) : PG_SPSQ(FA, FB, PA, PB, false, false)))
    *specialised = true;
    if (w[0] == 4 && w[1] == 1 && w[2] == 2 && w[3] == 1) return PG_SPSR(4, 1, 2, 1);
    *specialised = false;
    return has_a ? (has_b ? PG_SPSR(-1, -1, -1, -1) : PG_SPSR(-1, -1, -1, 0)) : (has_b ? PG_SPSR(-1, -1, 0, -1) : PG_SPSR(-1, -1, 0, 0));
#undef PG_SPSR
#undef PG_SPSQ
}

// (v - lo) <=u span on a STORED value; false when the range is empty
static bool unsigned_range(i64 slo, i64 shi, unsigned *lo, unsigned *span)
{
    if (slo > shi) return false;
    *lo = (unsigned)(int32_t)slo;
    *span = (unsigned)(u64)(shi - slo);
    return true;
}
// does [slo, shi] leave out any stored value the column holds?
static bool range_restricts(const Column &c, i64 slo, i64 shi)
{
    return slo > (i64)((i128)c.vmin - c.base) || shi < (i64)((i128)c.vmax - c.base);
}
// add a physical column to a stage description once; returns its index
static int stage_add(StageDesc *d, const Column &c)
{
    for (int i = 0; i < d->ncol; i++) if (d->src[i] == (const char *)c.d_data) return i;
    if (d->ncol >= ST_MAXCOL) return -1;
    d->src[d->ncol] = (const char *)c.d_data;
    d->pw[d->ncol] = c.phys_width();
    return d->ncol++;
}

struct SumProdPipeline : Pipeline {
    const pg_table *table = nullptr;
    SumProdParams prm{};
    SumProdSParams sprm{};
    bool staged = false, xr = false, yr = false, specialised = false;
    SumProdSKernel skern = nullptr;
    int qpt = 2;
    size_t smem = 0;
    bool has_a = false, has_b = false, wide = false;
    int grid = 1, vscale = 0;
    AggExpr agg;
    std::vector<std::pair<int, int>> outs;
    int bytes_per_row = 16;
    DevBuf d_part, d_final, d_gather;
    PinBuf h_final;
    EventPair ev_all, ev_main;

    int run(pg_result *res) override
    {
        Context &c = ctx();
        cudaStream_t st = c.stream;
        PG_TRY(ev_all.init());
        PG_TRY(ev_main.init());
        PG_CUDA(cudaEventRecord(ev_all.a, st));
        PG_CUDA(cudaEventRecord(ev_main.a, st));
        if (staged) skern<<<grid, ST_THREADS, smem, st>>>(sprm, d_part.as<i64>());
        else sumprod_variant(wide, has_a, has_b)<<<grid, SA_THREADS, 0, st>>>(prm, d_part.as<i64>());
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaEventRecord(ev_main.b, st));
        finalize128_kernel<<<2, 32, 0, st>>>(d_part.as<i64>(), grid, 2, d_final.as<u64>());
        PG_CUDA(cudaGetLastError());
        const void *src = d_final.p;
        const int nmerge = table->dist == PG_DIST_REPLICATED ? 1 : c.world;   // a replicated table is complete on every rank
        if (nmerge > 1) {
            PG_TRY(comm_allgather(d_final.p, d_gather.p, 32, st));
            src = d_gather.p;
        }
        PG_CUDA(cudaMemcpyAsync(h_final.p, src, 32 * (size_t)nmerge, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        PG_CUDA(cudaStreamSynchronize(st));
        const u64 *h = h_final.as<u64>();
        i128 sum = 0, cnt = 0;
        for (int r = 0; r < nmerge; r++) {   // merged in rank order (shards are contiguous row ranges)
            sum += make_i128(h[4 * r], h[4 * r + 1]);
            cnt += make_i128(h[4 * r + 2], h[4 * r + 3]);
        }
        res->stats.kernel_ms = ev_all.ms();
        res->stats.main_kernel_ms = ev_main.ms();
        res->stats.rows_scanned = table->nrows;
        res->stats.algorithmic_bytes = table->nrows * bytes_per_row;
        res->stats.main_kernel_bytes = res->stats.algorithmic_bytes;
        res->stats.kernel_launches = 2;
        res->stats.aux[0] = (i64)cnt;
        // the reference emits no row at all when nothing reached the aggregate
        // (aggregate_exec.go:160-185)
        res->nrows = cnt > 0 ? 1 : 0;
        for (auto &o : outs) {
            (void)o;
            ResCol col;
            col.type = PG_T_DECIMAL128;
            col.width = agg.width;
            col.scale = agg.scale;
            if (cnt > 0) {
                HDec d;
                // an ungrouped total beyond 19 digits would need the sequential half-even emulation of the
                // low-cardinality pipeline; refuse instead of rounding differently
                if (!hd_from_i128(sum, vscale, &d) || hd_digits((u128)(sum < 0 ? -sum : sum)) > HD_MAXPREC)
                    PG_FAIL(PG_EOVERFLOW, "sum exceeds 19 significant digits (order-dependent rounding regime)");
                col.push(to_pg_decimal(d));
            }
            res->cols.push_back(col);
        }
        return PG_OK;
    }
};

static int try_sumprod(pg_plan *plan, const Node &aggn, const Node &scan, const std::vector<Range> &ranges,
                       const std::vector<AffProd> &args, std::unique_ptr<Pipeline> *out, std::string *why)
{
    const pg_table *t = plan->slots[(size_t)scan.slot];
    if (!aggn.groups.empty()) { *why = "has group keys"; return PG_EUNSUPPORTED; }
    if (aggn.aggs.size() != 1 || aggn.aggs[0].fn != PG_AGG_SUM || aggn.aggs[0].ltype != PG_LT_DECIMAL) { *why = "not a single DECIMAL sum"; return PG_EUNSUPPORTED; }
    const AffProd &ap = args[0];
    if (ap.f.size() != 2 || ap.f[0].c != 0 || ap.f[0].s != 1 || ap.f[1].c != 0 || ap.f[1].s != 1) { *why = "argument is not column*column"; return PG_EUNSUPPORTED; }
    const Column &ca = t->cols[(size_t)ap.f[0].col], &cb = t->cols[(size_t)ap.f[1].col];
    std::unique_ptr<SumProdPipeline> p(new SumProdPipeline());
    p->table = t;
    p->agg = aggn.aggs[0];
    p->outs = aggn.outs;
    for (auto &o : aggn.outs) if (o.first != 1 || o.second != 0) { *why = "output list refers to something else than the aggregate"; return PG_EUNSUPPORTED; }
    p->vscale = ap.vscale();
    SumProdParams &q = p->prm;
    q.nrows = t->nrows;
    q.fa = ca.ncol();
    q.fb = cb.ncol();
    Range ra = find_range(ranges, ap.f[0].col), rb = find_range(ranges, ap.f[1].col);
    stored_range(ca, ra.lo, ra.hi, &q.fa_lo, &q.fa_hi);
    stored_range(cb, rb.lo, rb.hi, &q.fb_lo, &q.fb_hi);
    q.pa = q.pb = q.fa;
    q.a_lo = q.b_lo = 0;
    q.a_hi = q.b_hi = -1;
    bool narrow = col_is_narrow(ca) && col_is_narrow(cb);
    p->bytes_per_row = ca.phys_width() + cb.phys_width();
    int npred = 0;
    for (auto &r : ranges) {
        if (r.col == ap.f[0].col || r.col == ap.f[1].col) continue;
        const Column &col = t->cols[(size_t)r.col];
        if (!is_int_family(col.type)) { *why = "predicate on a column that is neither a factor nor an integer-family column"; return PG_EUNSUPPORTED; }
        if (npred == 2) { *why = "more than two predicate-only columns"; return PG_EUNSUPPORTED; }
        if (col.phys_width() > 4) narrow = false;
        if (npred == 0) { q.pa = col.ncol(); stored_range(col, r.lo, r.hi, &q.a_lo, &q.a_hi); p->has_a = true; }
        else { q.pb = col.ncol(); stored_range(col, r.lo, r.hi, &q.b_lo, &q.b_hi); p->has_b = true; }
        p->bytes_per_row += col.phys_width();
        npred++;
    }
    if (env_int("PG_FORCE_WIDE", 0)) narrow = false;
    p->wide = !narrow;
    i64 ntiles = (t->nrows + SA_TILE - 1) / SA_TILE;
    const void *kern = (const void *)sumprod_variant(p->wide, p->has_a, p->has_b);
    int full = sms_times(kern, SA_THREADS, 0);
    i64 tile_rows = SA_TILE;
    // narrow columns: the bulk-copy staged kernel (scan_staged.cuh)
    if (narrow && !env_int("PG_NO_STAGED", 0) && t->nrows > 0 && ca.vmin >= 0 && cb.vmin >= 0) {
        SumProdSParams &sp = p->sprm;
        sp = SumProdSParams{};
        sp.nrows = t->nrows;
        sp.xbase = (int)ca.base;
        sp.ybase = (int)cb.base;
        bool ok = unsigned_range(q.fa_lo, q.fa_hi, &sp.x_lo, &sp.x_span) && unsigned_range(q.fb_lo, q.fb_hi, &sp.y_lo, &sp.y_span);
        if (p->has_a) ok = ok && unsigned_range(q.a_lo, q.a_hi, &sp.a_lo, &sp.a_span);
        if (p->has_b) ok = ok && unsigned_range(q.b_lo, q.b_hi, &sp.b_lo, &sp.b_span);
        const Column *rc[4] = {&ca, &cb, nullptr, nullptr};
        {
            int np = 0;
            for (auto &r : ranges) {
                if (r.col == ap.f[0].col || r.col == ap.f[1].col) continue;
                rc[2 + np++] = &t->cols[(size_t)r.col];
            }
        }
        for (int r = 0; r < 4 && ok; r++) {
            if (!rc[r]) { sp.rpw[r] = 0; sp.roff[r] = 0; continue; }
            const int ci = stage_add(&sp.st, *rc[r]);
            if (ci < 0) { ok = false; break; }
            sp.rpw[r] = -1 - ci;            // resolved to offsets after the layout below
        }
        if (ok) {
            p->qpt = std::max(1, std::min(4, env_int("PG_QPT", 2)));
            if (p->qpt == 3) p->qpt = 2;
            tile_rows = (i64)ST_CONS_WARPS * 128 * p->qpt;
            stage_layout(&sp.st, (int)tile_rows);
            for (int r = 0; r < 4; r++) if (sp.rpw[r] < 0) { const int ci = -1 - sp.rpw[r]; sp.rpw[r] = sp.st.pw[ci]; sp.roff[r] = sp.st.off[ci]; }
            sp.st.nstage = std::max(2, std::min(ST_MAXSTAGE, env_int("PG_NSTAGE", 4)));
            p->xr = range_restricts(ca, q.fa_lo, q.fa_hi);
            p->yr = range_restricts(cb, q.fb_lo, q.fb_hi);
            p->smem = (size_t)ST_HDR + (size_t)sp.st.nstage * sp.st.stage_bytes;
            if (!p->has_a && p->has_b) {            // a single predicate column is role 2
                std::swap(sp.rpw[2], sp.rpw[3]); std::swap(sp.roff[2], sp.roff[3]);
                std::swap(sp.a_lo, sp.b_lo); std::swap(sp.a_span, sp.b_span);
            }
            int wsig[4] = {sp.rpw[0], sp.rpw[1], sp.rpw[2], sp.rpw[3]};
            if (env_int("PG_NO_SPECIALISED", 0)) wsig[0] = 8;
            p->skern = sumprod_staged_variant(wsig, p->xr, p->yr, p->qpt, &p->specialised);
            kern = (const void *)p->skern;
            PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
            full = sms_times(kern, ST_THREADS, p->smem);
            ntiles = (t->nrows + tile_rows - 1) / tile_rows;
            p->staged = true;
        }
    }
    p->grid = (int)std::max<i64>(1, std::min<i64>(full, ntiles));
    // per-CTA int64 partials must be exact: bound them with the statistics of the whole table (every rank takes
    // the same decision) and the largest shard
    i128 per_row = maxabs(std::max(ra.lo, ca.gmin()), std::min(ra.hi, ca.gmax())) * maxabs(std::max(rb.lo, cb.gmin()), std::min(rb.hi, cb.gmax()));
    const i64 max_tiles = (t->max_rows() + tile_rows - 1) / tile_rows;
    const i64 gmin_grid = std::max<i64>(1, std::min<i64>(full, max_tiles));
    i128 rows_per_cta = (i128)((max_tiles + gmin_grid - 1) / gmin_grid) * tile_rows;       // a CTA takes ceil(ntiles / grid) tiles
    if (ra.lo <= ra.hi && rb.lo <= rb.hi && per_row * rows_per_cta >= ((i128)1 << 62)) {
        *why = "per-CTA partial sum could exceed int64";
        return PG_EUNSUPPORTED;
    }
    PG_TRY(p->d_part.alloc(sizeof(i64) * 2 * (size_t)p->grid));
    PG_TRY(p->d_final.alloc(32));
    PG_TRY(p->d_gather.alloc(32 * (size_t)ctx().world));
    PG_TRY(p->h_final.alloc(32 * (size_t)ctx().world));
    char buf[512];
    snprintf(buf, sizeof buf,
             "ScanAgg[sumprod] table=%s rows=%lld kernel=%s<%s,%d,%d> grid=%d block=%d "
             "stored bytes/row=%d (widths: fa=%d fb=%d) stored ranges: a=[%lld,%lld] b=[%lld,%lld] fa=[%lld,%lld] fb=[%lld,%lld] value_scale=%d",
             t->name.c_str(), (long long)t->nrows, p->staged ? "sumprod_staged_kernel" : "sumprod_kernel", p->staged ? (p->specialised ? "bulk-copy ring, widths compiled in" : "bulk-copy ring, run-time widths") : p->wide ? "wide" : "narrow",
             (int)p->has_a, (int)p->has_b, p->grid, p->staged ? ST_THREADS : SA_THREADS,
             p->bytes_per_row, ca.phys_width(), cb.phys_width(), (long long)q.a_lo, (long long)q.a_hi, (long long)q.b_lo, (long long)q.b_hi,
             (long long)q.fa_lo, (long long)q.fa_hi, (long long)q.fb_lo, (long long)q.fb_hi, p->vscale);
    p->explain = buf;
    *out = std::move(p);
    return PG_OK;
}

// ------------------------------------------------------------- lowcard chain --

typedef void (*LowcardKernel)(const LowcardParams, i64 *, i64 *);
typedef void (*LowcardSKernel)(const LowcardSParams, i64 *);
// {pred, q, A, B, C} = {2, 1, 4, 1, 1} has its own instantiation (TPC-H Q1 at any SF); everything else: run-time widths
static LowcardSKernel lowcard_staged_variant(const int *w /* [7] role widths */, bool key1, int qpt, bool *specialised)
{
#define PG_LCS(P, Q, A, B, C) (qpt >= 2 ? (key1 ? lowcard_staged_kernel<P, Q, A, B, C, true, 2> : lowcard_staged_kernel<P, Q, A, B, C, false, 2>) \
                                        : (key1 ? lowcard_staged_kernel<P, Q, A, B, C, true, 1> : lowcard_staged_kernel<P, Q, A, B, C, false, 1>))
    *specialised = true;
    if (w[0] == 2 && w[3] == 1 && w[4] == 4 && w[5] == 1 && w[6] == 1) return PG_LCS(2, 1, 4, 1, 1);
    *specialised = false;
    return PG_LCS(-1, -1, -1, -1, -1);
#undef PG_LCS
}
static int bit_length(i128 v) { int n = 0; while (v > 0) { n++; v >>= 1; } return n; }
static LowcardKernel lowcard_variant(bool wide, bool acc32, bool key1, int unroll)
{
#define PG_LC(U) (wide ? (key1 ? lowcard_chain_kernel<true, false, true, U> : lowcard_chain_kernel<true, false, false, U>) \
                  : acc32 ? (key1 ? lowcard_chain_kernel<false, true, true, U> : lowcard_chain_kernel<false, true, false, U>) \
                          : (key1 ? lowcard_chain_kernel<false, false, true, U> : lowcard_chain_kernel<false, false, false, U>))
    return unroll >= 4 ? PG_LC(4) : PG_LC(2);
#undef PG_LC
}

// Rank-ordered exact merge of the ranks' partial aggregates as they come out of the all-gather: per rank
// [nvals x {u64 lo, u64 hi}] 128-bit totals followed by [ngroups] first-row ids (0x7f7f.. = group absent on that rank).
// Sums add in 128 bits (NCCL has no int128 sum), the first row of a group is the smallest over the ranks.
static void merge_rank_partials(const char *gathered, size_t rank_bytes, int nranks, int nvals, int ngroups, i128 *tot, i64 *first)
{
    for (int v = 0; v < nvals; v++) tot[v] = 0;
    for (int g = 0; g < ngroups; g++) first[g] = INT64_MAX;
    for (int r = 0; r < nranks; r++) {
        const char *base = gathered + rank_bytes * (size_t)r;
        const u64 *h = (const u64 *)base;
        const i64 *f = (const i64 *)(base + (size_t)nvals * 16);
        for (int v = 0; v < nvals; v++) tot[v] += make_i128(h[2 * v], h[2 * v + 1]);
        for (int g = 0; g < ngroups; g++) if (f[g] != 0x7f7f7f7f7f7f7f7fLL && f[g] < first[g]) first[g] = f[g];
    }
}

struct LowcardPipeline : Pipeline {
    const pg_table *table = nullptr;
    LowcardParams prm{};
    bool has_key1 = false, wide = false, acc32 = false;
    int unroll = 2;
    int grid = 1, G = 1;
    size_t smem = 0;
    bool staged = false, specialised = false;   // lowcard_staged_kernel (scan_staged.cuh) + first_rows_kernel
    LowcardSKernel skern = nullptr;
    LowcardSParams sprm{};
    int qpt = 2;
    i64 lc_per = 0;                      // LC_TILE-sized tiles per CTA (contiguous mode), whichever kernel runs
    int nkeys = 0;
    int key_col[2] = {-1, -1};
    std::vector<uint8_t> vals[2];        // dense id -> byte code, per key
    std::vector<AggExpr> aggs;
    std::vector<int> slot;               // accumulator slot per aggregate
    std::vector<int> slot_scale;         // value scale per accumulator slot
    std::vector<bool> agg_is_int;        // integer (HUGEINT / DOUBLE) vs DECIMAL result
    std::vector<std::pair<int, int>> outs;
    int bytes_per_row = 34;
    DevBuf d_part, d_final, d_first, d_luts, d_gather;
    PinBuf h_final;
    EventPair ev_all, ev_main;
    // the ordered-rounding stage (scanagg.cuh: ord_plan / ord_jobs / ord_fold), appended when the statistics allow
    // a DECIMAL total of 20 digits; identical launches and collectives on every rank
    bool ord_stage = false;
    unsigned emu_mask = 0;               // slots whose sum is a DECIMAL result with non-negative addends and scale >= 1
    i64 ord_stride = 0;                  // summaries reserved per job
    DevBuf d_jobs, d_ord, d_contrib;     // OrdJob[ORD_MAXJOBS] + njobs | summaries | OrdContrib[ORD_MAXJOBS] + gathered copies
    int emulations = 0;                  // (group, slot) sums that took the ordered path in the last run

    // ranks whose partials are merged: a replicated table is complete on every rank
    int nranks() const { return table->dist == PG_DIST_REPLICATED ? 1 : ctx().world; }
    int myrank() const { return table->dist == PG_DIST_REPLICATED ? 0 : ctx().rank; }

    size_t rank_bytes() const { return (size_t)G * LC_K * 16 + (size_t)LC_MAXG * 8; }
    size_t contrib_bytes() const { return sizeof(OrdContrib) * ORD_MAXJOBS; }
    // pinned result block: [nranks x rank_bytes] totals, then [nranks x contrib_bytes] contributions
    size_t host_bytes() const { return (rank_bytes() + contrib_bytes()) * (size_t)nranks(); }

    int run(pg_result *res) override
    {
        Context &c = ctx();
        cudaStream_t st = c.stream;
        PG_TRY(ev_all.init());
        PG_TRY(ev_main.init());
        Trace tr("lowcard");
        // d_final layout: [G*K][2] u64 totals followed by first_row[LC_MAXG]
        i64 *d_firstrow = (i64 *)((char *)d_final.p + (size_t)G * LC_K * 16);
        PG_CUDA(cudaEventRecord(ev_all.a, st));
        PG_CUDA(cudaMemsetAsync(d_firstrow, 0x7f, LC_MAXG * 8, st));   // 0x7f7f.. = "unset"
        PG_CUDA(cudaEventRecord(ev_main.a, st));
        if (staged) skern<<<grid, ST_THREADS, smem, st>>>(sprm, d_part.as<i64>());
        else lowcard_variant(wide, acc32, has_key1, unroll)<<<grid, LC_THREADS, smem, st>>>(prm, d_part.as<i64>(), d_firstrow);
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaEventRecord(ev_main.b, st));
        finalize128_kernel<<<G * LC_K, 32, 0, st>>>(d_part.as<i64>(), grid, G * LC_K, d_final.as<u64>());
        PG_CUDA(cudaGetLastError());
        if (staged) {       // group order: first passing row of every group, found outside the scan's inner loop
            if (has_key1) first_rows_kernel<true><<<64, 256, 0, st>>>(prm, d_final.as<u64>(), d_firstrow);
            else first_rows_kernel<false><<<64, 256, 0, st>>>(prm, d_final.as<u64>(), d_firstrow);
            PG_CUDA(cudaGetLastError());
        }
        // every rank's totals on every rank (world 1: a device copy)
        if (nranks() > 1) PG_TRY(comm_allgather(d_final.p, d_gather.p, rank_bytes(), st));
        else PG_CUDA(cudaMemcpyAsync(d_gather.p, d_final.p, rank_bytes(), cudaMemcpyDeviceToDevice, st));
        int launches = staged ? 3 : 2;
        char *h_tot = (char *)h_final.p, *h_con = h_tot + rank_bytes() * (size_t)nranks();
        if (ord_stage) {
            const i64 ntiles = (prm.nrows + LC_TILE - 1) / LC_TILE;
            const i64 per = lc_per;
            OrdJob *jobs = d_jobs.as<OrdJob>();
            int *njobs = (int *)(jobs + ORD_MAXJOBS);
            OrdContrib *mine = d_contrib.as<OrdContrib>(), *all = mine + ORD_MAXJOBS;
            ord_plan_kernel<<<1, ORD_PLAN_THREADS, 0, st>>>(d_gather.as<u64>(), (i64)(rank_bytes() / 8), nranks(), myrank(), G, emu_mask, d_part.as<i64>(), grid,
                                              per, ntiles, jobs, njobs);
            const int og = ctx().prop.multiProcessorCount * 4;
            const bool ow = prm.pred.pw > 4 || prm.A.pw > 4 || prm.B.pw > 4 || prm.C.pw > 4;       // 8-byte columns: 32-byte raw vectors
            if (staged && !env_int("PG_NO_ORD_FAST", 0)) {          // the staged kernel's preconditions are the staged summaries' too
                OrdStagedParams fp{};
                const NCol *rc[6] = {&prm.pred, &prm.key0, has_key1 ? &prm.key1 : nullptr, &prm.A, &prm.B, &prm.C};
                for (int r = 0; r < 6; r++) {
                    fp.rpw[r] = 0; fp.roff[r] = 0;
                    if (!rc[r] || !rc[r]->p) continue;
                    int ci = -1;
                    for (int i = 0; i < fp.st.ncol; i++) if (fp.st.src[i] == (const char *)rc[r]->p) ci = i;
                    if (ci < 0) { ci = fp.st.ncol++; fp.st.src[ci] = (const char *)rc[r]->p; fp.st.pw[ci] = rc[r]->pw; }
                    fp.rpw[r] = -1 - ci;
                }
                stage_layout(&fp.st, LC_TILE);
                for (int r = 0; r < 6; r++) if (fp.rpw[r] < 0) { const int ci = -1 - fp.rpw[r]; fp.rpw[r] = fp.st.pw[ci]; fp.roff[r] = fp.st.off[ci]; }
                // roles the chain does not use were aliased to A by the planner: they must read as zero here
                if (sprm.rpw[0] == 0) fp.rpw[0] = 0;
                if (sprm.rpw[5] == 0) fp.rpw[4] = 0;
                if (sprm.rpw[6] == 0) fp.rpw[5] = 0;
                fp.st.nstage = 6;
                fp.p_lo = sprm.p_lo; fp.p_span = sprm.p_span;
                fp.abase = sprm.abase; fp.bbase = (int)prm.B.base;
                fp.f1c = sprm.f1c; fp.f1s = sprm.f1s; fp.f2c = sprm.f2c; fp.f2s = sprm.f2s;
                fp.has_key1 = has_key1 ? 1 : 0;
                fp.n1 = prm.n1;
                for (size_t i = 0; i < vals[0].size() && i < (size_t)LC_MAXG; i++) fp.code0[i] = vals[0][i];
                for (size_t i = 0; i < vals[1].size() && i < (size_t)LC_MAXG; i++) fp.code1[i] = vals[1][i];
                fp.nrows = prm.nrows;
                typedef void (*OK)(const OrdStagedParams, const OrdJob *, const int *, OrdSummary *, i64, i64);
                OK k = (fp.rpw[0] == 2 && fp.rpw[3] == 4 && fp.rpw[4] == 1 && fp.rpw[5] == 1) ? ord_jobs_staged_kernel<2, 4, 1, 1> : ord_jobs_staged_kernel<-1, -1, -1, -1>;
                const size_t osm = (size_t)ST_HDR + (size_t)fp.st.nstage * fp.st.stage_bytes;
                PG_CUDA(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osm));
                k<<<sms_times((const void *)k, ST_THREADS, osm), ST_THREADS, osm, st>>>(fp, jobs, njobs, d_ord.as<OrdSummary>(), ord_stride, ntiles);
            } else {
#define PG_ORDJ(K, W) ord_jobs_kernel<K, W><<<og, LC_THREADS, 0, st>>>(prm, jobs, njobs, d_ord.as<OrdSummary>(), ord_stride, ntiles)
            if (has_key1) { if (ow) PG_ORDJ(true, true); else PG_ORDJ(true, false); }
            else { if (ow) PG_ORDJ(false, true); else PG_ORDJ(false, false); }
#undef PG_ORDJ
            }
            if (has_key1) ord_fold_kernel<true><<<ORD_MAXJOBS, LC_THREADS, 0, st>>>(prm, jobs, njobs, d_ord.as<OrdSummary>(), ord_stride, ntiles, mine);
            else ord_fold_kernel<false><<<ORD_MAXJOBS, LC_THREADS, 0, st>>>(prm, jobs, njobs, d_ord.as<OrdSummary>(), ord_stride, ntiles, mine);
            PG_CUDA(cudaGetLastError());
            if (nranks() > 1) PG_TRY(comm_allgather(mine, all, contrib_bytes(), st));
            else PG_CUDA(cudaMemcpyAsync(all, mine, contrib_bytes(), cudaMemcpyDeviceToDevice, st));
            PG_CUDA(cudaMemcpyAsync(h_con, all, contrib_bytes() * (size_t)nranks(), cudaMemcpyDeviceToHost, st));
            launches += 3;
        }
        PG_CUDA(cudaMemcpyAsync(h_tot, d_gather.p, rank_bytes() * (size_t)nranks(), cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        PG_CUDA(cudaStreamSynchronize(st));
        tr.mark("kernels+gather+d2h");

        // merge ranks in order; 128-bit exact (merge_rank_partials: also exported host-only as pg_host_merge_partials)
        std::vector<i128> tot((size_t)G * LC_K, 0);
        std::vector<i64> first((size_t)G, INT64_MAX);
        merge_rank_partials(h_tot, rank_bytes(), nranks(), G * LC_K, G, tot.data(), first.data());
        res->stats.kernel_ms = ev_all.ms();
        res->stats.main_kernel_ms = ev_main.ms();
        res->stats.rows_scanned = table->nrows;
        res->stats.algorithmic_bytes = table->nrows * bytes_per_row;
        res->stats.main_kernel_bytes = res->stats.algorithmic_bytes;
        res->stats.kernel_launches = launches;

        // The value the reference's sequential Decimal.Add fold holds for every (group, slot) whose exact total
        // needs 20 digits: the crossing rank's state, then the later ranks' transducers applied in rank order.
        // The job list is enumerated exactly as ord_plan_kernel does (group-major, slots 2..5).
        const i128 THR = (i128)10000000000000000000ULL;
        std::vector<HDec> rounded((size_t)G * LC_K);
        std::vector<char> has_rounded((size_t)G * LC_K, 0);
        emulations = 0;
        {
            int job = 0;
            for (int g = 0; g < G; g++)
                for (int s = 2; s < LC_K; s++) {
                    if (!((emu_mask >> s) & 1u) || !ord_stage) continue;
                    const i128 total = tot[(size_t)g * LC_K + (size_t)s];
                    if (total < THR) continue;
                    if (total >= THR * 9) PG_FAIL(PG_EOVERFLOW, "sum needs more than 20 digits");
                    if (job >= ORD_MAXJOBS) PG_FAIL(PG_EOVERFLOW, "more than %d sums need the ordered rounding emulation", ORD_MAXJOBS);
                    bool have = false;
                    u128 S = 0;
                    for (int r = 0; r < nranks(); r++) {
                        const OrdContrib &k = ((const OrdContrib *)(h_con + contrib_bytes() * (size_t)r))[job];
                        if (k.kind == 9) PG_FAIL(PG_ECUDA, "internal: ordered rounding could not locate the crossing row");
                        if (k.kind == 1) {
                            if (have) PG_FAIL(PG_ECUDA, "internal: two ranks reported a crossing state");
                            S = ((u128)k.s_hi << 64) | k.s_lo;
                            have = true;
                        } else if (k.kind == 2) {
                            if (!have) PG_FAIL(PG_ECUDA, "internal: a transducer precedes the crossing rank");
                            const u128 q = ((u128)k.q_hi << 64) | k.q_lo;
                            S += q + ((S & 1) ? k.c1 : k.c0);
                        }
                    }
                    if (!have) PG_FAIL(PG_ECUDA, "internal: crossing rank did not report a state");
                    if (S > (u128)HD_MAXCOEF) PG_FAIL(PG_EOVERFLOW, "sum needs more than 19 digits after rounding");
                    HDec d;
                    d.coef = (u64)S;
                    d.scale = slot_scale[(size_t)s] - 1;
                    d.neg = false;
                    rounded[(size_t)g * LC_K + (size_t)s] = d;
                    has_rounded[(size_t)g * LC_K + (size_t)s] = 1;
                    job++;
                }
        }
        // exact total -> the Decimal the reference would hold
        auto decimal_sum = [&](int g, int s, HDec *out) -> int {
            const i128 v = tot[(size_t)g * LC_K + (size_t)s];
            if (hd_digits((u128)(v < 0 ? -v : v)) > HD_MAXPREC) {
                if (!has_rounded[(size_t)g * LC_K + (size_t)s])
                    PG_FAIL(PG_EOVERFLOW, "sum needs 20 digits and the ordered rounding emulation does not apply (negative addends, scale 0, or not planned)");
                *out = rounded[(size_t)g * LC_K + (size_t)s];
                emulations++;
                return PG_OK;
            }
            if (!hd_from_i128(v, slot_scale[(size_t)s], out)) PG_FAIL(PG_EOVERFLOW, "decimal overflow");
            return PG_OK;
        };

        // groups in first-insertion order (aggregate_hash.go:424-438)
        std::vector<int> order;
        for (int g = 0; g < G; g++) if (tot[(size_t)g * LC_K] > 0) order.push_back(g);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return first[(size_t)a] < first[(size_t)b]; });
        i64 selected = 0;
        for (int g : order) selected += (i64)tot[(size_t)g * LC_K];
        res->stats.aux[0] = selected;
        res->nrows = (i64)order.size();

        std::vector<char> counted((size_t)G * LC_K, 0);
        for (auto &o : outs) {
            ResCol col;
            if (o.first == 0) {
                const Column &kc = table->cols[(size_t)key_col[o.second]];
                col.type = kc.type;
                if (kc.type == PG_T_DICT8) col.dict = kc.dict;      // results are self-describing (pg_result_column_dict)
                for (int g : order) {
                    int id = (nkeys == 2) ? (o.second == 0 ? g / prm.n1 : g % prm.n1) : g;
                    col.push<uint8_t>(vals[o.second][(size_t)id]);
                }
            } else {
                const AggExpr &a = aggs[(size_t)o.second];
                int s = slot[(size_t)o.second];
                bool is_int = agg_is_int[(size_t)o.second];
                col.width = a.width;
                col.scale = a.scale;
                if (a.fn == PG_AGG_COUNT) col.type = PG_T_HUGEINT;
                else if (a.fn == PG_AGG_SUM) col.type = is_int ? PG_T_HUGEINT : PG_T_DECIMAL128;
                else col.type = is_int ? PG_T_FLOAT64 : PG_T_DECIMAL128;
                for (int g : order) {
                    i128 v = tot[(size_t)g * LC_K + (size_t)s];
                    i128 n = tot[(size_t)g * LC_K];
                    if (a.fn == PG_AGG_COUNT || (a.fn == PG_AGG_SUM && is_int)) {
                        pg_hugeint h;
                        h.lower = (u64)v;
                        h.upper = (i64)(v >> 64);
                        col.push(h);
                    } else if (a.fn == PG_AGG_SUM) {
                        HDec d;
                        PG_TRY(decimal_sum(g, s, &d));
                        col.push(to_pg_decimal(d));
                    } else if (is_int) {   // avg(INT32): float64 sum / float64 count
                        i128 mag = v < 0 ? -v : v;
                        if (mag >= ((i128)1 << 53)) PG_FAIL(PG_EOVERFLOW, "avg(INT32): sum not exact in float64");
                        double x = (double)(i64)v / (double)(i64)n;
                        col.push(x);
                    } else {               // avg(DECIMAL) = sum.Quo(count)
                        HDec sd, nd, qd;
                        PG_TRY(decimal_sum(g, s, &sd));
                        hd_from_i128(n, 0, &nd);
                        if (!hd_quo(sd, nd, &qd)) PG_FAIL(PG_EOVERFLOW, "avg: decimal division failed");
                        col.push(to_pg_decimal(qd));
                    }
                }
            }
            res->cols.push_back(col);
        }
        // distinct (group, slot) sums that took the ordered path
        {
            int n = 0;
            for (size_t i = 0; i < has_rounded.size(); i++) n += has_rounded[i] ? 1 : 0;
            emulations = n;
        }
        res->stats.aux[1] = emulations;
        tr.mark("finalise");
        return PG_OK;
    }
};

static bool prefix_match(const AffProd &a, const AffProd &chain, size_t n)
{
    if (a.f.size() != n || chain.f.size() < n) return false;
    for (size_t i = 0; i < n; i++) if (!(a.f[i] == chain.f[i])) return false;
    return true;
}

// dense ids over the byte codes that occur anywhere in the (sharded) table
static void dense_codes(const Column &col, std::vector<uint8_t> *vals, uint8_t *lut /* [256] */)
{
    const uint32_t *present = col.gpresent();
    for (int code = 0; code < 256; code++)
        if (present[code >> 5] & (1u << (code & 31))) {
            lut[code] = (uint8_t)vals->size();
            vals->push_back((uint8_t)code);
        }
    if (vals->empty()) vals->push_back(0);   // empty table
}

static int try_lowcard(pg_plan *plan, const Node &aggn, const Node &scan, const std::vector<Range> &ranges,
                       const std::vector<AffProd> &args, std::unique_ptr<Pipeline> *out, std::string *why)
{
    const pg_table *t = plan->slots[(size_t)scan.slot];
    if (aggn.groups.empty() || aggn.groups.size() > 2) { *why = "needs 1 or 2 group keys"; return PG_EUNSUPPORTED; }
    if (!aggn.having.empty()) { *why = "HAVING not supported in this shape"; return PG_EUNSUPPORTED; }
    std::unique_ptr<LowcardPipeline> p(new LowcardPipeline());
    p->table = t;
    p->nkeys = (int)aggn.groups.size();
    std::vector<uint8_t> luts(512, 0);
    int dims[2] = {1, 1};
    for (int k = 0; k < p->nkeys; k++) {
        const Expr &ge = aggn.groups[(size_t)k];
        if (ge.kind != PG_TK_COL) { *why = "group key is not a column"; return PG_EUNSUPPORTED; }
        const Column &col = t->cols[(size_t)ge.idx];
        if (!is_byte_family(col.type) || col.any_nulls()) { *why = "group key is not a non-null byte-coded column"; return PG_EUNSUPPORTED; }
        p->key_col[k] = ge.idx;
        dense_codes(col, &p->vals[k], &luts[(size_t)k * 256]);
        dims[k] = (int)p->vals[k].size();
    }
    p->G = dims[0] * dims[1];
    if (p->G > LC_MAXG) { *why = "more dense groups than the low-cardinality kernel holds"; return PG_EUNSUPPORTED; }
    p->has_key1 = p->nkeys == 2;

    // the longest product is the chain A*(c1+s1*B)*(c2+s2*C)
    AffProd chain;
    for (auto &a : args) if (a.f.size() > chain.f.size()) chain = a;
    if (chain.f.size() > 3) { *why = "product of more than three factors"; return PG_EUNSUPPORTED; }
    int colA = -1, colB = -1, colC = -1, colQ = -1;
    if (!chain.f.empty()) {
        if (chain.f[0].c != 0 || chain.f[0].s != 1) { *why = "first factor of the chain is not a plain column"; return PG_EUNSUPPORTED; }
        colA = chain.f[0].col;
        if (chain.f.size() > 1) colB = chain.f[1].col;
        if (chain.f.size() > 2) colC = chain.f[2].col;
    }
    p->aggs = aggn.aggs;
    p->outs = aggn.outs;
    p->slot_scale.assign(LC_K, 0);
    std::vector<char> slot_is_decimal(LC_K, 0);
    for (size_t i = 0; i < aggn.aggs.size(); i++) {
        const AggExpr &a = aggn.aggs[i];
        const AffProd &ap = args[i];
        int s = -1;
        bool is_int = false;
        if (a.fn == PG_AGG_COUNT) { s = 0; is_int = true; }
        else if (a.fn != PG_AGG_SUM && a.fn != PG_AGG_AVG) { *why = "aggregate other than sum/avg/count"; return PG_EUNSUPPORTED; }
        else if (ap.f.size() == 1 && ap.f[0].c == 0 && ap.f[0].s == 1 && t->cols[(size_t)ap.f[0].col].type == PG_T_INT32) {
            if (colQ >= 0 && colQ != ap.f[0].col) { *why = "two different 32-bit aggregate columns"; return PG_EUNSUPPORTED; }
            colQ = ap.f[0].col;
            s = 1;
            is_int = true;
        } else if (prefix_match(ap, chain, 1)) s = 2;
        else if (prefix_match(ap, chain, 2)) s = 3;
        else if (prefix_match(ap, chain, 3)) s = 4;
        else if (ap.f.size() == 1 && ap.f[0].c == 0 && ap.f[0].s == 1 && ap.f[0].col == colB) s = 5;
        else { *why = "aggregate argument does not map onto the chain accumulators"; return PG_EUNSUPPORTED; }
        if (s >= 2) {
            int sc = 0;
            if (s == 5) sc = t->cols[(size_t)colB].type == PG_T_DECIMAL64 ? t->cols[(size_t)colB].scale : 0;
            else for (int k = 0; k < s - 1; k++) sc += chain.f[(size_t)k].scale;
            p->slot_scale[(size_t)s] = sc;
            is_int = a.ltype == PG_LT_HUGEINT || a.ltype == PG_LT_DOUBLE;
            if (is_int && sc != 0) { *why = "integer aggregate over a scaled value"; return PG_EUNSUPPORTED; }
            if (!is_int) slot_is_decimal[(size_t)s] = 1;
        }
        if (a.fn == PG_AGG_SUM && !is_int && a.ltype != PG_LT_DECIMAL) { *why = "sum result type mismatch"; return PG_EUNSUPPORTED; }
        p->slot.push_back(s);
        p->agg_is_int.push_back(is_int);
    }
    for (auto &o : aggn.outs) {
        if (o.first == 0 && (o.second < 0 || o.second >= p->nkeys)) { *why = "bad group output index"; return PG_EUNSUPPORTED; }
        if (o.first == 1 && (o.second < 0 || o.second >= (int)aggn.aggs.size())) { *why = "bad aggregate output index"; return PG_EUNSUPPORTED; }
        if (o.first != 0 && o.first != 1) { *why = "bad output kind"; return PG_EUNSUPPORTED; }
    }
    // predicate: at most one range, on an integer-family column
    if (ranges.size() > 1) { *why = "more than one predicate column"; return PG_EUNSUPPORTED; }
    int colP = -1;
    if (ranges.size() == 1) {
        colP = ranges[0].col;
        if (!is_int_family(t->cols[(size_t)colP].type)) { *why = "predicate column is not an integer-family column"; return PG_EUNSUPPORTED; }
    }
    if (colA < 0) { *why = "no chain column"; return PG_EUNSUPPORTED; }
    for (int cidx : {colA, colB, colC}) if (cidx >= 0 && !is_int_family(t->cols[(size_t)cidx].type)) { *why = "chain column is not an integer-family column"; return PG_EUNSUPPORTED; }
    // every kernel input must exist: alias the missing ones to a column that is read anyway
    LowcardParams &q = p->prm;
    q.nrows = t->nrows;
    q.row_base = t->global_offset;
    const Column &cA = t->cols[(size_t)colA];
    q.A = cA.ncol();
    q.c1 = 1; q.s1 = 0; q.c2 = 1; q.s2 = 0;
    q.B = q.A; q.C = q.A;
    p->bytes_per_row = cA.phys_width() + p->nkeys;
    bool narrow = col_is_narrow(cA);
    if (colB >= 0) { const Column &cB = t->cols[(size_t)colB]; q.B = cB.ncol(); q.c1 = chain.f[1].c; q.s1 = chain.f[1].s; p->bytes_per_row += cB.phys_width(); narrow = narrow && col_is_narrow(cB); }
    if (colC >= 0) { const Column &cC = t->cols[(size_t)colC]; q.C = cC.ncol(); q.c2 = chain.f[2].c; q.s2 = chain.f[2].s; p->bytes_per_row += cC.phys_width(); narrow = narrow && col_is_narrow(cC); }
    if (colQ >= 0) { q.q = t->cols[(size_t)colQ].ncol(); p->bytes_per_row += q.q.pw; narrow = narrow && col_is_narrow(t->cols[(size_t)colQ]); }
    else q.q = q.A;
    i64 slo = 1, shi = 0;
    if (colP >= 0) {
        const Column &cP = t->cols[(size_t)colP];
        q.pred = cP.ncol();
        stored_range(cP, ranges[0].lo, ranges[0].hi, &slo, &shi);
        if (colP != colQ && colP != colA && colP != colB && colP != colC) p->bytes_per_row += cP.phys_width();
        narrow = narrow && cP.phys_width() <= 4;
    } else {                       // no predicate: the chain column with its whole stored domain
        q.pred = q.A;
        stored_range(cA, INT64_MIN, INT64_MAX, &slo, &shi);
    }
    q.lo = slo; q.hi = shi;
    q.key0 = t->cols[(size_t)p->key_col[0]].ncol();
    if (p->has_key1) q.key1 = t->cols[(size_t)p->key_col[1]].ncol();
    else { q.key1 = q.key0; q.key1.p = nullptr; }
    q.n1 = dims[1];
    q.ngroups = p->G;
    // the factors as 32-bit values in the narrow kernel
    auto fac_fits = [&](i64 c, i64 s, const Column &col) {
        return fits_i32(c) && fits_i32(s) && fits_i32((i128)c + (i128)s * col.vmin) && fits_i32((i128)c + (i128)s * col.vmax) && fits_i32((i128)c + (i128)s * col.base);
    };
    if (colB >= 0) narrow = narrow && fac_fits(q.c1, q.s1, t->cols[(size_t)colB]);
    if (colC >= 0) narrow = narrow && fac_fits(q.c2, q.s2, t->cols[(size_t)colC]);
    if (env_int("PG_FORCE_WIDE", 0)) narrow = false;
    p->wide = !narrow;
    p->unroll = env_int("PG_LC_UNROLL", 2) >= 4 ? 4 : 2;
    const i64 ntiles = (t->nrows + LC_TILE - 1) / LC_TILE;

    // exactness proof from the statistics of the WHOLE table (every rank decides alike) and the largest shard
    const i64 max_tiles = (t->max_rows() + LC_TILE - 1) / LC_TILE;
    i128 bound = maxabs(cA.gmin(), cA.gmax());
    if (colB >= 0) {
        const Column &cB = t->cols[(size_t)colB];
        i128 f = std::max(maxabs(q.c1 + q.s1 * cB.gmin(), q.c1 + q.s1 * cB.gmax()), maxabs(cB.gmin(), cB.gmax()));
        bound *= f > 1 ? f : 1;
    }
    if (colC >= 0) {
        const Column &cC = t->cols[(size_t)colC];
        i128 f = maxabs(q.c2 + q.s2 * cC.gmin(), q.c2 + q.s2 * cC.gmax());
        bound *= f > 1 ? f : 1;
    }
    if (colQ >= 0) bound = std::max(bound, maxabs(t->cols[(size_t)colQ].gmin(), t->cols[(size_t)colQ].gmax()));
    // can any DECIMAL total need 20 digits?  Then the CTAs take contiguous tile runs so that their partials are
    // ordered, and the ordered-rounding stage is appended (on every rank: the bound is global)
    i128 table_bound = bound * (i128)std::max<i64>(t->total_rows(), 1);
    q.contig = table_bound >= (i128)1000000000000000000LL ? 1 : 0;   // 10^18: generous margin
    if (getenv("PG_LOWCARD_CONTIG")) q.contig = env_int("PG_LOWCARD_CONTIG", 0) ? 1 : 0;
    {
        const bool a_pos = cA.gmin() >= 0;
        const bool f1_pos = colB < 0 || (q.c1 + q.s1 * t->cols[(size_t)colB].gmin() >= 0 && q.c1 + q.s1 * t->cols[(size_t)colB].gmax() >= 0);
        const bool f2_pos = colC < 0 || (q.c2 + q.s2 * t->cols[(size_t)colC].gmin() >= 0 && q.c2 + q.s2 * t->cols[(size_t)colC].gmax() >= 0);
        const bool nonneg[LC_K] = {true, true, a_pos, a_pos && f1_pos, a_pos && f1_pos && f2_pos, colB < 0 || t->cols[(size_t)colB].gmin() >= 0};
        for (int s = 2; s < LC_K; s++)
            if (slot_is_decimal[(size_t)s] && nonneg[s] && p->slot_scale[(size_t)s] >= 1) p->emu_mask |= 1u << s;
    }
    p->ord_stage = q.contig && p->emu_mask != 0;

    // Narrow columns over non-negative values: the bulk-copy staged kernel with packed three-word table entries
    // (scan_staged.cuh) whenever the statistics prove that no packed field can overflow.
    if (narrow && !env_int("PG_NO_STAGED", 0) && t->nrows > 0) {
        LowcardSParams &sp = p->sprm;
        sp = LowcardSParams{};
        std::string no;
        const Column *cP = colP >= 0 ? &t->cols[(size_t)colP] : nullptr;
        const Column *cQ = colQ >= 0 ? &t->cols[(size_t)colQ] : nullptr;
        const Column *cB = colB >= 0 ? &t->cols[(size_t)colB] : nullptr;
        const Column *cC = colC >= 0 ? &t->cols[(size_t)colC] : nullptr;
        auto st_min = [](const Column *c) { return c ? (i128)c->vmin - c->base : (i128)0; };
        auto st_max = [](const Column *c) { return c ? (i128)c->vmax - c->base : (i128)0; };
        const i128 f1lo = cB ? std::min((i128)q.c1 + (i128)q.s1 * cB->vmin, (i128)q.c1 + (i128)q.s1 * cB->vmax) : 1;
        const i128 f1hi = cB ? std::max((i128)q.c1 + (i128)q.s1 * cB->vmin, (i128)q.c1 + (i128)q.s1 * cB->vmax) : 1;
        const i128 f2lo = cC ? std::min((i128)q.c2 + (i128)q.s2 * cC->vmin, (i128)q.c2 + (i128)q.s2 * cC->vmax) : 1;
        const i128 f2hi = cC ? std::max((i128)q.c2 + (i128)q.s2 * cC->vmin, (i128)q.c2 + (i128)q.s2 * cC->vmax) : 1;
        if (st_min(cQ) < 0 || st_min(&cA) < 0 || st_min(cB) < 0 || st_min(cC) < 0 || cA.vmin < 0 || f1lo < 0 || f2lo < 0) no = "negative values";
        const i128 t2max = (i128)std::max<i64>(cA.vmax, 0) * f1hi;
        if (no.empty() && (t2max >= ((i128)1 << 32) || f2hi >= ((i128)1 << 31))) no = "A*(c1+s1*B) does not fit 32 bits";
        unsigned plo = 0, pspan = 0xffffffffu;
        if (no.empty() && cP && !unsigned_range(q.lo, q.hi, &plo, &pspan)) no = "empty predicate range";
        // the slot hash: injective on every key combination that can occur
        unsigned M = 0;
        if (no.empty()) {
            u64 x = 0x9e3779b97f4a7c15ULL;
            bool found = false;
            for (int tries = 0; tries < 200000 && !found; tries++) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                const unsigned cand = (unsigned)(x >> 16) | 1u;
                unsigned used = 0;
                bool inj = true;
                for (size_t i0 = 0; i0 < p->vals[0].size() && inj; i0++)
                    for (size_t i1 = 0; i1 < (p->has_key1 ? p->vals[1].size() : (size_t)1) && inj; i1++) {
                        const unsigned sl = lc_slot_of(p->vals[0][i0], p->has_key1 ? p->vals[1][i1] : 0, p->has_key1, cand);
                        if (used & (1u << sl)) inj = false;
                        used |= 1u << sl;
                    }
                if (inj) { M = cand; found = true; }
            }
            if (!found) no = "no injective slot hash";
        }
        if (no.empty()) {
            for (int r = 0; r < 7; r++) { sp.rpw[r] = 0; sp.roff[r] = 0; }
            const Column *roles[7] = {cP, &t->cols[(size_t)p->key_col[0]], p->has_key1 ? &t->cols[(size_t)p->key_col[1]] : nullptr, cQ, &cA, cB, cC};
            for (int r = 0; r < 7; r++) {
                if (!roles[r]) continue;
                const int ci = stage_add(&sp.st, *roles[r]);
                if (ci < 0) { no = "too many columns"; break; }
                sp.rpw[r] = -1 - ci;
            }
        }
        if (no.empty()) {
            p->qpt = env_int("PG_QPT", 2) >= 2 ? 2 : 1;
            const i64 tile_rows = (i64)ST_CONS_WARPS * 128 * p->qpt;
            stage_layout(&sp.st, (int)tile_rows);
            for (int r = 0; r < 7; r++) if (sp.rpw[r] < 0) { const int ci = -1 - sp.rpw[r]; sp.rpw[r] = sp.st.pw[ci]; sp.roff[r] = sp.st.off[ci]; }
            // as many stages as let two CTAs share an SM
            const size_t budget = (size_t)(ctx().prop.sharedMemPerMultiprocessor / 2) - 1024 - 1024;
            int ns = env_int("PG_NSTAGE", 0);
            if (ns <= 0) ns = (int)std::min<size_t>(ST_MAXSTAGE, (budget - ST_HDR - LCS_TBL_BYTES) / (size_t)sp.st.stage_bytes);
            if (ns < 2) ns = 2;
            sp.st.nstage = ns;
            const size_t smem = (size_t)ST_HDR + (size_t)ns * sp.st.stage_bytes + LCS_TBL_BYTES;
            int wsig[7];
            for (int r = 0; r < 7; r++) wsig[r] = sp.rpw[r];
            if (env_int("PG_NO_SPECIALISED", 0)) wsig[0] = 8;
            p->skern = lowcard_staged_variant(wsig, p->has_key1, p->qpt, &p->specialised);
            const void *kern = (const void *)p->skern;
            if (smem > (size_t)ctx().prop.sharedMemPerBlockOptin) no = "stages do not fit in shared memory";
            else {
                PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const i64 ntiles_s = (t->nrows + tile_rows - 1) / tile_rows;
                const int full = sms_times(kern, ST_THREADS, smem);
                const int grid = (int)std::max<i64>(1, std::min<i64>(full, ntiles_s));
                const i64 per_s = (ntiles_s + grid - 1) / grid;
                // a consumer thread sees per_s tiles x qpt quads x 4 rows: field widths of the packed entry
                const i128 rows_thr = (i128)per_s * p->qpt * 4;
                const int b_n = bit_length(rows_thr), b_q = bit_length(rows_thr * st_max(cQ)), b_B = bit_length(rows_thr * st_max(cB));
                const int b_A = bit_length(rows_thr * st_max(&cA)), b_t2 = bit_length(rows_thr * t2max), b_t3 = bit_length(rows_thr * t2max * f2hi);
                const int sh0 = std::max(32, b_t3), sh1 = std::max(32, b_t2), sh2 = std::max(32, b_A);
                if (b_n > 64 - sh0 || b_q > 64 - sh1 || b_B > 64 - sh2) no = "packed table fields could overflow";
                else if ((i128)bound * per_s * tile_rows >= ((i128)1 << 62)) no = "per-CTA partial sum could exceed int64";
                else {
                    sp.p_lo = plo; sp.p_span = pspan;
                    sp.abase = (int)cA.base;
                    sp.f1c = cB ? (int)(q.c1 + q.s1 * cB->base) : 1; sp.f1s = cB ? (int)q.s1 : 0;
                    sp.f2c = cC ? (int)(q.c2 + q.s2 * cC->base) : 1; sp.f2s = cC ? (int)q.s2 : 0;
                    sp.hashM = M;
                    sp.sh0 = sh0; sp.sh1 = sh1; sp.sh2 = sh2;
                    sp.mul0 = 1u << (sh0 - 32); sp.mul1 = 1u << (sh1 - 32); sp.mul2 = 1u << (sh2 - 32);
                    memset(sp.slot_group, 0xff, 8);
                    for (size_t i0 = 0; i0 < p->vals[0].size(); i0++)
                        for (size_t i1 = 0; i1 < (p->has_key1 ? p->vals[1].size() : (size_t)1); i1++)
                            sp.slot_group[lc_slot_of(p->vals[0][i0], p->has_key1 ? p->vals[1][i1] : 0, p->has_key1, M)] = (unsigned char)(i0 * (size_t)dims[1] + i1);
                    sp.ngroups = p->G;
                    sp.contig = q.contig;
                    sp.qbase = cQ ? cQ->base : 0; sp.Abase = cA.base; sp.Bbase = cB ? cB->base : 0;
                    sp.nrows = t->nrows;
                    p->staged = true;
                    p->smem = smem;
                    p->grid = grid;
                    p->lc_per = per_s * (tile_rows / LC_TILE);
                }
            }
        }
        if (!p->staged && getenv("PG_TRACE")) fprintf(stderr, "[pg] lowcard: staged kernel not used: %s\n", no.c_str());
    }
    if (!p->staged) {
    // kernel variant, shared memory and grid.  ACC32: the three small accumulators as 32-bit table slots when a
    // thread's total over its stored values provably fits (rows per thread from the occupancy of the 64-bit variant,
    // which is never larger than the 32-bit variant's).
    auto prep = [&](LowcardKernel k, bool a32) -> int {      // opt in to the shared memory the variant needs; full-GPU grid
        const size_t sm = (size_t)lc_smem_per_thread(p->G, a32) * LC_THREADS;
        cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        return sms_times((const void *)k, LC_THREADS, sm);
    };
    p->acc32 = false;
    if (narrow && !env_int("PG_LC_NO_ACC32", 0)) {
        const int g64 = (int)std::max<i64>(1, std::min<i64>(prep(lowcard_variant(false, false, p->has_key1, p->unroll), false), ntiles));
        const i128 rows_thr = (i128)((ntiles + g64 - 1) / g64 + p->unroll) * SA_VEC;
        const Column &cq = t->cols[(size_t)(colQ >= 0 ? colQ : colA)], &cb = t->cols[(size_t)(colB >= 0 ? colB : colA)];
        const i128 m = std::max<i128>(1, std::max((i128)cq.vmax - cq.base, (i128)cb.vmax - cb.base));
        const bool nonneg_stored = (i128)cq.vmin - cq.base >= 0 && (i128)cb.vmin - cb.base >= 0;
        p->acc32 = nonneg_stored && rows_thr * m < ((i128)1 << 32);
    }
    p->smem = (size_t)lc_smem_per_thread(p->G, p->acc32) * LC_THREADS;
    const int full = prep(lowcard_variant(p->wide, p->acc32, p->has_key1, p->unroll), p->acc32);
    PG_CUDA(cudaGetLastError());
    p->grid = (int)std::max<i64>(1, std::min<i64>(full, ntiles));
    {
        const int full64 = prep(lowcard_variant(true, false, p->has_key1, 2), false);
        const i64 gmin_grid = std::max<i64>(1, std::min<i64>(full64, max_tiles));     // the smallest grid any rank could run
        i128 rows_per_cta = (i128)((max_tiles + gmin_grid - 1) / gmin_grid) * LC_TILE;
        if (bound * rows_per_cta >= ((i128)1 << 62)) { *why = "per-CTA partial sum could exceed int64"; return PG_EUNSUPPORTED; }
    }
    p->lc_per = (ntiles + p->grid - 1) / p->grid;
    }
    PG_TRY(p->d_part.alloc(sizeof(i64) * (size_t)p->grid * (size_t)p->G * LC_K));
    PG_TRY(p->d_final.alloc(p->rank_bytes()));
    PG_TRY(p->d_gather.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->h_final.alloc(p->host_bytes()));
    PG_TRY(p->d_luts.alloc(512));
    if (p->ord_stage) {
        const i64 per = p->lc_per;
        p->ord_stride = per + ntiles / ORD_CHUNK + 4;
        PG_TRY(p->d_jobs.alloc(sizeof(OrdJob) * ORD_MAXJOBS + 64));
        PG_TRY(p->d_ord.alloc(sizeof(OrdSummary) * (size_t)p->ord_stride * ORD_MAXJOBS));
        PG_TRY(p->d_contrib.alloc(p->contrib_bytes() * (size_t)(p->nranks() + 1)));
    }
    PG_CUDA(cudaMemcpyAsync(p->d_luts.p, luts.data(), 512, cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaStreamSynchronize(ctx().stream));
    q.luts = p->d_luts.as<uint8_t>();
    char buf[640];
    snprintf(buf, sizeof buf,
             "ScanAgg[lowcard-chain] table=%s rows=%lld kernel=%s<%s,%s,%d,%d> grid=%d block=%d smem=%zu "
             "groups=%dx%d stored bytes/row=%d (widths: pred=%d q=%d A=%d B=%d C=%d) stored pred=[%lld,%lld] chain: A*(%lld%+lld*B)*(%lld%+lld*C) tiles=%s%s",
             t->name.c_str(), (long long)t->nrows, p->staged ? "lowcard_staged_kernel" : "lowcard_chain_kernel", p->staged ? (p->specialised ? "bulk-copy ring, widths compiled in" : "bulk-copy ring, run-time widths") : p->wide ? "wide" : "narrow",
             p->staged ? "packed3" : p->acc32 ? "acc32" : "acc64", (int)p->has_key1, p->staged ? p->qpt : p->unroll, p->grid,
             p->staged ? ST_THREADS : LC_THREADS, p->smem, dims[0], dims[1], p->bytes_per_row, q.pred.pw, q.q.pw, q.A.pw, q.B.pw, q.C.pw, (long long)slo, (long long)shi,
             (long long)q.c1, (long long)q.s1, (long long)q.c2, (long long)q.s2,
             q.contig ? "contiguous-per-CTA(ordered partials)" : "interleaved", p->ord_stage ? " +ordered-rounding stage (device)" : "");
    p->explain = buf;
    *out = std::move(p);
    return PG_OK;
}

// ------------------------------------------------------------------ generic --

struct GenericPipeline : Pipeline {
    const pg_table *table = nullptr;
    GenParams prm{};
    int NT = 256, grid = 1, G = 1, P = 1;
    bool nulls = false;                  // some referenced column holds NULLs: per-aggregate valid counts are kept
    size_t smem = 0;
    int nkeys = 0, key_col[2] = {-1, -1};
    std::vector<uint8_t> vals[2];
    std::vector<AggExpr> aggs;
    std::vector<int> plane;              // accumulator plane per aggregate (0 = row count)
    std::vector<int> plane_scale;        // value scale per plane
    std::vector<int> plane_kind;         // GEN_* per plane
    std::vector<bool> agg_is_int;
    std::vector<std::pair<int, int>> outs;
    i64 bytes_per_row = 0, extra_bytes = 0;      // extra: string payload + offsets of LIKE'd VARCHAR columns
    DevBuf d_part, d_final, d_luts, d_gather, d_kinds;
    PinBuf h_final;
    EventPair ev_all, ev_main;
    // expression-driven mode (scanagg_vm.cuh): predicate and aggregate arguments are row programs
    bool vm = false;
    RvCompiler cc;
    VmAggParams vprm{};
    DevBuf d_code, d_err;

    int nranks() const { return table->dist == PG_DIST_REPLICATED ? 1 : ctx().world; }
    size_t rank_bytes() const { return (size_t)G * P * 16 + 64 * 8; }

    int run(pg_result *res) override
    {
        Context &c = ctx();
        cudaStream_t st = c.stream;
        PG_TRY(ev_all.init());
        PG_TRY(ev_main.init());
        i64 *d_firstrow = (i64 *)((char *)d_final.p + (size_t)G * P * 16);
        PG_CUDA(cudaEventRecord(ev_all.a, st));
        PG_CUDA(cudaMemsetAsync(d_firstrow, 0x7f, 64 * 8, st));
        PG_CUDA(cudaEventRecord(ev_main.a, st));
        if (vm) PG_CUDA(cudaMemsetAsync(d_err.p, 0, 4, st));
#define PG_GEN(N) do { if (vm) vm_scanagg_kernel<N><<<grid, N, smem, st>>>(vprm, d_part.as<i64>(), d_firstrow); \
                       else if (nulls) generic_scanagg_kernel<N, true><<<grid, N, smem, st>>>(prm, d_part.as<i64>(), d_firstrow); \
                       else generic_scanagg_kernel<N, false><<<grid, N, smem, st>>>(prm, d_part.as<i64>(), d_firstrow); } while (0)
        if (NT == 256) PG_GEN(256);
        else if (NT == 128) PG_GEN(128);
        else PG_GEN(64);
#undef PG_GEN
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaEventRecord(ev_main.b, st));
        finalize_generic_kernel<<<(G * P + 63) / 64, 64, 0, st>>>(d_part.as<i64>(), grid, G * P, d_kinds.as<int>(), d_final.as<u64>());
        PG_CUDA(cudaGetLastError());
        const void *src = d_final.p;
        if (nranks() > 1) {
            PG_TRY(comm_allgather(d_final.p, d_gather.p, rank_bytes(), st));
            src = d_gather.p;
        }
        PG_CUDA(cudaMemcpyAsync(h_final.p, src, rank_bytes() * (size_t)nranks(), cudaMemcpyDeviceToHost, st));
        int vm_err = 0;
        if (vm) PG_CUDA(cudaMemcpyAsync(&vm_err, d_err.p, 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        PG_CUDA(cudaStreamSynchronize(st));
        // (a rank that fails here leaves its peers' collectives matched: the all-gather above has already run)
        if (vm_err == RV_ERR_DIVZERO) PG_FAIL(PG_EOVERFLOW, "aggregate expression: division by zero");
        if (vm_err == RV_ERR_FLOAT) PG_FAIL(PG_EUNSUPPORTED, "aggregate expression: value outside the exactly reproducible float32 cast range");
        if (vm_err) PG_FAIL(PG_EOVERFLOW, "aggregate expression: a value exceeds the exact accumulation range");
        std::vector<i128> tot((size_t)G * P);
        std::vector<i64> first((size_t)G, INT64_MAX);
        for (int v = 0; v < G * P; v++) {
            int kind = plane_kind[(size_t)(v % P)];
            tot[(size_t)v] = kind == GEN_SUM ? 0 : kind == GEN_MIN ? (i128)INT64_MAX : (i128)INT64_MIN;
        }
        for (int r = 0; r < nranks(); r++) {
            const char *base = (const char *)h_final.p + rank_bytes() * (size_t)r;
            const u64 *h = (const u64 *)base;
            const i64 *f = (const i64 *)(base + (size_t)G * P * 16);
            for (int v = 0; v < G * P; v++) {
                i128 x = make_i128(h[2 * v], h[2 * v + 1]);
                int kind = plane_kind[(size_t)(v % P)];
                if (kind == GEN_SUM) tot[(size_t)v] += x;
                else if (kind == GEN_MIN) tot[(size_t)v] = std::min(tot[(size_t)v], x);
                else tot[(size_t)v] = std::max(tot[(size_t)v], x);
            }
            for (int g = 0; g < G; g++) if (f[g] != 0x7f7f7f7f7f7f7f7fLL && f[g] < first[(size_t)g]) first[(size_t)g] = f[g];
        }
        res->stats.kernel_ms = ev_all.ms();
        res->stats.main_kernel_ms = ev_main.ms();
        res->stats.rows_scanned = table->nrows;
        res->stats.algorithmic_bytes = table->nrows * bytes_per_row + extra_bytes;
        res->stats.main_kernel_bytes = res->stats.algorithmic_bytes;
        res->stats.kernel_launches = 2;
        std::vector<int> order;
        for (int g = 0; g < G; g++) if (tot[(size_t)g * P] > 0) order.push_back(g);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return first[(size_t)a] < first[(size_t)b]; });
        i64 selected = 0;
        for (int g : order) selected += (i64)tot[(size_t)g * P];
        res->stats.aux[0] = selected;
        res->nrows = (i64)order.size();
        for (auto &o : outs) {
            ResCol col;
            if (o.first == 0) {
                const Column &kc = table->cols[(size_t)key_col[o.second]];
                col.type = kc.type;
                if (kc.type == PG_T_DICT8) col.dict = kc.dict;      // results are self-describing (pg_result_column_dict)
                for (int g : order) {
                    int id = (nkeys == 2) ? (o.second == 0 ? g / prm.n1 : g % prm.n1) : g;
                    col.push<uint8_t>(vals[o.second][(size_t)id]);
                }
            } else {
                const AggExpr &a = aggs[(size_t)o.second];
                int pl = plane[(size_t)o.second];
                bool is_int = agg_is_int[(size_t)o.second];
                col.width = a.width;
                col.scale = a.scale;
                if (a.fn == PG_AGG_COUNT) col.type = PG_T_HUGEINT;
                else if (a.fn == PG_AGG_AVG) col.type = is_int ? PG_T_FLOAT64 : PG_T_DECIMAL128;
                else col.type = is_int ? PG_T_HUGEINT : PG_T_DECIMAL128;
                size_t nrow = 0;
                for (int g : order) {
                    i128 v = tot[(size_t)g * P + (size_t)pl], n = tot[(size_t)g * P];
                    if (nulls && pl > 0) {
                        // NULL inputs were skipped: the count that matters is the aggregate's own; no valid
                        // input at all => NULL (SumOp/AvgOp/CountOp/MinMaxOp.Finalize, function_aggr.go:815-1032)
                        n = tot[(size_t)g * P + (size_t)(prm.nacc + pl)];
                        if (n == 0) { col.push_null(nrow++, (size_t)type_size(col.type)); continue; }
                    }
                    col.mark_valid();
                    nrow++;
                    if (a.fn == PG_AGG_COUNT || (a.fn != PG_AGG_AVG && is_int)) {
                        pg_hugeint h;
                        h.lower = (u64)v;
                        h.upper = (i64)(v >> 64);
                        col.push(h);
                    } else if (a.fn == PG_AGG_AVG && is_int) {
                        i128 mag = v < 0 ? -v : v;
                        if (mag >= ((i128)1 << 53)) PG_FAIL(PG_EOVERFLOW, "avg(INT): sum not exact in float64");
                        col.push((double)(i64)v / (double)(i64)n);
                    } else {
                        HDec d;
                        if (hd_digits((u128)(v < 0 ? -v : v)) > HD_MAXPREC || !hd_from_i128(v, plane_scale[(size_t)pl], &d))
                            PG_FAIL(PG_EOVERFLOW, "decimal aggregate exceeds 19 significant digits (order-dependent rounding regime)");
                        if (a.fn == PG_AGG_AVG) {
                            HDec nd, qd;
                            hd_from_i128(n, 0, &nd);
                            if (!hd_quo(d, nd, &qd)) PG_FAIL(PG_EOVERFLOW, "avg: decimal division failed");
                            d = qd;
                        }
                        col.push(to_pg_decimal(d));
                    }
                }
            }
            res->cols.push_back(col);
        }
        return PG_OK;
    }
};

static int try_generic(pg_plan *plan, const Node &aggn, const Node &scan, const std::vector<Range> &ranges,
                       const std::vector<AffProd> &args, std::unique_ptr<Pipeline> *out, std::string *why)
{
    const pg_table *t = plan->slots[(size_t)scan.slot];
    if (aggn.groups.size() > 2) { *why = "more than two group keys"; return PG_EUNSUPPORTED; }
    if (!aggn.having.empty()) { *why = "HAVING"; return PG_EUNSUPPORTED; }
    if (ranges.size() > GEN_MAXPRED) { *why = "too many predicate columns"; return PG_EUNSUPPORTED; }
    std::unique_ptr<GenericPipeline> p(new GenericPipeline());
    p->table = t;
    p->nkeys = (int)aggn.groups.size();
    std::vector<uint8_t> luts(512, 0);
    int dims[2] = {1, 1};
    std::vector<std::pair<int, int>> used;   // (column, width) for the byte accounting
    auto use = [&](int col) { for (auto &u : used) if (u.first == col) return; used.push_back({col, t->cols[(size_t)col].phys_width()}); };
    for (int k = 0; k < p->nkeys; k++) {
        const Expr &ge = aggn.groups[(size_t)k];
        if (ge.kind != PG_TK_COL) { *why = "group key is not a column"; return PG_EUNSUPPORTED; }
        const Column &col = t->cols[(size_t)ge.idx];
        if (!is_byte_family(col.type) || col.any_nulls()) { *why = "group key is not a non-null byte-coded column"; return PG_EUNSUPPORTED; }
        p->key_col[k] = ge.idx;
        use(ge.idx);
        dense_codes(col, &p->vals[k], &luts[(size_t)k * 256]);     // over the codes of the whole (sharded) table
        dims[k] = (int)p->vals[k].size();
    }
    p->G = dims[0] * dims[1];
    if (p->G > 64) { *why = "more than 64 dense groups"; return PG_EUNSUPPORTED; }
    GenParams &q = p->prm;
    q.nrows = t->nrows;
    q.row_base = t->global_offset;
    // string predicates first out of the range list
    std::vector<Range> likes, plain;
    for (auto &r : ranges) (r.like ? likes : plain).push_back(r);
    if (likes.size() > GEN_MAXLIKE) { *why = "more than 2 string predicates"; return PG_EUNSUPPORTED; }
    q.nlike = (int)likes.size();
    for (size_t i = 0; i < likes.size(); i++) {
        const Column &col = t->cols[(size_t)likes[i].col];
        if (likes[i].pat.size() > GEN_PATMAX) { *why = "string pattern longer than 48 bytes"; return PG_EUNSUPPORTED; }
        if (!col.d_off) { *why = "VARCHAR column has no device copy"; return PG_EUNSUPPORTED; }
        q.like[i].bytes = col.d_bytes;
        q.like[i].off = (const i64 *)col.d_off;
        q.like[i].kind = likes[i].like;
        q.like[i].plen = (int)likes[i].pat.size();
        memcpy(q.like[i].pat, likes[i].pat.data(), likes[i].pat.size());
        std::string lit;
        if (likes[i].like <= 2 && like_is_contains(likes[i].pat, &lit)) {      // '%lit%': word-at-a-time search
            q.like[i].kind = likes[i].like == 1 ? 5 : 6;
            q.like[i].plen = (int)lit.size();
            memcpy(q.like[i].pat, lit.data(), lit.size());
        }
        p->extra_bytes += (i64)col.h_bytes.size() + 8 * t->nrows;
    }
    const std::vector<Range> &ranges_ = plain;
    q.npred = (int)ranges_.size();
    for (size_t i = 0; i < ranges_.size(); i++) {
        const Range *rp = &ranges_[i];
        const std::vector<Range> &ranges = ranges_;
        (void)rp;
        const Column &col = t->cols[(size_t)ranges[i].col];
        q.pcol[i].c = col.ncol();
        q.pcol[i].valid = col.has_nulls ? col.d_valid : nullptr;
        if (col.any_nulls()) p->nulls = true;           // agreed across ranks: the plane count shapes the merge
        stored_range(col, ranges[i].lo, ranges[i].hi, &q.plo[i], &q.phi[i]);
        q.pset[i] = ranges[i].is_set ? 1 : 0;
        memcpy(q.pmask[i], ranges[i].set, sizeof q.pmask[i]);
        use(ranges[i].col);
    }
    q.nkeys = p->nkeys;
    q.key0 = p->nkeys > 0 ? (const uint8_t *)t->cols[(size_t)p->key_col[0]].d_data : nullptr;
    q.key1 = p->nkeys > 1 ? (const uint8_t *)t->cols[(size_t)p->key_col[1]].d_data : nullptr;
    q.n1 = dims[1];
    q.ngroups = p->G;
    // planes: 0 = row count, then one per distinct (kind, product)
    p->plane_kind = {GEN_SUM};
    p->plane_scale = {0};
    std::vector<AffProd> plane_prod(1);
    i128 worst = 1;
    p->aggs = aggn.aggs;
    for (size_t i = 0; i < aggn.aggs.size(); i++) {
        const AggExpr &a = aggn.aggs[i];
        if (a.fn == PG_AGG_COUNT && args[i].f.empty()) { p->plane.push_back(0); p->agg_is_int.push_back(true); continue; }
        int kind = a.fn == PG_AGG_COUNT ? GEN_COUNTV : a.fn == PG_AGG_MIN ? GEN_MIN : a.fn == PG_AGG_MAX ? GEN_MAX : GEN_SUM;
        const AffProd &ap = args[i];
        if (ap.f.empty() || ap.f.size() > GEN_MAXFAC) { *why = "aggregate argument has an unsupported number of factors"; return PG_EUNSUPPORTED; }
        int found = -1;
        for (size_t pl = 1; pl < plane_prod.size(); pl++)
            if (p->plane_kind[pl] == kind && plane_prod[pl].f == ap.f) found = (int)pl;
        if (found < 0) {
            if ((int)plane_prod.size() > GEN_MAXACC) { *why = "too many distinct aggregate arguments"; return PG_EUNSUPPORTED; }
            GenAcc &A = q.acc[plane_prod.size() - 1];
            A.kind = kind;
            A.nfac = (int)ap.f.size();
            i128 bound = 1;
            for (size_t f = 0; f < ap.f.size(); f++) {
                const Column &col = t->cols[(size_t)ap.f[f].col];
                A.fac[f].c = col.ncol();
                A.fac[f].valid = col.has_nulls ? col.d_valid : nullptr;
                if (col.any_nulls()) p->nulls = true;
                A.c[f] = ap.f[f].c + ap.f[f].s * col.base;      // the kernel multiplies STORED values
                A.s[f] = ap.f[f].s;
                use(ap.f[f].col);
                i128 m = std::max(maxabs(ap.f[f].c + ap.f[f].s * col.gmin(), ap.f[f].c + ap.f[f].s * col.gmax()), (i128)1);
                bound *= m;
            }
            if (kind == GEN_SUM) worst = std::max(worst, bound);
            else if (kind != GEN_COUNTV && bound >= ((i128)1 << 62)) { *why = "min/max argument could exceed int64"; return PG_EUNSUPPORTED; }
            found = (int)plane_prod.size();
            plane_prod.push_back(ap);
            p->plane_kind.push_back(kind);
            p->plane_scale.push_back(ap.vscale());
        }
        p->plane.push_back(found);
        bool is_int = a.ltype == PG_LT_HUGEINT || a.ltype == PG_LT_DOUBLE || a.ltype == PG_LT_INTEGER || a.ltype == PG_LT_BIGINT;
        if (is_int && ap.vscale() != 0 && kind != GEN_COUNTV) { *why = "integer aggregate over a scaled value"; return PG_EUNSUPPORTED; }
        if (!is_int && a.ltype != PG_LT_DECIMAL) { *why = "aggregate result type"; return PG_EUNSUPPORTED; }
        if ((kind == GEN_MIN || kind == GEN_MAX) && is_int) { *why = "min/max are DECIMAL only in the reference"; return PG_EUNSUPPORTED; }
        p->agg_is_int.push_back(is_int);
    }
    q.nacc = (int)plane_prod.size() - 1;
    p->P = p->nulls ? 1 + 2 * q.nacc : 1 + q.nacc;
    for (int a = 0; a < q.nacc; a++) if (p->plane_kind[(size_t)(a + 1)] == GEN_COUNTV) p->plane_kind[(size_t)(a + 1)] = GEN_SUM;
    if (p->nulls) for (int a = 0; a < q.nacc; a++) { p->plane_kind.push_back(GEN_SUM); p->plane_scale.push_back(0); }
    for (auto &o : aggn.outs) {
        if (o.first == 0 && (o.second < 0 || o.second >= p->nkeys)) { *why = "bad group output index"; return PG_EUNSUPPORTED; }
        if (o.first == 1 && (o.second < 0 || o.second >= (int)aggn.aggs.size())) { *why = "bad aggregate output index"; return PG_EUNSUPPORTED; }
        if (o.first != 0 && o.first != 1) { *why = "bad output kind"; return PG_EUNSUPPORTED; }
    }
    p->outs = aggn.outs;
    for (auto &u : used) p->bytes_per_row += u.second;
    // thread-private tables must fit in shared memory
    p->NT = 256;
    while (p->NT >= 64 && (size_t)p->G * p->P * p->NT * 8 > (size_t)200 * 1024) p->NT /= 2;
    if (p->NT < 64) { *why = "group tables do not fit in shared memory"; return PG_EUNSUPPORTED; }
    p->smem = (size_t)p->G * p->P * p->NT * 8;
    const void *kern = p->nulls ? (p->NT == 256 ? (const void *)generic_scanagg_kernel<256, true> : p->NT == 128 ? (const void *)generic_scanagg_kernel<128, true> : (const void *)generic_scanagg_kernel<64, true>)
                                : (p->NT == 256 ? (const void *)generic_scanagg_kernel<256, false> : p->NT == 128 ? (const void *)generic_scanagg_kernel<128, false> : (const void *)generic_scanagg_kernel<64, false>);
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    const i64 g = sms_times(kern, p->NT, p->smem);
    const i64 tile = (i64)p->NT * SA_VEC;
    p->grid = (int)std::max<i64>(std::min(g, (t->nrows + tile - 1) / tile), 1);
    const i64 max_tiles = (t->max_rows() + tile - 1) / tile, gmin_grid = std::max<i64>(1, std::min<i64>(g, max_tiles));
    i128 rows_per_cta = (i128)((max_tiles + gmin_grid - 1) / gmin_grid) * tile;                 // largest shard, full grid
    if (worst * rows_per_cta >= ((i128)1 << 62)) { *why = "per-CTA partial sum could exceed int64"; return PG_EUNSUPPORTED; }
    PG_TRY(p->d_part.alloc(sizeof(i64) * (size_t)p->grid * (size_t)p->G * (size_t)p->P));
    PG_TRY(p->d_final.alloc(p->rank_bytes()));
    PG_TRY(p->d_gather.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->h_final.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->d_luts.alloc(512));
    PG_TRY(p->d_kinds.alloc(sizeof(int) * (size_t)p->G * (size_t)p->P));
    std::vector<int> kinds((size_t)p->G * (size_t)p->P);
    for (int v = 0; v < p->G * p->P; v++) kinds[(size_t)v] = p->plane_kind[(size_t)(v % p->P)];
    PG_CUDA(cudaMemcpyAsync(p->d_luts.p, luts.data(), 512, cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaMemcpyAsync(p->d_kinds.p, kinds.data(), sizeof(int) * kinds.size(), cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaStreamSynchronize(ctx().stream));
    q.luts = p->d_luts.as<uint8_t>();
    char buf[384];
    snprintf(buf, sizeof buf,
             "ScanAgg[generic] table=%s rows=%lld kernel=generic_scanagg_kernel<%d> grid=%d smem=%zu groups=%dx%d "
             "predicates=%d accumulators=%d stored bytes/row=%lld%s",
             t->name.c_str(), (long long)t->nrows, p->NT, p->grid, p->smem, dims[0], dims[1], q.npred, q.nacc, (long long)p->bytes_per_row,
             p->nulls ? " nulls=validity-bitmaps" : "");
    p->explain = buf;
    *out = std::move(p);
    return PG_OK;
}

// Expression-driven scan aggregate (scanagg_vm.cuh): taken when the filters or an aggregate argument do not lower to
// ranges / affine products.  Same group keys, planes and finalisation as the generic pipeline.
static int try_vm(pg_plan *plan, const Node &aggn, const Node &scan, std::unique_ptr<Pipeline> *out, std::string *why)
{
    const pg_table *t = plan->slots[(size_t)scan.slot];
    if (aggn.groups.size() > 2) { *why = "more than two group keys"; return PG_EUNSUPPORTED; }
    if (!aggn.having.empty()) { *why = "HAVING"; return PG_EUNSUPPORTED; }
    std::unique_ptr<GenericPipeline> p(new GenericPipeline());
    p->table = t;
    p->vm = true;
    p->nulls = true;                    // valid-input counts are always kept: an expression can be NULL without a NULL column
    p->nkeys = (int)aggn.groups.size();
    p->cc.tables[0] = t;
    std::vector<uint8_t> luts(512, 0);
    int dims[2] = {1, 1};
    for (int k = 0; k < p->nkeys; k++) {
        const Expr &ge = aggn.groups[(size_t)k];
        if (ge.kind != PG_TK_COL) { *why = "group key is not a column"; return PG_EUNSUPPORTED; }
        const Column &col = t->cols[(size_t)ge.idx];
        if (!is_byte_family(col.type) || col.any_nulls()) { *why = "group key is not a non-null byte-coded column"; return PG_EUNSUPPORTED; }
        p->key_col[k] = ge.idx;
        dense_codes(col, &p->vals[k], &luts[(size_t)k * 256]);
        dims[k] = (int)p->vals[k].size();
    }
    p->G = dims[0] * dims[1];
    if (p->G > 64) { *why = "more than 64 dense groups"; return PG_EUNSUPPORTED; }
    Resolver scope = [&](int idx, Src *s) { if (idx < 0 || idx >= (int)t->cols.size()) return false; s->side = 0; s->col = idx; s->mark = false; return true; };
    VmAggParams &q = p->vprm;
    q.nrows = t->nrows;
    q.row_base = t->global_offset;
    std::vector<const Expr *> fl;
    for (auto &f : scan.filters) fl.push_back(&f);
    if (!p->cc.compile_filters(fl, scope, &q.pred0, &q.pred1)) { *why = "filter: " + p->cc.why; return PG_EUNSUPPORTED; }
    q.nkeys = p->nkeys;
    q.key0 = p->nkeys > 0 ? (const uint8_t *)t->cols[(size_t)p->key_col[0]].d_data : nullptr;
    q.key1 = p->nkeys > 1 ? (const uint8_t *)t->cols[(size_t)p->key_col[1]].d_data : nullptr;
    q.n1 = dims[1];
    q.ngroups = p->G;
    p->prm.n1 = dims[1];
    p->plane_kind = {GEN_SUM};
    p->plane_scale = {0};
    p->aggs = aggn.aggs;
    int nacc = 0;
    for (size_t i = 0; i < aggn.aggs.size(); i++) {
        const AggExpr &a = aggn.aggs[i];
        if (a.fn == PG_AGG_COUNT && a.star) { p->plane.push_back(0); p->agg_is_int.push_back(true); continue; }
        const int kind = a.fn == PG_AGG_COUNT ? GEN_COUNTV : a.fn == PG_AGG_MIN ? GEN_MIN : a.fn == PG_AGG_MAX ? GEN_MAX : GEN_SUM;
        if (a.fn != PG_AGG_COUNT && a.fn != PG_AGG_MIN && a.fn != PG_AGG_MAX && a.fn != PG_AGG_SUM && a.fn != PG_AGG_AVG) { *why = "aggregate function"; return PG_EUNSUPPORTED; }
        int k = 0;
        if (nacc >= GEN_MAXACC) { *why = "more than 8 aggregate arguments"; return PG_EUNSUPPORTED; }
        q.a0[nacc] = p->cc.ncode;
        if (!p->cc.compile(a.arg, scope, &k)) { *why = "aggregate argument: " + p->cc.why; return PG_EUNSUPPORTED; }
        q.a1[nacc] = p->cc.ncode;
        const int sb = p->cc.scale_bound(a.arg, scope);
        if (kind != GEN_COUNTV) {
            if (k != RVK_INT && k != RVK_DEC) { *why = "aggregate over a non-numeric expression"; return PG_EUNSUPPORTED; }
            if (sb < 0 || sb > 18) { *why = "aggregate over a quotient (no fixed scale to accumulate at)"; return PG_EUNSUPPORTED; }
        }
        q.kind[nacc] = kind;
        q.ascale[nacc] = kind == GEN_COUNTV ? 0 : sb;
        const bool is_int = a.ltype == PG_LT_HUGEINT || a.ltype == PG_LT_DOUBLE || a.ltype == PG_LT_INTEGER || a.ltype == PG_LT_BIGINT;
        if (is_int && kind != GEN_COUNTV && (k != RVK_INT || sb != 0)) { *why = "integer aggregate over a scaled value"; return PG_EUNSUPPORTED; }
        if (!is_int && a.ltype != PG_LT_DECIMAL) { *why = "aggregate result type"; return PG_EUNSUPPORTED; }
        if ((kind == GEN_MIN || kind == GEN_MAX) && is_int) { *why = "min/max are DECIMAL only in the reference"; return PG_EUNSUPPORTED; }
        p->plane.push_back(nacc + 1);
        p->plane_kind.push_back(kind == GEN_COUNTV ? (int)GEN_SUM : kind);
        p->plane_scale.push_back(q.ascale[nacc]);
        p->agg_is_int.push_back(is_int);
        nacc++;
    }
    q.nacc = nacc;
    p->prm.nacc = nacc;
    p->P = 1 + 2 * nacc;
    for (int a = 0; a < nacc; a++) { p->plane_kind.push_back(GEN_SUM); p->plane_scale.push_back(0); }
    for (auto &o : aggn.outs) {
        if (o.first == 0 && (o.second < 0 || o.second >= p->nkeys)) { *why = "bad group output index"; return PG_EUNSUPPORTED; }
        if (o.first == 1 && (o.second < 0 || o.second >= (int)aggn.aggs.size())) { *why = "bad aggregate output index"; return PG_EUNSUPPORTED; }
        if (o.first != 0 && o.first != 1) { *why = "bad output kind"; return PG_EUNSUPPORTED; }
    }
    p->outs = aggn.outs;
    for (int i = 0; i < p->cc.ncols; i++) p->bytes_per_row += p->cc.code.cols[i].col.width;
    for (int k = 0; k < p->nkeys; k++) p->bytes_per_row += 1;
    p->NT = 256;
    while (p->NT >= 64 && (size_t)p->G * p->P * p->NT * 8 > (size_t)200 * 1024) p->NT /= 2;
    if (p->NT < 64) { *why = "group tables do not fit in shared memory"; return PG_EUNSUPPORTED; }
    p->smem = (size_t)p->G * p->P * p->NT * 8;
    const void *kern = p->NT == 256 ? (const void *)vm_scanagg_kernel<256> : p->NT == 128 ? (const void *)vm_scanagg_kernel<128> : (const void *)vm_scanagg_kernel<64>;
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    const i64 g = sms_times(kern, p->NT, p->smem);
    p->grid = (int)std::max<i64>(std::min(g, (t->nrows + p->NT - 1) / p->NT), 1);
    // a CTA's int64 partial sums stay exact while |value| * rows per CTA < 2^62 (largest shard, smallest full grid)
    const i64 gmin_grid = std::max<i64>(1, std::min<i64>(g, (t->max_rows() + p->NT - 1) / p->NT));
    const i64 rows_per_cta = (t->max_rows() + gmin_grid - 1) / gmin_grid + p->NT;
    q.absmax = (i64)(((i128)1 << 62) / (i128)std::max<i64>(rows_per_cta, 1));
    PG_TRY(p->d_part.alloc(sizeof(i64) * (size_t)p->grid * (size_t)p->G * (size_t)p->P));
    PG_TRY(p->d_final.alloc(p->rank_bytes()));
    PG_TRY(p->d_gather.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->h_final.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->d_luts.alloc(512));
    PG_TRY(p->d_kinds.alloc(sizeof(int) * (size_t)p->G * (size_t)p->P));
    PG_TRY(p->d_code.alloc(sizeof(RvCode)));
    PG_TRY(p->d_err.alloc(4));
    std::vector<int> kinds((size_t)p->G * (size_t)p->P);
    for (int v = 0; v < p->G * p->P; v++) kinds[(size_t)v] = p->plane_kind[(size_t)(v % p->P)];
    PG_CUDA(cudaMemcpyAsync(p->d_luts.p, luts.data(), 512, cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaMemcpyAsync(p->d_kinds.p, kinds.data(), sizeof(int) * kinds.size(), cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaMemcpyAsync(p->d_code.p, &p->cc.code, sizeof(RvCode), cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaStreamSynchronize(ctx().stream));
    q.luts = p->d_luts.as<uint8_t>();
    q.code = p->d_code.as<RvCode>();
    q.err = p->d_err.as<int>();
    char buf[384];
    snprintf(buf, sizeof buf,
             "ScanAgg[expression programs] table=%s rows=%lld kernel=vm_scanagg_kernel<%d> grid=%d smem=%zu groups=%dx%d "
             "instructions=%d columns=%d accumulators=%d stored bytes/row=%lld",
             t->name.c_str(), (long long)t->nrows, p->NT, p->grid, p->smem, dims[0], dims[1], p->cc.ncode, p->cc.ncols, nacc, (long long)p->bytes_per_row);
    p->explain = buf;
    *out = std::move(p);
    return PG_OK;
}

// ---------------------------------------------------------------------- entry --

int build_scan_agg(pg_plan *plan, const Node &aggn, const Node &scan, std::unique_ptr<Pipeline> *out)
{
    const pg_table *t = plan->slots[(size_t)scan.slot];
    LowerCtx cx;
    cx.table = t;
    cx.allow_nulls = true;     // columns that hold NULLs route the plan to the NULL-aware generic kernel
    std::vector<Range> ranges;
    std::string why_vm;
    auto vm_path = [&](const std::string &because) -> int {
        // not a conjunction of ranges / not a product of affine factors: evaluate the expressions themselves per row
        const int s = try_vm(plan, aggn, scan, out, &why_vm);
        if (s != PG_EUNSUPPORTED) return s;
        PG_FAIL(PG_EUNSUPPORTED, "%s; expression-driven kernel: %s", because.c_str(), why_vm.c_str());
    };
    if (getenv("PG_FORCE_VM") && atoi(getenv("PG_FORCE_VM"))) return vm_path("PG_FORCE_VM");
    if (!lower_filters(cx, scan.filters, ranges)) return vm_path("scan filter not a conjunction of ranges (" + cx.why + ")");
    std::vector<AffProd> args(aggn.aggs.size());
    for (size_t i = 0; i < aggn.aggs.size(); i++) {
        const AggExpr &a = aggn.aggs[i];
        if (a.fn == PG_AGG_COUNT) {
            // count(*) is rewritten to count(<first column>) by the binder (builder_binder.go:207-228);
            // on a NOT NULL column that is the row count
            if (!a.star) {
                const Expr *e = strip_value_preserving_casts(&a.arg);
                if (e->kind != PG_TK_COL || e->idx < 0 || e->idx >= (int)t->cols.size())
                    return vm_path("count() over a computed argument");
                if (t->cols[(size_t)e->idx].any_nulls()) {      // count(col) = rows where col is not NULL
                    Factor f;
                    f.col = e->idx;
                    args[i].f.push_back(f);
                    cx.saw_nulls = true;
                }
            }
            continue;
        }
        if (a.star) PG_FAIL(PG_EINVAL, "aggregate %zu has no argument", i);
        if (!lower_affprod(cx, a.arg, args[i])) return vm_path("aggregate argument not a product of affine factors (" + cx.why + ")");
    }
    std::string why1 = "NULLs or code-set predicates present", why2 = why1, why3;
    const char *force = getenv("PG_FORCE_GENERIC");      // testing: exercise the shape-agnostic kernel on every plan
    int s = PG_EUNSUPPORTED;
    bool any_set = false;
    for (auto &r : ranges) any_set = any_set || r.is_set || r.like;
    if (!(force && atoi(force)) && !cx.saw_nulls && !any_set) {
        s = try_sumprod(plan, aggn, scan, ranges, args, out, &why1);
        if (s != PG_EUNSUPPORTED) return s;
        s = try_lowcard(plan, aggn, scan, ranges, args, out, &why2);
        if (s != PG_EUNSUPPORTED) return s;
    }
    s = try_generic(plan, aggn, scan, ranges, args, out, &why3);
    if (s != PG_EUNSUPPORTED) return s;
    PG_FAIL(PG_EUNSUPPORTED, "no scan-aggregate kernel for this shape (sumprod: %s; lowcard: %s; generic: %s)",
            why1.c_str(), why2.c_str(), why3.c_str());
}

}  // namespace pg

// Host-only entry point (no CUDA call): the product's own cross-rank merge, for tests that run without a GPU
// (tests/test_dist_cpu.py feeds it per-rank partials exchanged over gloo).
extern "C" int pg_host_merge_partials(const void *gathered, int64_t rank_bytes, int nranks, int nvals, int ngroups,
                                      uint64_t *out_totals /* [nvals][2] */, int64_t *out_first /* [ngroups] */)
{
    using namespace pg;
    if (!gathered || !out_totals || !out_first || nranks < 1 || nvals < 0 || ngroups < 0 || rank_bytes < (int64_t)nvals * 16 + (int64_t)ngroups * 8)
        PG_FAIL(PG_EINVAL, "pg_host_merge_partials: bad arguments");
    std::vector<i128> tot((size_t)nvals);
    std::vector<i64> first((size_t)ngroups);
    merge_rank_partials((const char *)gathered, (size_t)rank_bytes, nranks, nvals, ngroups, tot.data(), first.data());
    for (int v = 0; v < nvals; v++) { out_totals[2 * v] = (u64)tot[(size_t)v]; out_totals[2 * v + 1] = (u64)(tot[(size_t)v] >> 64); }
    for (int g = 0; g < ngroups; g++) out_first[g] = first[(size_t)g];
    return PG_OK;
}

