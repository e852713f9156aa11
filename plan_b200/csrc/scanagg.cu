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

namespace pg {

static i128 maxabs(i64 lo, i64 hi)
{
    i128 a = lo < 0 ? -(i128)lo : (i128)lo, b = hi < 0 ? -(i128)hi : (i128)hi;
    return a > b ? a : b;
}

static void clamp_to_stats(Range &r, const Column &c)
{
    if (c.stats_ok) {
        if (r.lo < c.vmin) r.lo = c.vmin;
        if (r.hi > c.vmax) r.hi = c.vmax;
    }
}

static Range find_range(const std::vector<Range> &rs, int col)
{
    for (auto &r : rs) if (r.col == col) return r;
    Range r;
    r.col = col;
    return r;
}

static int grid_for(const void *kernel, int threads, size_t smem, i64 ntiles)
{
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    i64 g = (i64)ctx().prop.multiProcessorCount * per_sm;
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

// ------------------------------------------------------------------- sumprod --

struct SumProdPipeline : Pipeline {
    const pg_table *table = nullptr;
    SumProdParams prm{};
    bool has_a = false, has_b = false;
    int grid = 1, vscale = 0;
    AggExpr agg;
    std::vector<std::pair<int, int>> outs;
    int bytes_per_row = 16;
    DevBuf d_part, d_final, d_gather;
    PinBuf h_final;
    EventPair ev_all, ev_main;

    int launch()
    {
        cudaStream_t st = ctx().stream;
#define PG_SP(A, B) sumprod_kernel<A, B, 4><<<grid, SA_THREADS, 0, st>>>(prm, d_part.as<i64>())
        if (has_a && has_b) PG_SP(true, true);
        else if (has_a) PG_SP(true, false);
        else if (has_b) PG_SP(false, true);
        else PG_SP(false, false);
#undef PG_SP
        PG_CUDA(cudaGetLastError());
        return PG_OK;
    }

    int run(pg_result *res) override
    {
        Context &c = ctx();
        cudaStream_t st = c.stream;
        PG_TRY(ev_all.init());
        PG_TRY(ev_main.init());
        PG_CUDA(cudaEventRecord(ev_all.a, st));
        PG_CUDA(cudaEventRecord(ev_main.a, st));
        PG_TRY(launch());
        PG_CUDA(cudaEventRecord(ev_main.b, st));
        finalize128_kernel<<<1, 32, 0, st>>>(d_part.as<i64>(), grid, 2, d_final.as<u64>());
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
                // TODO(rounding regime): a total beyond 19 digits needs the sequential
                // half-even emulation; until then refuse instead of rounding differently.
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
    if (type_size(ca.type) != 8 || type_size(cb.type) != 8) { *why = "factor columns are not 64-bit"; return PG_EUNSUPPORTED; }
    std::unique_ptr<SumProdPipeline> p(new SumProdPipeline());
    p->table = t;
    p->agg = aggn.aggs[0];
    p->outs = aggn.outs;
    for (auto &o : aggn.outs) if (o.first != 1 || o.second != 0) { *why = "output list refers to something else than the aggregate"; return PG_EUNSUPPORTED; }
    p->vscale = ap.vscale();
    SumProdParams &q = p->prm;
    q.nrows = t->nrows;
    q.fa = (const i64 *)ca.d_data;
    q.fb = (const i64 *)cb.d_data;
    Range ra = find_range(ranges, ap.f[0].col), rb = find_range(ranges, ap.f[1].col);
    clamp_to_stats(ra, ca);
    clamp_to_stats(rb, cb);
    q.fa_lo = ra.lo; q.fa_hi = ra.hi; q.fb_lo = rb.lo; q.fb_hi = rb.hi;
    q.pa = q.pb = nullptr;
    q.a_lo = q.b_lo = INT32_MIN;
    q.a_hi = q.b_hi = INT32_MAX;
    int n32 = 0;
    for (auto &r : ranges) {
        if (r.col == ap.f[0].col || r.col == ap.f[1].col) continue;
        const Column &col = t->cols[(size_t)r.col];
        if (type_size(col.type) != 4) { *why = "predicate on a column that is neither a factor nor 32-bit"; return PG_EUNSUPPORTED; }
        if (n32 == 2) { *why = "more than two 32-bit predicate columns"; return PG_EUNSUPPORTED; }
        int lo = (int)std::max<i64>(r.lo, INT32_MIN), hi = (int)std::min<i64>(r.hi, INT32_MAX);
        if (r.lo > r.hi) { lo = 1; hi = 0; }
        if (n32 == 0) { q.pa = (const int *)col.d_data; q.a_lo = lo; q.a_hi = hi; p->has_a = true; }
        else { q.pb = (const int *)col.d_data; q.b_lo = lo; q.b_hi = hi; p->has_b = true; }
        n32++;
    }
    p->bytes_per_row = 16 + 4 * n32;
    i64 ntiles = (t->nrows + SA_TILE - 1) / SA_TILE;
    const void *kern = p->has_a && p->has_b ? (const void *)sumprod_kernel<true, true, 4>
                       : p->has_a           ? (const void *)sumprod_kernel<true, false, 4>
                       : p->has_b           ? (const void *)sumprod_kernel<false, true, 4>
                                            : (const void *)sumprod_kernel<false, false, 4>;
    p->grid = grid_for(kern, SA_THREADS, 0, ntiles);
    // per-CTA int64 partials must be exact: bound them with the column statistics
    i128 per_row = maxabs(ra.lo, ra.hi) * maxabs(rb.lo, rb.hi);
    i128 rows_per_cta = (i128)((ntiles + p->grid - 1) / p->grid) * SA_TILE;
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
             "ScanAgg[sumprod] table=%s rows=%lld kernel=sumprod_kernel<%d,%d,4> grid=%d block=%d "
             "bytes/row=%d ranges: a=[%d,%d] b=[%d,%d] fa=[%lld,%lld] fb=[%lld,%lld] value_scale=%d",
             t->name.c_str(), (long long)t->nrows, (int)p->has_a, (int)p->has_b, p->grid, SA_THREADS,
             p->bytes_per_row, q.a_lo, q.a_hi, q.b_lo, q.b_hi, (long long)q.fa_lo, (long long)q.fa_hi,
             (long long)q.fb_lo, (long long)q.fb_hi, p->vscale);
    p->explain = buf;
    *out = std::move(p);
    return PG_OK;
}

// ------------------------------------------------------------- lowcard chain --

struct LowcardPipeline : Pipeline {
    const pg_table *table = nullptr;
    LowcardParams prm{};
    bool has_key1 = false;
    int grid = 1, G = 1;
    size_t smem = 0;
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
    bool nonneg[LC_K] = {true, true, true, true, true, true};   // slot values proven >= 0 from statistics
    int emulations = 0;                                         // (group, slot) sums that took the ordered path

    // ranks whose partials are merged: a replicated table is complete on every rank
    int nranks() const { return table->dist == PG_DIST_REPLICATED ? 1 : ctx().world; }
    int myrank() const { return table->dist == PG_DIST_REPLICATED ? 0 : ctx().rank; }

    size_t rank_bytes() const { return (size_t)G * LC_K * 16 + (size_t)LC_MAXG * 8; }

    static constexpr int ORD_CHUNK = 16;   // tiles composed per summary on the device (tail pass)
    bool part_on_host = false;             // run() already copied the per-CTA partials to h_part
    // scratch of the ordered-rounding path, allocated once (no cudaMalloc while executing)
    DevBuf d_ord, d_contrib;
    PinBuf h_ord, h_part, h_tile, h_contrib;

    int ensure_ord_buffers()
    {
        if (d_ord.p) return PG_OK;
        const i64 ntiles = (prm.nrows + LC_TILE - 1) / LC_TILE;
        PG_TRY(d_ord.alloc(sizeof(OrdSummary) * (size_t)std::max<i64>(ntiles, 1)));
        PG_TRY(h_ord.alloc(sizeof(OrdSummary) * (size_t)std::max<i64>(ntiles, 1)));
        PG_TRY(h_part.alloc(sizeof(i64) * (size_t)grid * (size_t)G * LC_K));
        PG_TRY(h_tile.alloc((size_t)LC_TILE * 32 + 512));
        PG_TRY(d_contrib.alloc(64 * (size_t)(ctx().world + 1)));
        PG_TRY(h_contrib.alloc(64 * (size_t)(ctx().world + 1)));
        return PG_OK;
    }

    // summaries of tiles [tb, te) for (group, slot), `chunk` consecutive tiles composed per summary
    // (chunk = 1 where the crossing tile is searched) -> h_ord (pinned)
    int ord_summaries(int g, int s, i64 tb, i64 te, int chunk, const OrdSummary **out, i64 *n, i64 at = 0, bool sync = true)
    {
        *n = te > tb ? (te - tb + chunk - 1) / chunk : 0;
        *out = h_ord.as<OrdSummary>() + at;
        if (te <= tb) return PG_OK;
        cudaStream_t st = ctx().stream;
        OrdParams op;
        op.base = prm;
        op.group = g;
        op.slot = s;
        op.tile_begin = tb;
        op.tile_end = te;
        op.chunk = chunk;
        int gr = (int)std::min<i64>(*n, (i64)ctx().prop.multiProcessorCount * 8);
        if (has_key1) ord_tile_kernel<true><<<gr, LC_THREADS, 0, st>>>(op, d_ord.as<OrdSummary>() + at);
        else ord_tile_kernel<false><<<gr, LC_THREADS, 0, st>>>(op, d_ord.as<OrdSummary>() + at);
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaMemcpyAsync(h_ord.as<OrdSummary>() + at, d_ord.as<OrdSummary>() + at, sizeof(OrdSummary) * (size_t)*n, cudaMemcpyDeviceToHost, st));
        if (sync) PG_CUDA(cudaStreamSynchronize(st));
        return PG_OK;
    }

    // The value the reference's sequential Decimal.Add fold holds for (group g, slot s) when the
    // exact total needs 20 digits (see the comment above ord_tile_kernel).  Collective: every
    // rank calls it with the same arguments.  rank_tot[r] = exact total of rank r.
    int emulate_rounded_sum(int g, int s, const std::vector<i128> &rank_tot, int vscale, HDec *out)
    {
        Context &c = ctx();
        cudaStream_t st = c.stream;
        const i128 THR = (i128)10000000000000000000ULL;    // 10^19
        i128 total = 0;
        for (auto v : rank_tot) total += v;
        if (!nonneg[s]) PG_FAIL(PG_EOVERFLOW, "sum needs 20 digits and its addends may be negative: rounding order emulation not available");
        if (total >= THR * 10) PG_FAIL(PG_EOVERFLOW, "sum needs more than 20 digits");
        if (!prm.contig) PG_FAIL(PG_EOVERFLOW, "internal: ordered partials were not produced");
        if (vscale < 1) PG_FAIL(PG_EOVERFLOW, "decimal overflow: integer part exceeds 19 digits");
        int rstar = 0;
        i128 P0 = 0;
        while (P0 + rank_tot[(size_t)rstar] < THR) { P0 += rank_tot[(size_t)rstar]; rstar++; }
        const i64 ntiles = (prm.nrows + LC_TILE - 1) / LC_TILE;
        PG_TRY(ensure_ord_buffers());
        const OrdSummary *sums = nullptr;
        i64 nsums = 0;
        // per-rank contribution: absolute state (rank == rstar) or a transducer summary (rank > rstar)
        struct Contrib { u64 kind, s_lo, s_hi, q_lo, q_hi; u64 c0, c1, p0p1; } mine{};
        if (myrank() == rstar) {
            // a. which CTA range crosses
            const i64 *part = h_part.as<i64>();
            if (!part_on_host) {
                PG_CUDA(cudaMemcpyAsync(h_part.p, d_part.p, sizeof(i64) * (size_t)grid * (size_t)G * LC_K, cudaMemcpyDeviceToHost, st));
                PG_CUDA(cudaStreamSynchronize(st));
            }
            i64 per = (ntiles + grid - 1) / grid;
            i128 P = P0;
            int cstar = 0;
            for (; cstar < grid; cstar++) {
                i128 v = part[(size_t)cstar * (size_t)G * LC_K + (size_t)g * LC_K + (size_t)s];
                if (P + v >= THR) break;
                P += v;
            }
            if (cstar == grid) PG_FAIL(PG_ECUDA, "internal: crossing CTA not found");
            // b. per-tile summaries of the crossing CTA's tiles and, in the same round trip, chunk summaries
            //    of every tile after that CTA's range
            i64 tb = (i64)cstar * per, te = std::min<i64>(ntiles, tb + per);
            const OrdSummary *tail = nullptr;
            i64 ntail = 0;
            PG_TRY(ord_summaries(g, s, tb, te, 1, &sums, &nsums, 0, false));
            PG_TRY(ord_summaries(g, s, te, ntiles, ORD_CHUNK, &tail, &ntail, te - tb, true));
            i64 tstar = tb;
            for (; tstar < te; tstar++) {
                i128 v = sums[tstar - tb].sum_x;
                if (P + v >= THR) break;
                P += v;
            }
            if (tstar == te) PG_FAIL(PG_ECUDA, "internal: crossing tile not found");
            // c. that tile row by row, exactly as the reference would add them
            i64 row0 = tstar * LC_TILE;
            int n = (int)std::min<i64>(LC_TILE, prm.nrows - row0);
            // pinned staging: [A | B | C] int64, [pred] int32, [key0 | key1] bytes, [luts]
            i64 *h_a = h_tile.as<i64>(), *h_b = h_a + LC_TILE, *h_c = h_b + LC_TILE;
            int *h_pred = (int *)(h_c + LC_TILE);
            uint8_t *h_k0 = (uint8_t *)(h_pred + LC_TILE), *h_k1 = h_k0 + LC_TILE, *h_lut = h_k1 + LC_TILE;
            PG_CUDA(cudaMemcpyAsync(h_pred, prm.pred + row0, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaMemcpyAsync(h_k0, prm.key0 + row0, (size_t)n, cudaMemcpyDeviceToHost, st));
            if (has_key1) PG_CUDA(cudaMemcpyAsync(h_k1, prm.key1 + row0, (size_t)n, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaMemcpyAsync(h_a, prm.A + row0, sizeof(i64) * (size_t)n, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaMemcpyAsync(h_b, prm.B + row0, sizeof(i64) * (size_t)n, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaMemcpyAsync(h_c, prm.C + row0, sizeof(i64) * (size_t)n, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaMemcpyAsync(h_lut, prm.luts, 512, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            bool rounded = false;
            u128 S = 0;
            for (int i = 0; i < n; i++) {
                if (h_pred[(size_t)i] < prm.lo || h_pred[(size_t)i] > prm.hi) continue;
                int gg = h_lut[h_k0[(size_t)i]];
                if (has_key1) gg = gg * prm.n1 + h_lut[256 + h_k1[(size_t)i]];
                if (gg != g) continue;
                i64 a = h_a[(size_t)i], b = h_b[(size_t)i], cc = h_c[(size_t)i];
                i64 t2 = a * (prm.c1 + prm.s1 * b);
                i64 x = s == 2 ? a : s == 3 ? t2 : s == 4 ? t2 * (prm.c2 + prm.s2 * cc) : b;
                if (!rounded) {
                    P += x;
                    if (P >= THR) { S = hd_shift_right_even((u128)P, 1); rounded = true; }
                } else {
                    u128 q = (u128)(x / 10), t = S + q;
                    int d = (int)(x % 10);
                    S = t + ((d > 5 || (d == 5 && (t & 1))) ? 1 : 0);
                }
            }
            if (!rounded) PG_FAIL(PG_ECUDA, "internal: crossing row not found");
            // d. the rest of the crossing CTA's tiles (per-tile summaries), then the chunk summaries after it
            for (i64 t = tstar + 1; t < te; t++) S += (u128)sums[t - tb].sum_q + ((S & 1) ? sums[t - tb].c1 : sums[t - tb].c0);
            for (i64 i = 0; i < ntail; i++) S += (u128)tail[i].sum_q + ((S & 1) ? tail[i].c1 : tail[i].c0);
            mine.kind = 1;
            mine.s_lo = (u64)S;
            mine.s_hi = (u64)(S >> 64);
        } else if (myrank() > rstar) {
            PG_TRY(ord_summaries(g, s, 0, ntiles, ORD_CHUNK, &sums, &nsums));
            i128 q = 0;
            u64 cc[2] = {0, 0}, pp[2] = {0, 1};
            for (i64 i = 0; i < nsums; i++) {
                const OrdSummary &o = sums[i];
                q += o.sum_q;
                for (int k = 0; k < 2; k++) {
                    cc[k] += pp[k] ? o.c1 : o.c0;
                    pp[k] = pp[k] ? o.p1 : o.p0;
                }
            }
            mine.kind = 2;
            mine.q_lo = (u64)q;
            mine.q_hi = (u64)((u128)q >> 64);
            mine.c0 = cc[0];
            mine.c1 = cc[1];
            mine.p0p1 = pp[0] | (pp[1] << 1);
        }
        static_assert(sizeof(Contrib) == 64, "Contrib is exchanged as 64 bytes");
        Contrib *all = h_contrib.as<Contrib>() + 1;
        if (nranks() > 1) {
            h_contrib.as<Contrib>()[0] = mine;
            PG_CUDA(cudaMemcpyAsync(d_contrib.p, h_contrib.p, 64, cudaMemcpyHostToDevice, st));
            PG_TRY(comm_allgather(d_contrib.p, (char *)d_contrib.p + 64, 64, st));
            PG_CUDA(cudaMemcpyAsync(all, (char *)d_contrib.p + 64, 64 * (size_t)nranks(), cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
        } else {
            all[0] = mine;
        }
        if (all[(size_t)rstar].kind != 1) PG_FAIL(PG_ECUDA, "internal: crossing rank did not report a state");
        u128 S = ((u128)all[(size_t)rstar].s_hi << 64) | all[(size_t)rstar].s_lo;
        for (int r = rstar + 1; r < nranks(); r++) {
            const Contrib &k = all[(size_t)r];
            u128 q = ((u128)k.q_hi << 64) | k.q_lo;
            S += q + ((S & 1) ? k.c1 : k.c0);
        }
        if (S > (u128)HD_MAXCOEF) PG_FAIL(PG_EOVERFLOW, "sum needs more than 19 digits after rounding");
        out->coef = (u64)S;
        out->scale = vscale - 1;
        out->neg = false;
        emulations++;
        return PG_OK;
    }

    // exact total -> the Decimal the reference would hold (rank_tot = per-rank exact totals)
    int decimal_sum(int g, int s, const std::vector<i128> &rank_tot, HDec *out)
    {
        i128 v = 0;
        for (auto x : rank_tot) v += x;
        if (hd_digits((u128)(v < 0 ? -v : v)) > HD_MAXPREC) return emulate_rounded_sum(g, s, rank_tot, slot_scale[(size_t)s], out);
        if (!hd_from_i128(v, slot_scale[(size_t)s], out)) PG_FAIL(PG_EOVERFLOW, "decimal overflow");
        return PG_OK;
    }

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
        if (has_key1) lowcard_chain_kernel<true, 2><<<grid, LC_THREADS, smem, st>>>(prm, d_part.as<i64>(), d_firstrow);
        else lowcard_chain_kernel<false, 2><<<grid, LC_THREADS, smem, st>>>(prm, d_part.as<i64>(), d_firstrow);
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaEventRecord(ev_main.b, st));
        finalize128_kernel<<<1, 64, 0, st>>>(d_part.as<i64>(), grid, G * LC_K, d_final.as<u64>());
        PG_CUDA(cudaGetLastError());
        const void *src = d_final.p;
        if (nranks() > 1) {
            PG_TRY(comm_allgather(d_final.p, d_gather.p, rank_bytes(), st));
            src = d_gather.p;
        }
        PG_CUDA(cudaMemcpyAsync(h_final.p, src, rank_bytes() * (size_t)nranks(), cudaMemcpyDeviceToHost, st));
        part_on_host = false;
        if (prm.contig) {      // the ordered partials ride along: the rounding path then needs no extra round trip
            PG_TRY(ensure_ord_buffers());
            PG_CUDA(cudaMemcpyAsync(h_part.p, d_part.p, sizeof(i64) * (size_t)grid * (size_t)G * LC_K, cudaMemcpyDeviceToHost, st));
            part_on_host = true;
        }
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        PG_CUDA(cudaStreamSynchronize(st));
        tr.mark("kernels+gather+d2h");

        // merge ranks in order; 128-bit exact
        std::vector<i128> tot((size_t)G * LC_K, 0);
        std::vector<std::vector<i128>> rtot((size_t)G * LC_K, std::vector<i128>((size_t)nranks(), 0));
        std::vector<i64> first((size_t)G, INT64_MAX);
        emulations = 0;
        for (int r = 0; r < nranks(); r++) {
            const char *base = (const char *)h_final.p + rank_bytes() * (size_t)r;
            const u64 *h = (const u64 *)base;
            const i64 *f = (const i64 *)(base + (size_t)G * LC_K * 16);
            for (int v = 0; v < G * LC_K; v++) {
                rtot[(size_t)v][(size_t)r] = make_i128(h[2 * v], h[2 * v + 1]);
                tot[(size_t)v] += rtot[(size_t)v][(size_t)r];
            }
            for (int g = 0; g < G; g++) if (f[g] != 0x7f7f7f7f7f7f7f7fLL && f[g] < first[(size_t)g]) first[(size_t)g] = f[g];
        }
        res->stats.kernel_ms = ev_all.ms();
        res->stats.main_kernel_ms = ev_main.ms();
        res->stats.rows_scanned = table->nrows;
        res->stats.algorithmic_bytes = table->nrows * bytes_per_row;
        res->stats.main_kernel_bytes = res->stats.algorithmic_bytes;
        res->stats.kernel_launches = 2;

        // groups in first-insertion order (aggregate_hash.go:424-438)
        std::vector<int> order;
        for (int g = 0; g < G; g++) if (tot[(size_t)g * LC_K] > 0) order.push_back(g);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return first[(size_t)a] < first[(size_t)b]; });
        i64 selected = 0;
        for (int g : order) selected += (i64)tot[(size_t)g * LC_K];
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
                        PG_TRY(decimal_sum(g, s, rtot[(size_t)g * LC_K + (size_t)s], &d));
                        col.push(to_pg_decimal(d));
                    } else if (is_int) {   // avg(INT32): float64 sum / float64 count
                        i128 mag = v < 0 ? -v : v;
                        if (mag >= ((i128)1 << 53)) PG_FAIL(PG_EOVERFLOW, "avg(INT32): sum not exact in float64");
                        double x = (double)(i64)v / (double)(i64)n;
                        col.push(x);
                    } else {               // avg(DECIMAL) = sum.Quo(count)
                        HDec sd, nd, qd;
                        PG_TRY(decimal_sum(g, s, rtot[(size_t)g * LC_K + (size_t)s], &sd));
                        hd_from_i128(n, 0, &nd);
                        if (!hd_quo(sd, nd, &qd)) PG_FAIL(PG_EOVERFLOW, "avg: decimal division failed");
                        col.push(to_pg_decimal(qd));
                    }
                }
            }
            res->cols.push_back(col);
        }
        res->stats.aux[1] = emulations;
        tr.mark("finalise(+rounding emulation)");
        return PG_OK;
    }
};

static bool prefix_match(const AffProd &a, const AffProd &chain, size_t n)
{
    if (a.f.size() != n || chain.f.size() < n) return false;
    for (size_t i = 0; i < n; i++) if (!(a.f[i] == chain.f[i])) return false;
    return true;
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
        if (!is_byte_family(col.type) || col.has_nulls) { *why = "group key is not a non-null byte-coded column"; return PG_EUNSUPPORTED; }
        p->key_col[k] = ge.idx;
        // dense ids over the codes that occur anywhere (union over ranks so every rank agrees)
        uint32_t present[8];
        memcpy(present, col.present, sizeof present);
        if (ctx().world > 1 && t->dist != PG_DIST_REPLICATED) {
            DevBuf ds, dr;
            PG_TRY(ds.alloc(32));
            PG_TRY(dr.alloc(32 * (size_t)ctx().world));
            PG_CUDA(cudaMemcpyAsync(ds.p, present, 32, cudaMemcpyHostToDevice, ctx().stream));
            PG_TRY(comm_allgather(ds.p, dr.p, 32, ctx().stream));
            std::vector<uint32_t> all(8 * (size_t)ctx().world);
            PG_CUDA(cudaMemcpyAsync(all.data(), dr.p, 32 * (size_t)ctx().world, cudaMemcpyDeviceToHost, ctx().stream));
            PG_CUDA(cudaStreamSynchronize(ctx().stream));
            for (int r = 0; r < ctx().world; r++) for (int w = 0; w < 8; w++) present[w] |= all[(size_t)r * 8 + (size_t)w];
        }
        for (int code = 0; code < 256; code++)
            if (present[code >> 5] & (1u << (code & 31))) {
                luts[(size_t)k * 256 + (size_t)code] = (uint8_t)p->vals[k].size();
                p->vals[k].push_back((uint8_t)code);
            }
        if (p->vals[k].empty()) p->vals[k].push_back(0);   // empty table
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
    for (size_t i = 0; i < aggn.aggs.size(); i++) {
        const AggExpr &a = aggn.aggs[i];
        const AffProd &ap = args[i];
        int s = -1;
        bool is_int = false;
        if (a.fn == PG_AGG_COUNT) { s = 0; is_int = true; }
        else if (a.fn != PG_AGG_SUM && a.fn != PG_AGG_AVG) { *why = "aggregate other than sum/avg/count"; return PG_EUNSUPPORTED; }
        else if (ap.f.size() == 1 && ap.f[0].c == 0 && ap.f[0].s == 1 && type_size(t->cols[(size_t)ap.f[0].col].type) == 4) {
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
    // predicate: at most one range, on a 32-bit column
    if (ranges.size() > 1) { *why = "more than one predicate column"; return PG_EUNSUPPORTED; }
    int colP = -1;
    int lo = INT32_MIN, hi = INT32_MAX;
    if (ranges.size() == 1) {
        colP = ranges[0].col;
        if (type_size(t->cols[(size_t)colP].type) != 4) { *why = "predicate column is not 32-bit"; return PG_EUNSUPPORTED; }
        lo = (int)std::max<i64>(ranges[0].lo, INT32_MIN);
        hi = (int)std::min<i64>(ranges[0].hi, INT32_MAX);
        if (ranges[0].lo > ranges[0].hi) { lo = 1; hi = 0; }
    }
    // every kernel input must exist: alias the missing ones to a column that is read anyway
    if (colA < 0) { *why = "no 64-bit aggregate column"; return PG_EUNSUPPORTED; }
    for (int cidx : {colA, colB, colC}) if (cidx >= 0 && type_size(t->cols[(size_t)cidx].type) != 8) { *why = "chain column is not 64-bit"; return PG_EUNSUPPORTED; }
    LowcardParams &q = p->prm;
    q.nrows = t->nrows;
    q.row_base = t->global_offset;
    q.A = (const i64 *)t->cols[(size_t)colA].d_data;
    q.c1 = 1; q.s1 = 0; q.c2 = 1; q.s2 = 0;
    q.B = q.A; q.C = q.A;
    int ncol8 = 1;
    if (colB >= 0) { q.B = (const i64 *)t->cols[(size_t)colB].d_data; q.c1 = chain.f[1].c; q.s1 = chain.f[1].s; ncol8++; }
    if (colC >= 0) { q.C = (const i64 *)t->cols[(size_t)colC].d_data; q.c2 = chain.f[2].c; q.s2 = chain.f[2].s; ncol8++; }
    int ncol4 = 0;
    if (colQ >= 0) { q.q = (const int *)t->cols[(size_t)colQ].d_data; ncol4++; }
    if (colP >= 0) { q.pred = (const int *)t->cols[(size_t)colP].d_data; if (colP != colQ) ncol4++; }
    if (colQ < 0 && colP < 0) { *why = "no 32-bit column at all"; return PG_EUNSUPPORTED; }
    if (colQ < 0) q.q = q.pred;
    if (colP < 0) q.pred = q.q;
    q.lo = lo; q.hi = hi;
    q.key0 = (const uint8_t *)t->cols[(size_t)p->key_col[0]].d_data;
    q.key1 = p->has_key1 ? (const uint8_t *)t->cols[(size_t)p->key_col[1]].d_data : nullptr;
    q.n1 = dims[1];
    q.ngroups = p->G;
    p->bytes_per_row = 8 * ncol8 + 4 * ncol4 + p->nkeys;
    p->smem = (size_t)p->G * LC_K * LC_THREADS * sizeof(i64);
    const void *kern = p->has_key1 ? (const void *)lowcard_chain_kernel<true, 2> : (const void *)lowcard_chain_kernel<false, 2>;
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    i64 ntiles = (t->nrows + LC_TILE - 1) / LC_TILE;
    p->grid = grid_for(kern, LC_THREADS, p->smem, ntiles);
    // exactness proof from column statistics
    {
        const Column &cA = t->cols[(size_t)colA];
        i128 bound = maxabs(cA.vmin, cA.vmax);
        if (colB >= 0) {
            const Column &cB = t->cols[(size_t)colB];
            i128 f = std::max(maxabs(q.c1 + q.s1 * cB.vmin, q.c1 + q.s1 * cB.vmax), maxabs(cB.vmin, cB.vmax));
            bound *= f > 1 ? f : 1;
        }
        if (colC >= 0) {
            const Column &cC = t->cols[(size_t)colC];
            i128 f = maxabs(q.c2 + q.s2 * cC.vmin, q.c2 + q.s2 * cC.vmax);
            bound *= f > 1 ? f : 1;
        }
        i128 rows_per_cta = (i128)((ntiles + p->grid - 1) / p->grid) * LC_TILE;
        if (bound * rows_per_cta >= ((i128)1 << 62)) { *why = "per-CTA partial sum could exceed int64"; return PG_EUNSUPPORTED; }
        // can any DECIMAL total need 20 digits?  Then the CTAs take contiguous tile runs so that
        // their partials are ordered (sequential-rounding emulation); the bound uses the largest
        // table any rank could hold relative to this one (shards are balanced to within 2x).
        i128 table_bound = bound * (i128)std::max<i64>(t->nrows, 1) * (i128)ctx().world * 2;
        q.contig = table_bound >= (i128)1000000000000000000LL ? 1 : 0;   // 10^18: generous margin
        const char *force = getenv("PG_LOWCARD_CONTIG");
        if (force) q.contig = atoi(force) ? 1 : 0;
        bool a_pos = cA.vmin >= 0;
        bool f1_pos = colB < 0 || (q.c1 + q.s1 * t->cols[(size_t)colB].vmin >= 0 && q.c1 + q.s1 * t->cols[(size_t)colB].vmax >= 0);
        bool f2_pos = colC < 0 || (q.c2 + q.s2 * t->cols[(size_t)colC].vmin >= 0 && q.c2 + q.s2 * t->cols[(size_t)colC].vmax >= 0);
        p->nonneg[2] = a_pos;
        p->nonneg[3] = a_pos && f1_pos;
        p->nonneg[4] = a_pos && f1_pos && f2_pos;
        p->nonneg[5] = colB < 0 || t->cols[(size_t)colB].vmin >= 0;
    }
    PG_TRY(p->d_part.alloc(sizeof(i64) * (size_t)p->grid * (size_t)p->G * LC_K));
    PG_TRY(p->d_final.alloc(p->rank_bytes()));
    PG_TRY(p->d_gather.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->h_final.alloc(p->rank_bytes() * (size_t)p->nranks()));
    PG_TRY(p->d_luts.alloc(512));
    PG_CUDA(cudaMemcpyAsync(p->d_luts.p, luts.data(), 512, cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaStreamSynchronize(ctx().stream));
    q.luts = p->d_luts.as<uint8_t>();
    char buf[512];
    snprintf(buf, sizeof buf,
             "ScanAgg[lowcard-chain] table=%s rows=%lld kernel=lowcard_chain_kernel<%d,2> grid=%d block=%d smem=%zu "
             "groups=%dx%d bytes/row=%d pred=[%d,%d] chain: A*(%lld%+lld*B)*(%lld%+lld*C) tiles=%s",
             t->name.c_str(), (long long)t->nrows, (int)p->has_key1, p->grid, LC_THREADS, p->smem, dims[0], dims[1],
             p->bytes_per_row, lo, hi, (long long)q.c1, (long long)q.s1, (long long)q.c2, (long long)q.s2,
             q.contig ? "contiguous-per-CTA(ordered partials)" : "interleaved");
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
#define PG_GEN(N) do { if (nulls) generic_scanagg_kernel<N, true><<<grid, N, smem, st>>>(prm, d_part.as<i64>(), d_firstrow); \
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
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        PG_CUDA(cudaStreamSynchronize(st));
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
    auto use = [&](int col) { for (auto &u : used) if (u.first == col) return; used.push_back({col, type_size(t->cols[(size_t)col].type)}); };
    for (int k = 0; k < p->nkeys; k++) {
        const Expr &ge = aggn.groups[(size_t)k];
        if (ge.kind != PG_TK_COL) { *why = "group key is not a column"; return PG_EUNSUPPORTED; }
        const Column &col = t->cols[(size_t)ge.idx];
        if (!is_byte_family(col.type) || col.has_nulls) { *why = "group key is not a non-null byte-coded column"; return PG_EUNSUPPORTED; }
        p->key_col[k] = ge.idx;
        use(ge.idx);
        uint32_t present[8];
        memcpy(present, col.present, sizeof present);
        if (ctx().world > 1 && t->dist != PG_DIST_REPLICATED) {
            DevBuf ds, dr;
            PG_TRY(ds.alloc(32));
            PG_TRY(dr.alloc(32 * (size_t)ctx().world));
            PG_CUDA(cudaMemcpyAsync(ds.p, present, 32, cudaMemcpyHostToDevice, ctx().stream));
            PG_TRY(comm_allgather(ds.p, dr.p, 32, ctx().stream));
            std::vector<uint32_t> all(8 * (size_t)ctx().world);
            PG_CUDA(cudaMemcpyAsync(all.data(), dr.p, 32 * (size_t)ctx().world, cudaMemcpyDeviceToHost, ctx().stream));
            PG_CUDA(cudaStreamSynchronize(ctx().stream));
            for (int r = 0; r < ctx().world; r++) for (int w = 0; w < 8; w++) present[w] |= all[(size_t)r * 8 + (size_t)w];
        }
        for (int code = 0; code < 256; code++)
            if (present[code >> 5] & (1u << (code & 31))) {
                luts[(size_t)k * 256 + (size_t)code] = (uint8_t)p->vals[k].size();
                p->vals[k].push_back((uint8_t)code);
            }
        if (p->vals[k].empty()) p->vals[k].push_back(0);
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
        q.pcol[i].p = col.d_data;
        q.pcol[i].width = type_size(col.type);
        q.pcol[i].valid = col.has_nulls ? col.d_valid : nullptr;
        if (col.has_nulls) p->nulls = true;
        q.plo[i] = ranges[i].lo;
        q.phi[i] = ranges[i].hi;
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
                A.fac[f].p = col.d_data;
                A.fac[f].width = type_size(col.type);
                A.fac[f].valid = col.has_nulls ? col.d_valid : nullptr;
                if (col.has_nulls) p->nulls = true;
                A.c[f] = ap.f[f].c;
                A.s[f] = ap.f[f].s;
                use(ap.f[f].col);
                i128 m = std::max(maxabs(ap.f[f].c + ap.f[f].s * col.vmin, ap.f[f].c + ap.f[f].s * col.vmax), (i128)1);
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
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, p->NT, p->smem);
    i64 g = (i64)ctx().prop.multiProcessorCount * std::max(per_sm, 1);
    i64 maxg = (t->nrows + p->NT - 1) / p->NT;
    p->grid = (int)std::max<i64>(std::min(g, maxg), 1);
    i128 rows_per_cta = (i128)((t->nrows + p->grid - 1) / p->grid) + p->NT;
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
             "predicates=%d accumulators=%d bytes/row=%lld%s",
             t->name.c_str(), (long long)t->nrows, p->NT, p->grid, p->smem, dims[0], dims[1], q.npred, q.nacc, (long long)p->bytes_per_row,
             p->nulls ? " nulls=validity-bitmaps" : "");
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
    if (!lower_filters(cx, scan.filters, ranges)) PG_FAIL(PG_EUNSUPPORTED, "scan filter not off-loadable: %s", cx.why.c_str());
    std::vector<AffProd> args(aggn.aggs.size());
    for (size_t i = 0; i < aggn.aggs.size(); i++) {
        const AggExpr &a = aggn.aggs[i];
        if (a.fn == PG_AGG_COUNT) {
            // count(*) is rewritten to count(<first column>) by the binder (builder_binder.go:207-228);
            // on a NOT NULL column that is the row count
            if (!a.star) {
                const Expr *e = strip_value_preserving_casts(&a.arg);
                if (e->kind != PG_TK_COL || e->idx < 0 || e->idx >= (int)t->cols.size())
                    PG_FAIL(PG_EUNSUPPORTED, "count() over a computed argument");
                if (t->cols[(size_t)e->idx].has_nulls) {      // count(col) = rows where col is not NULL
                    Factor f;
                    f.col = e->idx;
                    args[i].f.push_back(f);
                    cx.saw_nulls = true;
                }
            }
            continue;
        }
        if (a.star) PG_FAIL(PG_EINVAL, "aggregate %zu has no argument", i);
        if (!lower_affprod(cx, a.arg, args[i])) PG_FAIL(PG_EUNSUPPORTED, "aggregate argument not off-loadable: %s", cx.why.c_str());
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
