// rows.cu -- ROW-EMITTING operators: Scan / Filter / Project / Join whose parent is not an aggregate.
//
// Reference operators replaced:
//   filterExecutor   /root/reference/pkg/compute/executor_filter.go:12-118   rows for which every filter is TRUE
//   projectExecutor  /root/reference/pkg/compute/executor_project.go:24-82   one output column per expression
//   joinExecutor     /root/reference/pkg/compute/executor_join.go:62-123,209-235 with Scan.Next* of join_scan.go:
//       INNER (:182-299)  every (probe row, matching build row) pair
//       LEFT  (:67-88)    the INNER pairs plus every probe row without a match, build columns NULL
//       SEMI / ANTI (:90-121)  probe rows with / without a match (a NULL key never matches: ANTI keeps the row)
//       MARK  (:123-165)  every probe row plus a boolean "found a match", NULL for a NULL probe key
//   CASE / arithmetic / comparisons in their expressions: rowvm.cuh
// NULL join keys never match (join_table.go:152-195).
//
// Plan shape:  [Project] <- [Filter]* <- ( Scan | Join( [Filter]* Scan , [Filter]* Scan ) ), 1 or 2 integer key columns.
// Execution: the build side's passing rows go into a bucketized hash table (join.cuh jt_insert); one pass over the
// probe side COUNTS the output rows of every probe row (filters above the join are evaluated per candidate pair, so
// they see exactly the joined rows the reference's Filter operator sees), an exclusive scan turns counts into
// offsets, a second pass writes the (probe row, build row) pairs in probe-row order, and the projection kernel
// evaluates the output expressions per pair into columnar buffers that go back to the host in one copy per column
// (the shim re-emits them as <= 2048-row chunks through pg_result_next).
//
// AGGREGATE MODE (Agg <- [Filter]* <- Join(scan, scan) whose expressions do not lower to the affine-product kernels of
// join.cu: CASE WHEN / OR / IN inside the aggregates, TPC-H Q12 / Q14 style): the same pair list feeds vm_scanagg_kernel
// (scanagg_vm.cuh) instead of the projection -- <= 2 byte-coded group keys from either side, sum / avg / min / max /
// count of row programs over the joined row, thread-private shared-memory planes, exact 128-bit totals, the ranks'
// partials all-gathered and merged in rank order, finalisation as in the generic scan aggregate (aggExecutor,
// executor_aggr.go:106-262; SumOp / AvgOp / CountOp / MinMaxOp, function_aggr.go:620-1032).
#include <algorithm>
#include <functional>

#include <cub/device/device_scan.cuh>

#include "hostdec.hpp"
#include "pipeline.hpp"
#include "rowvm_compile.hpp"
#include "scanagg_vm.cuh"

namespace pg {

enum { RO_I32 = 1, RO_I64, RO_DEC, RO_BOOL, RO_U8, RO_ROW0, RO_ROW1 };

struct RowsParams {
    const RvCode *code;
    int jointype;                 // 0: no join
    i64 build_rows, probe_rows;
    int bf0, bf1, pf0, pf1, jf0, jf1;     // filter programs: build scan, probe scan, above the join
    int nkey, bkey[2], pkey[2];   // column slots of the key parts
    JoinTable jt;
    unsigned *cnt;                // [probe_rows + 1]
    const i64 *off;               // [probe_rows + 1]
    i64 *pair0, *pair1;
    int *err;
    int nout;
    int o0[RV_MAXOUT], o1[RV_MAXOUT], okind[RV_MAXOUT];
    void *odata[RV_MAXOUT];
    uint8_t *ovalid[RV_MAXOUT];
    i64 nout_rows;
};

__device__ __forceinline__ bool rows_key(const RowsParams &p, const int *slots, i64 row, i64 *key)
{
    const RvCol &c0 = p.code->cols[slots[0]];
    if (!typed_valid(c0.col, row)) return false;
    i64 k = load_typed(c0.col, row);
    if (p.nkey == 2) {
        const RvCol &c1 = p.code->cols[slots[1]];
        if (!typed_valid(c1.col, row)) return false;
        k = (k << 32) | (load_typed(c1.col, row) & 0xffffffffLL);
    }
    *key = k;
    return true;
}

static __global__ void __launch_bounds__(256)
rows_build_kernel(const RowsParams p)
{
    int err = 0;
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < p.build_rows; r += (i64)gridDim.x * blockDim.x) {
        if (!rv_true(*p.code, p.bf0, p.bf1, -1, r, &err)) continue;
        i64 key;
        if (!rows_key(p, p.bkey, r, &key)) continue;            // NULL keys are not inserted
        jt_insert(p.jt, key, (u64)r);
    }
    if (err) *p.err = err;
}

template <bool EMIT>
static __global__ void __launch_bounds__(256)
rows_probe_kernel(const RowsParams p)
{
    int err = 0;
    // one probe row that passed the inline pre-tests of its scan filter: the interpreted rest of the filter, the join, the
    // filters above the join per candidate pair; COUNT mode writes how many rows it yields, EMIT mode the pairs
    auto process = [&](i64 r, bool pre_done) {
        unsigned n = 0;
        const i64 base = EMIT ? p.off[r] : 0;
        auto emit = [&](i64 b) {
            if (!rv_true(*p.code, p.jf0, p.jf1, r, b, &err)) return;
            if (EMIT) { p.pair0[base + n] = r; p.pair1[base + n] = b; }
            n++;
        };
        if (pre_done ? rv_post(*p.code, p.pf0, p.pf1, r, -1, &err) : rv_true(*p.code, p.pf0, p.pf1, r, -1, &err)) {
            if (p.jointype == 0) {
                emit(-1);
            } else {
                i64 key = 0;
                const bool has_key = rows_key(p, p.pkey, r, &key);
                if (p.jointype == PG_JOIN_INNER || p.jointype == PG_JOIN_LEFT) {
                    bool matched = false;
                    if (has_key) jt_probe(p.jt, key, [&](u64 b) { matched = true; emit((i64)b); });
                    if (!matched && p.jointype == PG_JOIN_LEFT) emit(-1);
                } else {
                    i64 first = -1;
                    if (has_key) jt_probe(p.jt, key, [&](u64 b) { if (first < 0) first = (i64)b; });
                    if (p.jointype == PG_JOIN_SEMI) { if (first >= 0) emit(-1); }
                    else if (p.jointype == PG_JOIN_ANTI) { if (first < 0) emit(-1); }
                    else emit(!has_key ? -2 : first);          // MARK: one row per probe row, the mark rides in the build row id
                }
            }
        }
        if (!EMIT) p.cnt[r] = n;
    };
    if (rv_has_pre(*p.code, p.pf0, p.pf1)) {
        // A selective scan filter: a warp would run the interpreter and the probe while ANY of its 32 rows survives the
        // pre-tests, so the survivors of a block are compacted into a shared-memory queue and processed 256 at a time
        // (the same scheme as vm_scanagg_kernel).  Counts and offsets are indexed by row, so the output order is unchanged.
        constexpr int NT = 256, R = 4;
        __shared__ i64 s_q[(R + 1) * NT];
        __shared__ int s_woff[R][NT / 32];
        __shared__ int s_cnt;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (i64 base = (i64)blockIdx.x * NT * R; base < p.probe_rows; base += (i64)gridDim.x * NT * R) {      // block-uniform trip count
            bool pass[R];
            unsigned m[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const i64 row = base + (i64)k * NT + threadIdx.x;
                const bool in = row < p.probe_rows;
                pass[k] = in && rv_pre(*p.code, p.pf0, p.pf1, row, -1);
                if (!EMIT && in && !pass[k]) p.cnt[row] = 0;
            }
#pragma unroll
            for (int k = 0; k < R; k++) {
                m[k] = __ballot_sync(0xffffffffu, pass[k]);
                if (lane == 0) s_woff[k][warp] = __popc(m[k]);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int acc = s_cnt;
                for (int k = 0; k < R; k++)
                    for (int w = 0; w < NT / 32; w++) { const int c = s_woff[k][w]; s_woff[k][w] = acc; acc += c; }
                s_cnt = acc;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < R; k++)
                if (pass[k]) s_q[s_woff[k][warp] + __popc(m[k] & ((1u << lane) - 1u))] = base + (i64)k * NT + threadIdx.x;
            __syncthreads();
            int n = s_cnt;
            while (n >= NT) {                       // block-uniform
                process(s_q[n - NT + threadIdx.x], true);
                n -= NT;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_cnt = n;
            __syncthreads();
        }
        if ((int)threadIdx.x < s_cnt) process(s_q[threadIdx.x], true);
    } else {
        for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < p.probe_rows; r += (i64)gridDim.x * blockDim.x) process(r, false);
    }
    if (err) *p.err = err;
}

static __global__ void __launch_bounds__(256)
rows_project_kernel(const RowsParams p)
{
    int err = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < p.nout_rows; i += (i64)gridDim.x * blockDim.x) {
        const i64 r0 = p.pair0[i], r1 = p.pair1[i];
        for (int j = 0; j < p.nout; j++) {
            const int kind = p.okind[j];
            if (kind == RO_ROW0 || kind == RO_ROW1) {
                const i64 r = kind == RO_ROW0 ? r0 : r1;
                ((i64 *)p.odata[j])[i] = r;
                p.ovalid[j][i] = r >= 0;
                continue;
            }
            const RvVal v = rv_eval(*p.code, p.o0[j], p.o1[j], r0, r1, &err);
            p.ovalid[j][i] = v.null ? 0 : 1;
            switch (kind) {
            case RO_I32: ((int32_t *)p.odata[j])[i] = v.null ? 0 : (int32_t)(i64)v.v; break;
            case RO_I64: ((i64 *)p.odata[j])[i] = v.null ? 0 : (i64)v.v; break;
            case RO_BOOL: case RO_U8: ((uint8_t *)p.odata[j])[i] = v.null ? 0 : (uint8_t)(i64)v.v; break;
            default: {
                pg_decimal d;
                d.neg = (!v.null && v.v < 0) ? 1u : 0u;
                d.coef = v.null ? 0 : (u64)(v.v < 0 ? -v.v : v.v);
                d.scale = v.null ? 0 : v.scale;
                ((pg_decimal *)p.odata[j])[i] = d;
                break;
            }
            }
        }
    }
    if (err) *p.err = err;
}

namespace {

struct RowsPipeline : Pipeline {
    pg_plan *plan = nullptr;
    RvCompiler cc;
    int slot[2] = {-1, -1};              // table slots of side 0 (probe / scanned) and side 1 (build)
    RowsParams prm{};
    struct Out { int type = 0, width = 0, scale = 0, side = 0, col = -1; };
    std::vector<Out> outs;
    DevBuf d_code, d_slots, d_cnt, d_off, d_pairs, d_err, d_scan_tmp, d_out;
    EventPair ev_all;
    // aggregate mode
    bool agg_mode = false;
    VmAggParams aprm{};
    int G = 1, P = 1, NT = 256, agrid = 1, nkeys = 0, key_side[2] = {0, 0}, key_col[2] = {-1, -1};
    size_t asmem = 0;
    std::vector<uint8_t> vals[2];
    std::vector<AggExpr> aggs;
    std::vector<int> plane, plane_scale, plane_kind;
    std::vector<bool> agg_is_int;
    std::vector<std::pair<int, int>> agg_outs;
    DevBuf d_part, d_final, d_gather, d_luts, d_kinds;
    PinBuf h_final;
    int nranks() const { return tab(0)->dist == PG_DIST_REPLICATED ? 1 : ctx().world; }
    size_t rank_bytes() const { return (size_t)G * P * 16 + 64 * 8; }

    const pg_table *tab(int side) const { return plan->slots[(size_t)slot[side]]; }

    int run(pg_result *res) override
    {
        Context &c = ctx();
        cudaStream_t st = c.stream;
        PG_TRY(ev_all.init());
        Trace tr("rows");
        RowsParams p = prm;
        p.code = d_code.as<RvCode>();
        p.err = d_err.as<int>();
        PG_CUDA(cudaEventRecord(ev_all.a, st));
        PG_CUDA(cudaMemsetAsync(d_err.p, 0, 4, st));
        const int grid_cap = c.prop.multiProcessorCount * 8;
        auto grid_for = [&](i64 n) { return (int)std::max<i64>(std::min<i64>((n + 255) / 256, grid_cap), 1); };
        int launches = 0;
        if (p.jointype) {
            u64 nb = 16;
            while (nb * HT_BUCKET < (u64)std::max<i64>(p.build_rows, 1) * 2) nb <<= 1;
            if (d_slots.bytes < nb * HT_BUCKET * sizeof(longlong2)) PG_TRY(d_slots.alloc(nb * HT_BUCKET * sizeof(longlong2)));
            PG_CUDA(cudaMemsetAsync(d_slots.p, 0x80, nb * HT_BUCKET * sizeof(longlong2), st));
            p.jt = JoinTable{};
            p.jt.slots = d_slots.as<longlong2>();
            p.jt.bucket_mask = nb - 1;
            p.jt.domain = 1;
            if (p.build_rows > 0) {
                rows_build_kernel<<<grid_for(p.build_rows), 256, 0, st>>>(p);
                PG_CUDA(cudaGetLastError());
                launches++;
            }
            tr.mark("build");
        }
        const i64 n = p.probe_rows;
        if (d_cnt.bytes < (size_t)(n + 1) * 4) PG_TRY(d_cnt.alloc((size_t)(n + 1) * 4));
        if (d_off.bytes < (size_t)(n + 1) * 8) PG_TRY(d_off.alloc((size_t)(n + 1) * 8));
        p.cnt = d_cnt.as<unsigned>();
        p.off = d_off.as<i64>();
        PG_CUDA(cudaMemsetAsync(p.cnt + n, 0, 4, st));
        if (n > 0) {
            rows_probe_kernel<false><<<grid_for(n), 256, 0, st>>>(p);
            PG_CUDA(cudaGetLastError());
        }
        size_t tmp = 0;
        PG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, p.cnt, d_off.as<i64>(), (int)(n + 1), st));
        if (d_scan_tmp.bytes < tmp) PG_TRY(d_scan_tmp.alloc(tmp));
        PG_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, tmp, p.cnt, d_off.as<i64>(), (int)(n + 1), st));
        i64 total = 0;
        int err = 0;
        PG_CUDA(cudaMemcpyAsync(&total, d_off.as<i64>() + n, 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(&err, d_err.p, 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        launches += 2;
        tr.mark("count + scan");
        // (aggregate mode: a fault of ONE rank's rows must not keep that rank out of the all-gather its peers are about to
        //  enter -- d_err is sticky, the check follows the collective in run_agg)
        if (!agg_mode) PG_TRY(check_err(err));
        if (d_pairs.bytes < (size_t)std::max<i64>(total, 1) * 16) PG_TRY(d_pairs.alloc((size_t)std::max<i64>(total, 1) * 16));
        p.pair0 = d_pairs.as<i64>();
        p.pair1 = p.pair0 + std::max<i64>(total, 1);
        if (agg_mode) {
            if (total > 0) {
                rows_probe_kernel<true><<<grid_for(n), 256, 0, st>>>(p);
                PG_CUDA(cudaGetLastError());
                launches++;
            }
            return run_agg(res, p, total, launches, tr);
        }
        // output buffers: per column data (8-byte aligned blocks) then validity bytes
        std::vector<size_t> doff(outs.size()), voff(outs.size());
        size_t at = 0;
        for (size_t j = 0; j < outs.size(); j++) {
            const int k = p.okind[j];
            const size_t w = k == RO_I32 ? 4 : (k == RO_BOOL || k == RO_U8) ? 1 : k == RO_DEC ? 16 : 8;
            doff[j] = at; at += ((size_t)total * w + 15) / 16 * 16;
            voff[j] = at; at += ((size_t)total + 15) / 16 * 16;
        }
        if (d_out.bytes < std::max<size_t>(at, 16)) PG_TRY(d_out.alloc(std::max<size_t>(at, 16)));
        for (size_t j = 0; j < outs.size(); j++) { p.odata[j] = (char *)d_out.p + doff[j]; p.ovalid[j] = (uint8_t *)d_out.p + voff[j]; }
        p.nout_rows = total;
        if (total > 0) {
            rows_probe_kernel<true><<<grid_for(n), 256, 0, st>>>(p);
            PG_CUDA(cudaGetLastError());
            rows_project_kernel<<<grid_for(total), 256, 0, st>>>(p);
            PG_CUDA(cudaGetLastError());
            launches += 2;
        }
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        std::vector<std::vector<uint8_t>> hv(outs.size());
        std::vector<std::vector<uint8_t>> hd(outs.size());
        for (size_t j = 0; j < outs.size(); j++) {
            const int k = p.okind[j];
            const size_t w = k == RO_I32 ? 4 : (k == RO_BOOL || k == RO_U8) ? 1 : k == RO_DEC ? 16 : 8;
            hd[j].resize((size_t)total * w);
            hv[j].resize((size_t)total);
            if (total > 0) {
                PG_CUDA(cudaMemcpyAsync(hd[j].data(), p.odata[j], (size_t)total * w, cudaMemcpyDeviceToHost, st));
                PG_CUDA(cudaMemcpyAsync(hv[j].data(), p.ovalid[j], (size_t)total, cudaMemcpyDeviceToHost, st));
            }
        }
        PG_CUDA(cudaMemcpyAsync(&err, d_err.p, 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        tr.mark("emit + project + d2h");
        PG_TRY(check_err(err));
        res->nrows = total;
        for (size_t j = 0; j < outs.size(); j++) {
            ResCol col;
            col.type = outs[j].type;
            col.width = outs[j].width;
            col.scale = outs[j].scale;
            bool any_null = false;
            for (uint8_t v : hv[j]) any_null = any_null || !v;
            if (p.okind[j] == RO_ROW0 || p.okind[j] == RO_ROW1) {
                // VARCHAR carried to the output: fetched by row id from the host-resident column
                const Column &cc = tab(outs[j].side)->cols[(size_t)outs[j].col];
                const i64 *rows = (const i64 *)hd[j].data();
                size_t bytes = 0;
                for (i64 i = 0; i < total; i++)
                    if (rows[i] >= 0) {
                        if (rows[i] + 1 >= (i64)cc.h_off.size()) PG_FAIL(PG_ECUDA, "internal: VARCHAR row %lld out of range", (long long)rows[i]);
                        bytes += (size_t)(cc.h_off[(size_t)rows[i] + 1] - cc.h_off[(size_t)rows[i]]);
                    }
                col.heap.resize(bytes + 1);
                col.data.resize((size_t)total * sizeof(pg_string));
                pg_string *d = (pg_string *)col.data.data();
                size_t pos = 0;
                for (i64 i = 0; i < total; i++) {
                    if (rows[i] < 0) { d[i].data = col.heap.data(); d[i].len = 0; continue; }
                    const size_t b = (size_t)cc.h_off[(size_t)rows[i]], len = (size_t)cc.h_off[(size_t)rows[i] + 1] - b;
                    memcpy(col.heap.data() + pos, cc.h_bytes.data() + b, len);
                    d[i].data = col.heap.data() + pos;
                    d[i].len = (int64_t)len;
                    pos += len;
                }
            } else {
                col.data = std::move(hd[j]);
                if (col.type == PG_T_DICT8) col.dict = tab(outs[j].side)->cols[(size_t)outs[j].col].dict;
            }
            if (any_null) col.valid = std::move(hv[j]);
            res->cols.push_back(std::move(col));
        }
        res->stats.kernel_ms = ev_all.ms();
        res->stats.main_kernel_ms = res->stats.kernel_ms;
        res->stats.rows_scanned = n;
        res->stats.kernel_launches = launches;
        res->stats.aux[0] = n;
        res->stats.aux[1] = total;
        res->stats.aux[6] = total;
        return PG_OK;
    }

    // the pair list -> group totals (collective when the probe table is sharded: every rank all-gathers its partials)
    int run_agg(pg_result *res, const RowsParams &p, i64 total, int launches, Trace &tr)
    {
        cudaStream_t st = ctx().stream;
        VmAggParams q = aprm;
        q.code = d_code.as<RvCode>();
        q.err = d_err.as<int>();
        q.nrows = total;
        q.pair0 = p.pair0;
        q.pair1 = p.pair1;
        q.luts = d_luts.as<uint8_t>();
        i64 *d_firstrow = (i64 *)((char *)d_final.p + (size_t)G * P * 16);
        PG_CUDA(cudaMemsetAsync(d_firstrow, 0x7f, 64 * 8, st));
        const int grid = (int)std::max<i64>(std::min<i64>((total + NT - 1) / NT, agrid), 1);
        // a CTA's int64 partial sums are exact while |value| x its pairs < 2^62 (rank-local fact: reported after the collective)
        const bool inexact = (i128)q.absmax * (i128)((total + grid - 1) / grid + NT) >= ((i128)1 << 62);
        if (NT == 256) vm_scanagg_kernel<256><<<grid, 256, asmem, st>>>(q, d_part.as<i64>(), d_firstrow);
        else if (NT == 128) vm_scanagg_kernel<128><<<grid, 128, asmem, st>>>(q, d_part.as<i64>(), d_firstrow);
        else vm_scanagg_kernel<64><<<grid, 64, asmem, st>>>(q, d_part.as<i64>(), d_firstrow);
        PG_CUDA(cudaGetLastError());
        finalize_generic_kernel<<<(G * P + 63) / 64, 64, 0, st>>>(d_part.as<i64>(), grid, G * P, d_kinds.as<int>(), d_final.as<u64>());
        PG_CUDA(cudaGetLastError());
        const void *src = d_final.p;
        const int R = nranks();
        if (R > 1) {
            PG_TRY(comm_allgather(d_final.p, d_gather.p, rank_bytes(), st));
            src = d_gather.p;
        }
        PG_CUDA(cudaMemcpyAsync(h_final.p, src, rank_bytes() * (size_t)R, cudaMemcpyDeviceToHost, st));
        int err = 0;
        PG_CUDA(cudaMemcpyAsync(&err, d_err.p, 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaEventRecord(ev_all.b, st));
        PG_CUDA(cudaStreamSynchronize(st));
        tr.mark("aggregate over the pairs + gather");
        PG_TRY(check_err(err));
        if (inexact) PG_FAIL(PG_EUNSUPPORTED, "aggregate over %lld joined rows: exactness of the per-CTA partial sums cannot be proven", (long long)total);
        std::vector<i128> tot((size_t)G * P);
        std::vector<i64> first((size_t)G, INT64_MAX);
        for (int v = 0; v < G * P; v++) {
            const int kind = plane_kind[(size_t)(v % P)];
            tot[(size_t)v] = kind == GEN_SUM ? 0 : kind == GEN_MIN ? (i128)INT64_MAX : (i128)INT64_MIN;
        }
        for (int r = 0; r < R; r++) {
            const char *base = (const char *)h_final.p + rank_bytes() * (size_t)r;
            const u64 *h = (const u64 *)base;
            const i64 *f = (const i64 *)(base + (size_t)G * P * 16);
            for (int v = 0; v < G * P; v++) {
                const i128 x = (i128)(((u128)h[2 * v + 1] << 64) | (u128)h[2 * v]);
                const int kind = plane_kind[(size_t)(v % P)];
                if (kind == GEN_SUM) tot[(size_t)v] += x;
                else if (kind == GEN_MIN) tot[(size_t)v] = std::min(tot[(size_t)v], x);
                else tot[(size_t)v] = std::max(tot[(size_t)v], x);
            }
            // group order: first joined row, ranks in order (shards are contiguous row ranges)
            for (int g = 0; g < G; g++) if (f[g] != 0x7f7f7f7f7f7f7f7fLL && first[(size_t)g] == INT64_MAX) first[(size_t)g] = ((i64)r << 48) + f[g];
        }
        const int nacc = aprm.nacc;
        std::vector<int> order;
        for (int g = 0; g < G; g++) if (tot[(size_t)g * P] > 0) order.push_back(g);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return first[(size_t)a] < first[(size_t)b]; });
        res->nrows = (i64)order.size();
        for (auto &o : agg_outs) {
            ResCol col;
            if (o.first == 0) {
                const Column &kc = tab(key_side[o.second])->cols[(size_t)key_col[o.second]];
                col.type = kc.type;
                if (kc.type == PG_T_DICT8) col.dict = kc.dict;
                for (int g : order) {
                    const int id = (nkeys == 2) ? (o.second == 0 ? g / aprm.n1 : g % aprm.n1) : g;
                    col.push<uint8_t>(vals[o.second][(size_t)id]);
                }
            } else {
                const AggExpr &a = aggs[(size_t)o.second];
                const int pl = plane[(size_t)o.second];
                const bool is_int = agg_is_int[(size_t)o.second];
                col.width = a.width;
                col.scale = a.scale;
                if (a.fn == PG_AGG_COUNT) col.type = PG_T_HUGEINT;
                else if (a.fn == PG_AGG_AVG) col.type = is_int ? PG_T_FLOAT64 : PG_T_DECIMAL128;
                else col.type = is_int ? PG_T_HUGEINT : PG_T_DECIMAL128;
                size_t nrow = 0;
                for (int g : order) {
                    i128 v = tot[(size_t)g * P + (size_t)pl], cnt = tot[(size_t)g * P];
                    if (pl > 0) {
                        // NULL arguments were skipped; no valid input at all => NULL, count(x) included (CountOp.Finalize,
                        // function_aggr.go:949-960 -- the reference's q13.txt prints NULL for customers without orders)
                        cnt = tot[(size_t)g * P + (size_t)(nacc + pl)];
                        if (cnt == 0) { col.push_null(nrow++, (size_t)type_size(col.type)); continue; }
                    }
                    col.mark_valid();
                    nrow++;
                    if (a.fn == PG_AGG_COUNT || (a.fn != PG_AGG_AVG && is_int)) {
                        pg_hugeint h;
                        h.lower = (u64)v;
                        h.upper = (i64)(v >> 64);
                        col.push(h);
                    } else if (a.fn == PG_AGG_AVG && is_int) {
                        const i128 mag = v < 0 ? -v : v;
                        if (mag >= ((i128)1 << 53)) PG_FAIL(PG_EOVERFLOW, "avg(INT): sum not exact in float64");
                        col.push((double)(i64)v / (double)(i64)cnt);
                    } else {
                        HDec d;
                        if (hd_digits((u128)(v < 0 ? -v : v)) > HD_MAXPREC || !hd_from_i128(v, plane_scale[(size_t)pl], &d))
                            PG_FAIL(PG_EOVERFLOW, "decimal aggregate exceeds 19 significant digits (order-dependent rounding regime)");
                        if (a.fn == PG_AGG_AVG) {
                            HDec nd, qd;
                            hd_from_i128(cnt, 0, &nd);
                            if (!hd_quo(d, nd, &qd)) PG_FAIL(PG_EOVERFLOW, "avg: decimal division failed");
                            d = qd;
                        }
                        pg_decimal pd;
                        pd.coef = d.coef; pd.scale = d.scale; pd.neg = d.neg ? 1u : 0u;
                        col.push(pd);
                    }
                }
            }
            res->cols.push_back(col);
        }
        res->stats.kernel_ms = ev_all.ms();
        res->stats.main_kernel_ms = res->stats.kernel_ms;
        res->stats.rows_scanned = p.probe_rows;
        res->stats.kernel_launches = launches + 2;
        res->stats.aux[0] = p.probe_rows;
        res->stats.aux[1] = total;
        res->stats.aux[6] = res->nrows;
        return PG_OK;
    }

    static int check_err(int err)
    {
        // the reference panics on these and the query fails (executor_bench.go:184-189 recover)
        if (err == RV_ERR_OVERFLOW) PG_FAIL(PG_EOVERFLOW, "row expression: decimal overflow (integer part needs more than 19 digits)");
        if (err == RV_ERR_DIVZERO) PG_FAIL(PG_EOVERFLOW, "row expression: division by zero");
        if (err == RV_ERR_FLOAT) PG_FAIL(PG_EUNSUPPORTED, "row expression: value outside the exactly reproducible float32 cast range");
        return PG_OK;
    }
};

}  // namespace

int build_rows(pg_plan *plan, std::unique_ptr<Pipeline> *out, const Node *aggn)
{
    std::unique_ptr<RowsPipeline> p(new RowsPipeline());
    p->plan = plan;
    const Node *n = aggn ? &aggn->children[0] : &plan->root;
    const Node *project = nullptr;
    if (n->op == PG_OP_PROJECT) { project = n; n = &n->children[0]; }
    std::vector<const Expr *> above;                 // filters above the join (or above the scan)
    while (n->op == PG_OP_FILTER) { for (auto &f : n->filters) above.push_back(&f); n = &n->children[0]; }
    const Node *top = n;
    auto leaf_scan = [&](const Node *x, std::vector<const Expr *> *fl) -> const Node * {
        while (x->op == PG_OP_FILTER) { for (auto &f : x->filters) fl->push_back(&f); x = &x->children[0]; }
        if (x->op != PG_OP_SCAN) return nullptr;
        for (auto &f : x->filters) fl->push_back(&f);
        return x;
    };
    std::vector<const Expr *> pf, bf;
    const Node *pscan = nullptr, *bscan = nullptr;
    if (top->op == PG_OP_JOIN) {
        pscan = leaf_scan(&top->children[0], &pf);
        bscan = leaf_scan(&top->children[1], &bf);
        if (!pscan || !bscan) PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: both inputs must be (filtered) scans");
        if (top->jointype != PG_JOIN_INNER && top->jointype != PG_JOIN_LEFT && top->jointype != PG_JOIN_SEMI && top->jointype != PG_JOIN_ANTI &&
            top->jointype != PG_JOIN_MARK)
            PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: join type %d", top->jointype);
        p->slot[0] = pscan->slot;
        p->slot[1] = bscan->slot;
    } else if (top->op == PG_OP_SCAN) {
        pscan = top;
        for (auto &f : top->filters) pf.push_back(&f);
        p->slot[0] = pscan->slot;
    } else {
        PG_FAIL(PG_EUNSUPPORTED, "row-emitting pipeline over operator %d", top->op);
    }
    const pg_table *pt = p->tab(0), *bt = bscan ? p->tab(1) : nullptr;
    p->cc.tables[0] = pt;
    p->cc.tables[1] = bt;
    // several ranks: every rank joins ITS probe rows (its shard, or the whole replicated table) against the WHOLE build side --
    // a sharded build side would make "no match on this rank" (LEFT / ANTI / MARK, and SEMI duplicates) wrong
    if (ctx().world > 1 && bt && bt->dist != PG_DIST_REPLICATED)
        PG_FAIL(PG_EUNSUPPORTED, "row-emitting join against a sharded build side (replicate it)");
    if (bt && bt->nrows >= ((i64)1 << 31)) PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: build side of 2^31 rows or more");
    if (pt->nrows >= ((i64)1 << 31) - 1) PG_FAIL(PG_EUNSUPPORTED, "row-emitting pipeline over 2^31 rows or more");
    // scopes
    Resolver probe_scope = [&](int idx, Src *s) { if (idx < 0 || idx >= (int)pt->cols.size()) return false; s->side = 0; s->col = idx; s->mark = false; return true; };
    Resolver build_scope = [&](int idx, Src *s) { if (!bt || idx < 0 || idx >= (int)bt->cols.size()) return false; s->side = 1; s->col = idx; s->mark = false; return true; };
    Resolver top_scope = probe_scope;
    if (top->op == PG_OP_JOIN) {
        top_scope = [&](int idx, Src *s) {
            if (idx < 0 || idx >= (int)top->outs.size()) return false;
            const auto &o = top->outs[(size_t)idx];
            if (o.first == 2) { s->mark = true; return top->jointype == PG_JOIN_MARK; }
            if (o.first == 1 && top->jointype != PG_JOIN_INNER && top->jointype != PG_JOIN_LEFT) return false;
            return o.first == 0 ? probe_scope(o.second, s) : o.first == 1 ? build_scope(o.second, s) : false;
        };
    }
    RowsParams &prm = p->prm;
    prm.jointype = top->op == PG_OP_JOIN ? top->jointype : 0;
    prm.probe_rows = pt->nrows;
    prm.build_rows = bt ? bt->nrows : 0;
#define PG_ROWS_TRY(x) do { if (!(x)) PG_FAIL(PG_EUNSUPPORTED, "row-emitting pipeline: %s", p->cc.why.c_str()); } while (0)
    PG_ROWS_TRY(p->cc.compile_filters(pf, probe_scope, &prm.pf0, &prm.pf1));
    // build-side filters run with the build row on side 1
    PG_ROWS_TRY(p->cc.compile_filters(bf, build_scope, &prm.bf0, &prm.bf1));
    PG_ROWS_TRY(p->cc.compile_filters(above, top_scope, &prm.jf0, &prm.jf1));
    if (top->op == PG_OP_JOIN) {
        if (top->conds.empty() || top->conds.size() > 2) PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: 1 or 2 key columns");
        prm.nkey = (int)top->conds.size();
        for (int k = 0; k < prm.nkey; k++) {
            const Expr *pe = strip_value_preserving_casts(&top->conds[(size_t)k].first), *be = strip_value_preserving_casts(&top->conds[(size_t)k].second);
            Src ps, bs;
            if (pe->kind != PG_TK_COL || be->kind != PG_TK_COL || !probe_scope(pe->idx, &ps) || !build_scope(be->idx, &bs))
                PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: condition is not column = column");
            const Column &pc = pt->cols[(size_t)ps.col], &bc = bt->cols[(size_t)bs.col];
            if (!is_int_family(pc.type) || !is_int_family(bc.type)) PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: key columns must be integer / date / decimal");
            if (pc.type == PG_T_DECIMAL64 || bc.type == PG_T_DECIMAL64) {
                if (pc.type != bc.type || pc.scale != bc.scale) PG_FAIL(PG_EUNSUPPORTED, "row-emitting join: DECIMAL keys of different scales");
            }
            if (prm.nkey == 2) {
                auto fits32 = [](const Column &c) { return c.stats_ok && c.gmin() >= INT32_MIN && c.gmax() <= INT32_MAX; };
                if (!fits32(pc) || !fits32(bc)) PG_FAIL(PG_EUNSUPPORTED, "two-column join keys must both fit 32 bits");
            } else if (bc.gmin() <= HT_EMPTY && bc.gmax() >= HT_EMPTY) {
                PG_FAIL(PG_EUNSUPPORTED, "join key range contains the empty-slot sentinel");
            }
            prm.pkey[k] = p->cc.col_slot(0, ps.col);
            prm.bkey[k] = p->cc.col_slot(1, bs.col);
            if (prm.pkey[k] < 0 || prm.bkey[k] < 0) PG_FAIL(PG_EUNSUPPORTED, "row-emitting pipeline: too many columns referenced");
        }
    }
    if (aggn) {
        // ---- aggregate mode: group keys and aggregate arguments over the joined row ----
        if (top->op != PG_OP_JOIN) PG_FAIL(PG_EUNSUPPORTED, "aggregate over rows: expected a join below the aggregate");
        if (aggn->groups.size() > 2) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a join with expression arguments: more than two group keys");
        if (!aggn->having.empty()) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a join with expression arguments: HAVING");
        p->agg_mode = true;
        p->nkeys = (int)aggn->groups.size();
        std::vector<uint8_t> luts(512, 0);
        int dims[2] = {1, 1};
        VmAggParams &q = p->aprm;
        for (int k = 0; k < p->nkeys; k++) {
            const Expr *ge = strip_value_preserving_casts(&aggn->groups[(size_t)k]);
            Src s;
            if (ge->kind != PG_TK_COL || ge->side != 0 || !top_scope(ge->idx, &s) || s.mark) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a join: group key %d is not a column", k);
            const Column &col = p->tab(s.side)->cols[(size_t)s.col];
            if (!is_byte_family(col.type) || col.any_nulls()) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a join with expression arguments: group key %s is not a non-null dictionary / char column", col.name.c_str());
            if (s.side == 1 && top->jointype != PG_JOIN_INNER) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a join: a build-side group key needs an INNER join");
            p->key_side[k] = s.side;
            p->key_col[k] = s.col;
            const uint32_t *present = col.gpresent();
            for (int code = 0; code < 256; code++)
                if (present[code >> 5] & (1u << (code & 31))) { luts[(size_t)k * 256 + (size_t)code] = (uint8_t)p->vals[k].size(); p->vals[k].push_back((uint8_t)code); }
            if (p->vals[k].empty()) p->vals[k].push_back(0);
            dims[k] = (int)p->vals[k].size();
            (k == 0 ? q.key0 : q.key1) = (const uint8_t *)col.d_data;
            q.key_side[k] = s.side;
        }
        p->G = dims[0] * dims[1];
        if (p->G > 64) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a join with expression arguments: more than 64 dense groups");
        q.nkeys = p->nkeys;
        q.n1 = dims[1];
        q.ngroups = p->G;
        q.pred0 = q.pred1 = 0;                        // the filters were applied while the pairs were made
        q.row_base = 0;
        p->plane_kind = {GEN_SUM};
        p->plane_scale = {0};
        p->aggs = aggn->aggs;
        int nacc = 0;
        for (size_t i = 0; i < aggn->aggs.size(); i++) {
            const AggExpr &a = aggn->aggs[i];
            if (a.fn == PG_AGG_COUNT && a.star) { p->plane.push_back(0); p->agg_is_int.push_back(true); continue; }
            const int kind = a.fn == PG_AGG_COUNT ? GEN_COUNTV : a.fn == PG_AGG_MIN ? GEN_MIN : a.fn == PG_AGG_MAX ? GEN_MAX : GEN_SUM;
            if (a.fn != PG_AGG_COUNT && a.fn != PG_AGG_MIN && a.fn != PG_AGG_MAX && a.fn != PG_AGG_SUM && a.fn != PG_AGG_AVG) PG_FAIL(PG_EUNSUPPORTED, "aggregate function %d", a.fn);
            if (nacc >= GEN_MAXACC) PG_FAIL(PG_EUNSUPPORTED, "more than 8 aggregate arguments");
            int k = 0;
            q.a0[nacc] = p->cc.ncode;
            PG_ROWS_TRY(p->cc.compile(a.arg, top_scope, &k));
            q.a1[nacc] = p->cc.ncode;
            const int sb = p->cc.scale_bound(a.arg, top_scope);
            if (kind != GEN_COUNTV) {
                if (k != RVK_INT && k != RVK_DEC) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a non-numeric expression");
                if (sb < 0 || sb > 18) PG_FAIL(PG_EUNSUPPORTED, "aggregate over a quotient (no fixed scale to accumulate at)");
            }
            q.kind[nacc] = kind;
            q.ascale[nacc] = kind == GEN_COUNTV ? 0 : sb;
            const bool is_int = a.ltype == PG_LT_HUGEINT || a.ltype == PG_LT_DOUBLE || a.ltype == PG_LT_INTEGER || a.ltype == PG_LT_BIGINT;
            if (is_int && kind != GEN_COUNTV && (k != RVK_INT || sb != 0)) PG_FAIL(PG_EUNSUPPORTED, "integer aggregate over a scaled value");
            if (!is_int && a.ltype != PG_LT_DECIMAL) PG_FAIL(PG_EUNSUPPORTED, "aggregate result type");
            if ((kind == GEN_MIN || kind == GEN_MAX) && is_int) PG_FAIL(PG_EUNSUPPORTED, "min/max are DECIMAL only in the reference");
            p->plane.push_back(nacc + 1);
            p->plane_kind.push_back(kind == GEN_COUNTV ? (int)GEN_SUM : kind);
            p->plane_scale.push_back(q.ascale[nacc]);
            p->agg_is_int.push_back(is_int);
            nacc++;
        }
        q.nacc = nacc;
        p->P = 1 + 2 * nacc;
        for (int a = 0; a < nacc; a++) { p->plane_kind.push_back(GEN_SUM); p->plane_scale.push_back(0); }
        for (auto &o : aggn->outs) {
            if ((o.first == 0 && (o.second < 0 || o.second >= p->nkeys)) || (o.first == 1 && (o.second < 0 || o.second >= (int)aggn->aggs.size())) || (o.first != 0 && o.first != 1))
                PG_FAIL(PG_EUNSUPPORTED, "bad aggregate output reference");
        }
        p->agg_outs = aggn->outs;
        p->NT = 256;
        while (p->NT >= 64 && (size_t)p->G * p->P * p->NT * 8 > (size_t)200 * 1024) p->NT /= 2;
        if (p->NT < 64) PG_FAIL(PG_EUNSUPPORTED, "group tables do not fit in shared memory");
        p->asmem = (size_t)p->G * p->P * p->NT * 8;
        const void *kern = p->NT == 256 ? (const void *)vm_scanagg_kernel<256> : p->NT == 128 ? (const void *)vm_scanagg_kernel<128> : (const void *)vm_scanagg_kernel<64>;
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->asmem));
        int per_sm = 1;
        PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, p->NT, p->asmem));
        p->agrid = std::max(per_sm, 1) * ctx().prop.multiProcessorCount;
        // int64 partial sums of a CTA stay exact while |value| x pairs per CTA < 2^62; the pair count is not known at
        // plan time, so bound it by what the emit pass can address (2^31 probe rows x matches is checked at run time:
        // the count pass refuses more than 2^40 pairs below) -- values above 2^22 per pair slot are refused instead
        q.absmax = (i64)1 << 40;
        PG_TRY(p->d_part.alloc(sizeof(i64) * (size_t)p->agrid * (size_t)p->G * (size_t)p->P));
        PG_TRY(p->d_final.alloc(p->rank_bytes()));
        PG_TRY(p->d_gather.alloc(p->rank_bytes() * (size_t)std::max(ctx().world, 1)));
        PG_TRY(p->h_final.alloc(p->rank_bytes() * (size_t)std::max(ctx().world, 1)));
        PG_TRY(p->d_luts.alloc(512));
        PG_TRY(p->d_kinds.alloc(sizeof(int) * (size_t)p->G * (size_t)p->P));
        std::vector<int> kinds((size_t)p->G * (size_t)p->P);
        for (int v = 0; v < p->G * p->P; v++) kinds[(size_t)v] = p->plane_kind[(size_t)(v % p->P)];
        PG_CUDA(cudaMemcpyAsync(p->d_luts.p, luts.data(), 512, cudaMemcpyHostToDevice, ctx().stream));
        PG_CUDA(cudaMemcpyAsync(p->d_kinds.p, kinds.data(), sizeof(int) * kinds.size(), cudaMemcpyHostToDevice, ctx().stream));
    }
    // outputs
    std::vector<Expr> owned;
    size_t nout = aggn ? 0 : project ? project->exprs.size() : top->op == PG_OP_JOIN ? top->outs.size() : pt->cols.size();
    if ((nout == 0 && !aggn) || nout > RV_MAXOUT) PG_FAIL(PG_EUNSUPPORTED, "row-emitting pipeline: %zu output columns", nout);
    owned.resize(nout);
    prm.nout = (int)nout;
    for (size_t j = 0; j < nout; j++) {
        const Expr *e;
        if (project) e = &project->exprs[j];
        else { owned[j].kind = PG_TK_COL; owned[j].side = 0; owned[j].idx = (int)j; e = &owned[j]; }
        RowsPipeline::Out o;
        const Expr *bare = strip_value_preserving_casts(e);
        Src s;
        if (bare->kind == PG_TK_COL && bare->side == 0 && top_scope(bare->idx, &s) && !s.mark) {
            // a column carried through: keeps its column type (and dictionary / host-resident strings)
            const Column &c = p->tab(s.side)->cols[(size_t)s.col];
            o.type = c.type; o.width = c.width; o.scale = c.scale; o.side = s.side; o.col = s.col;
            if (c.type == PG_T_VARCHAR) {
                prm.okind[j] = s.side ? RO_ROW1 : RO_ROW0;
                prm.o0[j] = prm.o1[j] = 0;
                p->outs.push_back(o);
                continue;
            }
            int k;
            prm.o0[j] = p->cc.ncode;
            PG_ROWS_TRY(p->cc.compile(*bare, top_scope, &k));
            prm.o1[j] = p->cc.ncode;
            prm.okind[j] = type_size(c.type) == 4 ? RO_I32 : type_size(c.type) == 1 ? RO_U8 : RO_I64;
            p->outs.push_back(o);
            continue;
        }
        int k;
        prm.o0[j] = p->cc.ncode;
        PG_ROWS_TRY(p->cc.compile(*e, top_scope, &k));
        prm.o1[j] = p->cc.ncode;
        switch (k) {
        case RVK_BOOL: o.type = PG_T_BOOL; prm.okind[j] = RO_BOOL; break;
        case RVK_INT:
            if (e->ltype == PG_LT_DATE) { o.type = PG_T_DATE32; prm.okind[j] = RO_I32; }
            else if (e->ltype == PG_LT_INTEGER) { o.type = PG_T_INT32; prm.okind[j] = RO_I32; }
            else { o.type = PG_T_INT64; prm.okind[j] = RO_I64; }
            break;
        case RVK_DEC: o.type = PG_T_DECIMAL128; o.width = e->width; o.scale = e->scale; prm.okind[j] = RO_DEC; break;
        default: PG_FAIL(PG_EUNSUPPORTED, "row-emitting pipeline: output column %zu has a type that is not returned (FLOAT / computed string)", j);
        }
        p->outs.push_back(o);
    }
#undef PG_ROWS_TRY
    PG_TRY(p->d_code.alloc(sizeof(RvCode)));
    PG_TRY(p->d_err.alloc(4));
    PG_CUDA(cudaMemcpyAsync(p->d_code.p, &p->cc.code, sizeof(RvCode), cudaMemcpyHostToDevice, ctx().stream));
    PG_CUDA(cudaStreamSynchronize(ctx().stream));
    static const char *jn[] = {"", "INNER", "SEMI", "ANTI", "MARK", "LEFT", "ANTI-MARK"};
    char b[320];
    if (aggn)
        snprintf(b, sizeof b, "JoinAgg[expression programs] %s join %s x %s (%d key column%s, bucketized hash table) -> count / scan / emit pairs -> vm_scanagg_kernel over the pairs, groups=%d accumulators=%d instructions=%d",
                 jn[top->jointype], pt->name.c_str(), bt->name.c_str(), prm.nkey, prm.nkey > 1 ? "s" : "", p->G, p->aprm.nacc, p->cc.ncode);
    else if (top->op == PG_OP_JOIN)
        snprintf(b, sizeof b, "Rows[%s%s join %s x %s (%d key column%s, bucketized hash table) -> count / scan / emit pairs -> project %zu columns] instructions=%d",
                 project ? "project <- " : "", jn[top->jointype], pt->name.c_str(), bt->name.c_str(), prm.nkey, prm.nkey > 1 ? "s" : "", nout, p->cc.ncode);
    else
        snprintf(b, sizeof b, "Rows[%sfilter scan(%s) -> count / scan / emit -> project %zu columns] instructions=%d", project ? "project <- " : "", pt->name.c_str(), nout, p->cc.ncode);
    p->explain = b;
    *out = std::move(p);
    return PG_OK;
}

}  // namespace pg
