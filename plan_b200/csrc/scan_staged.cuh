// scan_staged.cuh -- the scan-aggregate kernels over NARROW physical columns (every referenced column stored in
// <= 4 bytes), fed by the bulk-copy tile ring of stage.cuh.
//
// Reference path replaced (as scanagg.cuh): scan filter ExprExec.executeSelect pkg/compute/expr_exec.go:342-530,
// projection executeExprs :85-340, decimal ops function_operator_binary.go:134-191, group lookup
// FindOrCreateGroups aggregate_hash.go:201-391, state update function_aggr.go:770-1161 -- one pass, every column
// byte read once.
//
// Why a second family next to scanagg.cuh's register-staged kernels: at 8-11 stored bytes per row the HBM
// roofline allows ~2 rows / clk / SM, i.e. a budget of ~60 issued instructions per row over the whole SM.  The
// register-staged kernels spend that on address arithmetic and sub-word loads (ncu, r2: 52 instructions / row for
// Q6, issue-bound at 0.49 of the roofline; Q1 0.25).  Here the producer lane issues one bulk copy per column per
// 2048-row tile and the consumers' inner loops hold nothing but the predicate and the arithmetic.
#pragma once
#include "stage.cuh"

namespace pg {

template <int J> struct IC { static constexpr int v = J; };
#define PG_FOR4(f) do { f(IC<0>()); f(IC<1>()); f(IC<2>()); f(IC<3>()); } while (0)

// does this CTA's LAST tile hold pad rows?  (only the table's last tile can)
__device__ __forceinline__ bool last_tile_is_partial(const TileSeq &seq, i64 ntiles, i64 nrows, int tile_rows)
{
    return seq.count > 0 && seq.first + (seq.count - 1) * seq.step == ntiles - 1 && nrows % tile_rows != 0;
}

// ------------------------------------------------------------------------------
// sumprod: ungrouped sum(fa * fb) under inclusive range predicates (TPC-H Q6).
// Roles: 0 = fa, 1 = fb, 2 = pa, 3 = pb (width 0: absent).  A range is  (v - lo) <=u span  on the STORED value.
// ------------------------------------------------------------------------------
struct SumProdSParams {
    StageDesc st;                     // the distinct physical columns
    int roff[4], rpw[4];              // role -> stage offset / width
    unsigned a_lo, a_span, b_lo, b_span, x_lo, x_span, y_lo, y_span;
    int xbase, ybase;                 // logical = stored + base
    i64 nrows;
};

template <int WFA, int WFB, int WPA, int WPB, bool XR, bool YR, int QPT, bool MASK>
__device__ __forceinline__ void sumprod_tile(const SumProdSParams &p, const StageRing &ring, StageCursor &cur, int warp, int lane,
                                             int rows_in_tile, i64 &sum, unsigned &cnt)
{
    const char *stg = stage_acquire(ring, p.st, cur);
    Quad<WFA> x[QPT];
    Quad<WFB> y[QPT];
    Quad<WPA> a[QPT];
    Quad<WPB> b[QPT];
#pragma unroll
    for (int q = 0; q < QPT; q++) {
        const int qrow = (warp * QPT + q) * 128 + lane * 4;
        x[q].load(stg + p.roff[0], qrow, p.rpw[0]);
        y[q].load(stg + p.roff[1], qrow, p.rpw[1]);
        a[q].load(stg + p.roff[2], qrow, p.rpw[2]);
        b[q].load(stg + p.roff[3], qrow, p.rpw[3]);
    }
    stage_release(ring, p.st, cur);
#pragma unroll
    for (int q = 0; q < QPT; q++) {
        const int rem = rows_in_tile - ((warp * QPT + q) * 128 + lane * 4);
        auto row = [&](auto jc) {
            constexpr int J = decltype(jc)::v;
            bool ok = !MASK || J < rem;
            if (WPA != 0) ok = ok && (a[q].template get<J>() - p.a_lo) <= p.a_span;
            if (WPB != 0) ok = ok && (b[q].template get<J>() - p.b_lo) <= p.b_span;
            const unsigned xs = x[q].template get<J>(), ys = y[q].template get<J>();
            if (XR) ok = ok && (xs - p.x_lo) <= p.x_span;
            if (YR) ok = ok && (ys - p.y_lo) <= p.y_span;
            const unsigned xl = xs + (unsigned)p.xbase;     // logical values are non-negative (host-checked)
            const unsigned ym = ok ? ys + (unsigned)p.ybase : 0u;
            sum += (i64)((u64)xl * (u64)ym);               // one IMAD.WIDE.U32
            cnt += ok ? 1u : 0u;
        };
        PG_FOR4(row);
    }
}

template <int WFA, int WFB, int WPA, int WPB, bool XR, bool YR, int QPT>
__global__ void __launch_bounds__(ST_THREADS)
sumprod_staged_kernel(const SumProdSParams p, i64 *__restrict__ partials /* [grid][2] = {sum, count} */)
{
    extern __shared__ __align__(128) unsigned char st_smem[];
    __shared__ i64 s_sum[ST_CONS_WARPS], s_cnt[ST_CONS_WARPS];
    const StageDesc &d = p.st;
    const StageRing ring = stage_ring_init(st_smem, d);
    const i64 ntiles = (p.nrows + d.tile_rows - 1) / d.tile_rows;
    const TileSeq seq = tile_seq(ntiles, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == ST_CONS_WARPS) {
        stage_produce(ring, d, seq);
    } else {
        i64 sum = 0, cnt64 = 0;
        StageCursor cur = {0, 0};
        const bool partial = last_tile_is_partial(seq, ntiles, p.nrows, d.tile_rows);
        const int nwhole = (int)seq.count - (partial ? 1 : 0);
        for (int k = 0; k < nwhole; k++) {
            unsigned cnt = 0;
            sumprod_tile<WFA, WFB, WPA, WPB, XR, YR, QPT, false>(p, ring, cur, warp, lane, 0, sum, cnt);
            cnt64 += cnt;
        }
        if (partial) {
            unsigned cnt = 0;
            sumprod_tile<WFA, WFB, WPA, WPB, XR, YR, QPT, true>(p, ring, cur, warp, lane, (int)(p.nrows - (ntiles - 1) * d.tile_rows), sum, cnt);
            cnt64 += cnt;
        }
        sum = warp_sum(sum);
        cnt64 = warp_sum(cnt64);
        if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt64; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        i64 s = 0, c = 0;
#pragma unroll
        for (int w = 0; w < ST_CONS_WARPS; w++) { s += s_sum[w]; c += s_cnt[w]; }
        partials[2 * blockIdx.x] = s;
        partials[2 * blockIdx.x + 1] = c;
    }
}

// ------------------------------------------------------------------------------
// lowcard chain (TPC-H Q1): GROUP BY <= 2 byte-coded keys (<= 8 dense groups), accumulators
//   [0] count(*)  [1] sum(q)  [2] sum(A)  [3] sum(A*(c1+s1*B))  [4] sum(A*(c1+s1*B)*(c2+s2*C))  [5] sum(B)
// over rows passing one inclusive range.  Output format identical to lowcard_chain_kernel (scanagg.cuh).
//
// Group state: every consumer thread owns a private table in shared memory, updated for EVERY row with plain
// LDS / STS (no atomics, no data-dependent branch: the cost does not depend on how the groups are clustered).
// What bounds the kernel is shared-memory bandwidth, so an entry is packed into THREE 64-bit words, each
// holding a wide sum in its low bits and a small one in its high (>= 32) bits:
//     w0 = sum t3 | count << sh0      w1 = sum t2 | sum q << sh1      w2 = sum A | sum B << sh2
// (all over non-negative values; the host proves from the statistics and the rows a thread can see that no
// field overflows into its neighbour).  24 bytes read + 24 written per row instead of 2 x 48.
// The table slot of a row is a multiplicative hash of its two key bytes, chosen by the host to be injective
// on the key combinations that can occur:  slot = ((k0 | k1 << 8) * 0x10001 * M) >> 29.
// Roles: 0 pred, 1 key0, 2 key1, 3 q, 4 A, 5 B, 6 C   (width 0: absent, reads as 0).
// ------------------------------------------------------------------------------
struct LowcardSParams {
    StageDesc st;                     // the distinct physical columns
    int roff[7], rpw[7];              // role -> stage offset / width
    unsigned p_lo, p_span;
    int abase;                        // logical A = stored + abase
    int f1c, f1s, f2c, f2s;           // factors on the STORED values of B and C
    unsigned hashM;
    unsigned mul0, mul1, mul2;        // 1 << (sh - 32): what one unit of the small field adds to the high word
    int sh0, sh1, sh2;
    unsigned char slot_group[8];      // table slot -> dense group id, 0xff: unused
    int ngroups;
    int contig;
    i64 qbase, Abase, Bbase;          // stored -> logical for the sums of q, A, B
    i64 nrows;
};

__host__ __device__ inline unsigned lc_slot_of(unsigned k0, unsigned k1, bool has_key1, unsigned M)
{
    const unsigned kk = has_key1 ? (k0 | (k1 << 8)) * 0x00010001u : k0 * 0x01010101u;
    return (kk * M) >> 29;
}

constexpr int LCS_TBL_BYTES = 8 * 3 * 8 * ST_CONS_THREADS;       // 8 slots x 3 words per consumer thread

template <int WP, int WQ, int WA, int WB, int WC, bool HAS_KEY1, int QPT, bool MASK>
__device__ __forceinline__ void lowcard_tile(const LowcardSParams &p, const StageRing &ring, StageCursor &cur, int warp, int lane,
                                             int rows_in_tile, char *my01, char *my2)
{
    const char *stg = stage_acquire(ring, p.st, cur);
    Quad<WP> dv[QPT];
    Quad<WQ> qv[QPT];
    Quad<WA> av[QPT];
    Quad<WB> bv[QPT];
    Quad<WC> cv[QPT];
    unsigned k0[QPT], k1[QPT];
#pragma unroll
    for (int q = 0; q < QPT; q++) {
        const int qrow = (warp * QPT + q) * 128 + lane * 4;
        dv[q].load(stg + p.roff[0], qrow, p.rpw[0]);
        k0[q] = *(const unsigned *)(stg + p.roff[1] + qrow);
        k1[q] = HAS_KEY1 ? *(const unsigned *)(stg + p.roff[2] + qrow) : 0u;
        qv[q].load(stg + p.roff[3], qrow, p.rpw[3]);
        av[q].load(stg + p.roff[4], qrow, p.rpw[4]);
        bv[q].load(stg + p.roff[5], qrow, p.rpw[5]);
        cv[q].load(stg + p.roff[6], qrow, p.rpw[6]);
    }
    stage_release(ring, p.st, cur);
#pragma unroll
    for (int q = 0; q < QPT; q++) {
        const int rem = rows_in_tile - ((warp * QPT + q) * 128 + lane * 4);
        auto row = [&](auto jc) {
            constexpr int J = decltype(jc)::v;
            const bool ok = (!MASK || J < rem) && (dv[q].template get<J>() - p.p_lo) <= p.p_span;
            if (ok) {
                // both key bytes twice (k0 k1 k0 k1) = (k0 | k1 << 8) * 0x10001: the hash needs no masking
                const unsigned kk = HAS_KEY1 ? __byte_perm(k0[q], k1[q], (unsigned)(J | ((4 + J) << 4) | (J << 8) | ((4 + J) << 12)))
                                             : __byte_perm(k0[q], 0u, (unsigned)(J * 0x1111));
                const unsigned slot = (kk * p.hashM) >> 29;
                ulonglong2 *e01 = (ulonglong2 *)(my01 + slot * (16 * ST_CONS_THREADS));
                u64 *e2 = (u64 *)(my2 + slot * (8 * ST_CONS_THREADS));
                ulonglong2 w01 = *e01;
                u64 w2 = *e2;
                const unsigned as = av[q].template get<J>(), bs = bv[q].template get<J>();
                const unsigned t2 = (as + (unsigned)p.abase) * (unsigned)(p.f1c + p.f1s * (int)bs);      // proven < 2^32
                const unsigned f2 = (unsigned)(p.f2c + p.f2s * (int)cv[q].template get<J>());
                w01.x += (u64)t2 * f2 + ((u64)p.mul0 << 32);
                w01.y += (u64)t2 + ((u64)(qv[q].template get<J>() * p.mul1) << 32);
                w2 += (u64)as + ((u64)(bs * p.mul2) << 32);
                *e01 = w01;
                *e2 = w2;
            }
        };
        PG_FOR4(row);
    }
}

template <int WP, int WQ, int WA, int WB, int WC, bool HAS_KEY1, int QPT>
__global__ void __launch_bounds__(ST_THREADS, 2)
lowcard_staged_kernel(const LowcardSParams p, i64 *__restrict__ partials /* [grid][G*6] */)
{
    extern __shared__ __align__(128) unsigned char st_smem[];
    __shared__ i64 s_tot[8 * 6];
    const StageDesc &d = p.st;
    const StageRing ring = stage_ring_init(st_smem, d);
    // tables behind the stage ring: [slot][thread] pairs {w0, w1} (16 B), then [slot][thread] w2 (8 B)
    ulonglong2 *t01 = (ulonglong2 *)(st_smem + ST_HDR + (size_t)d.nstage * d.stage_bytes);
    u64 *t2p = (u64 *)(t01 + 8 * ST_CONS_THREADS);
    for (int i = threadIdx.x; i < 8 * ST_CONS_THREADS; i += ST_THREADS) { t01[i] = make_ulonglong2(0, 0); t2p[i] = 0; }
    if (threadIdx.x < 48) s_tot[threadIdx.x] = 0;
    __syncthreads();
    const i64 ntiles = (p.nrows + d.tile_rows - 1) / d.tile_rows;
    const TileSeq seq = tile_seq(ntiles, p.contig);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == ST_CONS_WARPS) {
        stage_produce(ring, d, seq);
    } else {
        ulonglong2 *my01 = t01 + threadIdx.x;
        u64 *my2 = t2p + threadIdx.x;
        StageCursor cur = {0, 0};
        const bool partial = last_tile_is_partial(seq, ntiles, p.nrows, d.tile_rows);
        const int nwhole = (int)seq.count - (partial ? 1 : 0);
        for (int k = 0; k < nwhole; k++)
            lowcard_tile<WP, WQ, WA, WB, WC, HAS_KEY1, QPT, false>(p, ring, cur, warp, lane, 0, (char *)my01, (char *)my2);
        if (partial)
            lowcard_tile<WP, WQ, WA, WB, WC, HAS_KEY1, QPT, true>(p, ring, cur, warp, lane, (int)(p.nrows - (ntiles - 1) * d.tile_rows), (char *)my01, (char *)my2);
        // this thread's fields -> block totals (shared-memory atomics: once per thread, not per row)
        for (int s = 0; s < 8; s++) {
            const int g = p.slot_group[s];
            if (g == 0xff) continue;                               // warp-uniform
            const ulonglong2 w01 = my01[s * ST_CONS_THREADS];
            const u64 w2 = my2[s * ST_CONS_THREADS];
            i64 f[6];
            f[0] = (i64)(w01.x >> p.sh0);                          // count
            f[1] = (i64)(w01.y >> p.sh1);                          // sum q (stored)
            f[2] = (i64)(w2 & (((u64)1 << p.sh2) - 1));            // sum A (stored)
            f[3] = (i64)(w01.y & (((u64)1 << p.sh1) - 1));         // sum t2
            f[4] = (i64)(w01.x & (((u64)1 << p.sh0) - 1));         // sum t3
            f[5] = (i64)(w2 >> p.sh2);                             // sum B (stored)
#pragma unroll
            for (int kx = 0; kx < 6; kx++) {
                const i64 v = warp_sum(f[kx]);
                if (lane == 0) atomicAdd((unsigned long long *)&s_tot[g * 6 + kx], (unsigned long long)v);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < p.ngroups * 6) {
        const int g = threadIdx.x / 6, kx = threadIdx.x % 6;
        i64 s = s_tot[threadIdx.x];
        const i64 n = s_tot[g * 6];
        if (kx == 1) s += n * p.qbase;                              // stored -> logical sums
        else if (kx == 2) s += n * p.Abase;
        else if (kx == 5) s += n * p.Bbase;
        partials[(i64)blockIdx.x * (p.ngroups * 6) + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------
// Ordered-rounding summaries (scanagg.cuh, "sequential-rounding emulation") over narrow columns, fed by the tile ring.
// One CTA per work item (a 1024-row tile, or ORD_CHUNK of them); tile = 8 warps x 128 rows, a lane simulates its 4
// consecutive rows for both entering parities, a warp composes its 128-row slice with ballots (see ord_jobs_fast_kernel),
// slices are composed in row order by one thread per item.  Digits and parities without 64-bit division:
// x = 10 q + d, d from residues, parity(q) = parity((x - d) / 2), sum q = (sum x - sum d) / 10.
// Roles: 0 pred, 1 key0, 2 key1, 3 A, 4 B, 5 C.
// ------------------------------------------------------------------------------
struct OrdStagedParams {
    StageDesc st;
    int roff[6], rpw[6];
    unsigned p_lo, p_span;
    int abase, bbase, f1c, f1s, f2c, f2s;
    int has_key1, n1;
    unsigned char code0[LC_MAXG], code1[LC_MAXG];
    i64 nrows;
};

template <int WP, int WA, int WB, int WC>
__global__ void __launch_bounds__(ST_THREADS)
ord_jobs_staged_kernel(const OrdStagedParams p, const OrdJob *__restrict__ jobs, const int *__restrict__ njobs, OrdSummary *__restrict__ out,
                       i64 job_stride, i64 ntiles)
{
    extern __shared__ __align__(128) unsigned char st_smem[];
    __shared__ int4 s_slice[2][ORD_CHUNK][ST_CONS_WARPS];          // per (item parity, tile, warp): {D0, D1, G0, G1}
    __shared__ unsigned long long s_acc[2][3];                      // per item parity: sum x, sum digits, base carries
    const int nj = *njobs < ORD_MAXJOBS ? *njobs : ORD_MAXJOBS;
    if (nj == 0) return;
    const StageDesc &d = p.st;
    const StageRing ring = stage_ring_init(st_smem, d);
    if (threadIdx.x < 6) s_acc[threadIdx.x / 3][threadIdx.x % 3] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == ST_CONS_WARPS) {                 // producer: the tiles of this CTA's items, in the consumers' order
        if (lane != 0) return;
        int s = 0;
        unsigned ph = 0, tx = 0;
        for (int c = 0; c < d.ncol; c++) tx += (unsigned)(d.tile_rows * d.pw[c]);
        for (int e = 0; e < nj; e++) {
            const OrdJob job = jobs[e];
            if (job.role == 0) continue;
            const i64 ntile1 = job.te - job.tb, nch = (ntiles - job.te + ORD_CHUNK - 1) / ORD_CHUNK;
            for (i64 w = blockIdx.x; w < ntile1 + nch; w += gridDim.x) {
                i64 t0, t1;
                if (w < ntile1) { t0 = job.tb + w; t1 = t0 + 1; }
                else { t0 = job.te + (w - ntile1) * ORD_CHUNK; t1 = t0 + ORD_CHUNK < ntiles ? t0 + ORD_CHUNK : ntiles; }
                for (i64 tile = t0; tile < t1; tile++) {
                    mbar_wait(ring.empty0 + 8 * s, ph ^ 1u);
                    const unsigned bar = ring.full0 + 8 * s;
                    mbar_expect_tx(bar, tx);
                    const unsigned dst = ring.data0 + (unsigned)s * (unsigned)d.stage_bytes;
                    for (int c = 0; c < d.ncol; c++) {
                        const unsigned bytes = (unsigned)(d.tile_rows * d.pw[c]);
                        bulk_g2s(dst + (unsigned)d.off[c], d.src[c] + tile * (i64)bytes, bytes, bar);
                    }
                    if (++s == d.nstage) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }
    StageCursor cur = {0, 0};
    const unsigned lt = (1u << lane) - 1u;
    int par = 0;                                  // item parity: which half of the shared scratch this item uses
    for (int e = 0; e < nj; e++) {
        const OrdJob job = jobs[e];
        if (job.role == 0) continue;
        const int slot = job.s;
        const int i0 = p.has_key1 ? job.g / p.n1 : job.g, i1 = p.has_key1 ? job.g % p.n1 : 0;
        const unsigned target = (unsigned)p.code0[i0] | (p.has_key1 ? (unsigned)p.code1[i1] << 8 : 0u);
        const i64 ntile1 = job.te - job.tb, nch = (ntiles - job.te + ORD_CHUNK - 1) / ORD_CHUNK;
        for (i64 w = blockIdx.x; w < ntile1 + nch; w += gridDim.x) {
            i64 t0, t1;
            if (w < ntile1) { t0 = job.tb + w; t1 = t0 + 1; }
            else { t0 = job.te + (w - ntile1) * ORD_CHUNK; t1 = t0 + ORD_CHUNK < ntiles ? t0 + ORD_CHUNK : ntiles; }
            u64 sx = 0;
            unsigned sd = 0, cbase = 0;
            for (i64 tile = t0; tile < t1; tile++) {
                const char *stg = stage_acquire(ring, d, cur);
                Quad<WP> dv;
                Quad<WA> av;
                Quad<WB> bv;
                Quad<WC> cv;
                const int qrow = warp * 128 + lane * 4;
                dv.load(stg + p.roff[0], qrow, p.rpw[0]);
                const unsigned k0 = *(const unsigned *)(stg + p.roff[1] + qrow);
                const unsigned k1 = p.has_key1 ? *(const unsigned *)(stg + p.roff[2] + qrow) : 0u;
                av.load(stg + p.roff[3], qrow, p.rpw[3]);
                bv.load(stg + p.roff[4], qrow, p.rpw[4]);
                cv.load(stg + p.roff[5], qrow, p.rpw[5]);
                stage_release(ring, d, cur);
                const i64 rem = p.nrows - (tile * 1024 + qrow);
                unsigned c0 = 0, c1 = 0, q0 = 0, q1 = 1;
                auto row = [&](auto jc) {
                    constexpr int J = decltype(jc)::v;
                    const unsigned kk = __byte_perm(k0, k1, (unsigned)(J | ((4 + J) << 4) | 0x4400)) & 0xffffu;
                    const bool ok = J < rem && (dv.template get<J>() - p.p_lo) <= p.p_span && kk == target;
                    if (ok) {
                        const unsigned bs = bv.template get<J>();
                        const unsigned al = av.template get<J>() + (unsigned)p.abase;
                        const unsigned t2 = al * (unsigned)(p.f1c + p.f1s * (int)bs);
                        unsigned xlo, dg;
                        if (slot == 4) {
                            const unsigned f2 = (unsigned)(p.f2c + p.f2s * (int)cv.template get<J>());
                            sx += (u64)t2 * f2;
                            xlo = t2 * f2;
                            dg = ((t2 % 10u) * (f2 % 10u)) % 10u;
                        } else {
                            xlo = slot == 3 ? t2 : slot == 2 ? al : bs + (unsigned)p.bbase;
                            sx += xlo;
                            dg = xlo % 10u;
                        }
                        const unsigned t = ((xlo - dg) >> 1) & 1u;
                        sd += dg;
                        if (dg == 5u) {
                            c0 += q0 ^ t;
                            c1 += q1 ^ t;
                            q0 = 0;
                            q1 = 0;
                        } else {
                            const unsigned cy = dg > 5u ? 1u : 0u;
                            c0 += cy;
                            c1 += cy;
                            q0 ^= t ^ cy;
                            q1 ^= t ^ cy;
                        }
                    }
                };
                PG_FOR4(row);
                // this warp's 128-row slice: parity entering each lane for the slice entered even (E0) / odd (E1)
                const bool reset = q0 == q1;
                const unsigned R = __ballot_sync(0xffffffffu, reset), F = __ballot_sync(0xffffffffu, q0 != 0u);
                unsigned E0, E1;
                const unsigned belowR = R & lt;
                if (belowR) {
                    const int m = 31 - __clz(belowR);
                    E0 = E1 = ((F >> m) & 1u) ^ (__popc(F & ~R & lt & ~((2u << m) - 1u)) & 1u);
                } else {
                    E0 = __popc(F & lt) & 1u;
                    E1 = E0 ^ 1u;
                }
                // carries = c0 (order independent, kept per lane) + the part that depends on the entering parity
                cbase += c0;
                const bool up = c1 > c0, dn = c1 < c0;                 // |c1 - c0| <= 1
                const int D0 = __popc(__ballot_sync(0xffffffffu, E0 && up)) - __popc(__ballot_sync(0xffffffffu, E0 && dn));
                const int D1 = __popc(__ballot_sync(0xffffffffu, E1 && up)) - __popc(__ballot_sync(0xffffffffu, E1 && dn));
                unsigned G0, G1;
                if (R) {
                    const int m = 31 - __clz(R);
                    G0 = G1 = ((F >> m) & 1u) ^ (__popc(F & ~R & ~((2u << m) - 1u)) & 1u);
                } else {
                    G0 = __popc(F) & 1u;
                    G1 = G0 ^ 1u;
                }
                if (lane == 0) s_slice[par][tile - t0][warp] = make_int4(D0, D1, (int)G0, (int)G1);
            }
            const i64 tx = warp_sum((i64)sx), td = warp_sum((i64)sd), tc = warp_sum((i64)cbase);
            if (lane == 0) {
                atomicAdd(&s_acc[par][0], (unsigned long long)tx);
                atomicAdd(&s_acc[par][1], (unsigned long long)td);
                atomicAdd(&s_acc[par][2], (unsigned long long)tc);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(ST_CONS_THREADS) : "memory");     // the consumer warps only
            if (threadIdx.x == 0) {
                i64 C0 = 0, C1 = 0;
                unsigned P0 = 0, P1 = 1;
                for (int t = 0; t < (int)(t1 - t0); t++)
                    for (int wq = 0; wq < ST_CONS_WARPS; wq++) {
                        const int4 sl = s_slice[par][t][wq];
                        C0 += P0 ? sl.y : sl.x;
                        C1 += P1 ? sl.y : sl.x;
                        P0 = P0 ? (unsigned)sl.w : (unsigned)sl.z;
                        P1 = P1 ? (unsigned)sl.w : (unsigned)sl.z;
                    }
                OrdSummary o;
                const i64 ax = (i64)s_acc[par][0], ad = (i64)s_acc[par][1], ac = (i64)s_acc[par][2];
                o.sum_q = (ax - ad) / 10;
                o.sum_x = ax;
                o.c0 = (unsigned)(ac + C0);
                o.c1 = (unsigned)(ac + C1);
                o.p0 = P0;
                o.p1 = P1;
                out[(i64)e * job_stride + w] = o;
                s_acc[par][0] = s_acc[par][1] = s_acc[par][2] = 0;      // this half is next used two items later: a barrier lies between
            }
            par ^= 1;
        }
    }
}

// ------------------------------------------------------------------------------
// First occurrence of every group among the rows that pass the predicate (the reference emits groups in
// first-insertion order, aggregate_hash.go:424-438).  Kept out of the scan's inner loop: CTAs walk 1024-row
// chunks from the start of the table and stop as soon as every group the totals say is present has been seen
// before their next chunk -- normally after one wave.
// ------------------------------------------------------------------------------
__device__ __forceinline__ i64 ncol_stored(const NCol &c, i64 row)
{
    switch (c.pw) {
    case 8: return __ldg((const i64 *)c.p + row);
    case 4: return (i64)__ldg((const int *)c.p + row);
    case 2: return (i64)__ldg((const unsigned short *)c.p + row);
    default: return (i64)__ldg((const uint8_t *)c.p + row);
    }
}

template <bool HAS_KEY1>
__global__ void __launch_bounds__(256)
first_rows_kernel(const LowcardParams p, const u64 *__restrict__ totals /* [G*6][2], this rank */, i64 *__restrict__ first_row /* [G], preset huge */)
{
    __shared__ i64 s_first[LC_MAXG];
    __shared__ int s_go;
    const int G = p.ngroups;
    const i64 nchunks = (p.nrows + 1023) / 1024;
    for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (threadIdx.x == 0) {
            int go = 0;
            for (int g = 0; g < G; g++) {
                const bool present = (totals[(size_t)(g * LC_K) * 2] | totals[(size_t)(g * LC_K) * 2 + 1]) != 0;
                if (present && *(volatile i64 *)&first_row[g] >= p.row_base + c * 1024) go = 1;
            }
            s_go = go;
        }
        if (threadIdx.x < LC_MAXG) s_first[threadIdx.x] = INT64_MAX;
        __syncthreads();
        if (!s_go) break;
        const i64 row = c * 1024 + threadIdx.x * 4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const i64 r = row + j;
            if (r >= p.nrows) break;
            const i64 dv = ncol_stored(p.pred, r);
            if (dv < p.lo || dv > p.hi) continue;
            int g = __ldg(p.luts + __ldg((const uint8_t *)p.key0.p + r));
            if (HAS_KEY1) g = g * p.n1 + __ldg(p.luts + 256 + __ldg((const uint8_t *)p.key1.p + r));
            if (p.row_base + r < s_first[g]) atomicMin((long long *)&s_first[g], (long long)(p.row_base + r));
        }
        __syncthreads();
        if (threadIdx.x < G && s_first[threadIdx.x] != INT64_MAX) atomicMin((long long *)&first_row[threadIdx.x], (long long)s_first[threadIdx.x]);
        __syncthreads();
    }
}

}  // namespace pg
