// scanagg.cuh -- fused scan + predicate + projection + aggregate kernels (sm_100a).
//
// These replace, for `Agg <- Scan[filters]` pipelines, the reference's per-chunk chain
//   scan filter     ExprExec.executeSelect        pkg/compute/expr_exec.go:342-530
//   projection      ExprExec.executeExprs         pkg/compute/expr_exec.go:85-340
//   decimal ops     binDecimalDecimal{Sub,Add,Mul}Op  pkg/compute/function_operator_binary.go:134-191
//   group lookup    GroupedAggrHashTable.FindOrCreateGroups pkg/compute/aggregate_hash.go:201-391
//   state update    UnaryScatter / SumOp / CountOp pkg/compute/function_aggr.go:770-1161
// with ONE pass over device-native columns: every referenced column is read exactly once
// with 16-byte coalesced streaming loads, the predicate and the fixed-point arithmetic run
// in registers, and no selection vector or intermediate vector is materialised.
//
// HBM-bound integer work: no tensor cores.  Grid = SMs x resident CTAs, grid-stride over
// 1024-row tiles so that concurrently running CTAs read neighbouring DRAM pages.
#pragma once
#include "common.cuh"

namespace pg {

constexpr int SA_THREADS = 256;
constexpr int SA_VEC = 4;                          // rows per thread per tile (16 B of int32)
constexpr int SA_TILE = SA_THREADS * SA_VEC;       // 1024 rows

// streaming loads: read-only path, do not allocate in L1 (each byte is used once)
__device__ __forceinline__ int4 ld_stream16(const void *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ longlong2 ld_stream16_ll(const void *p)
{
    longlong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ld_stream4(const void *p)
{
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ i64 warp_sum(i64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------
// Shape "sumprod": ungrouped  sum(fa * fb)  with inclusive range predicates on up to two
// int32 columns and on the two int64 factor columns (TPC-H Q6).
// Algorithmic bytes per row: 4*[HAS_A] + 4*[HAS_B] + 16.
// ------------------------------------------------------------------------------
struct SumProdParams {
    const int *pa, *pb;       // int32 / date32 predicate columns
    const i64 *fa, *fb;       // int64 factor columns (DECIMAL64 / BIGINT)
    int a_lo, a_hi, b_lo, b_hi;
    i64 fa_lo, fa_hi, fb_lo, fb_hi;
    i64 nrows;
};

template <bool HAS_A, bool HAS_B, int UNROLL>
__global__ void __launch_bounds__(SA_THREADS)
sumprod_kernel(const SumProdParams p, i64 *__restrict__ partials /* [grid][2] = {sum, count} */)
{
    const i64 ntiles = (p.nrows + SA_TILE - 1) / SA_TILE;
    i64 sum = 0, cnt = 0;
    for (i64 tile0 = blockIdx.x; tile0 < ntiles; tile0 += (i64)gridDim.x * UNROLL) {
        int4 a[UNROLL], b[UNROLL];
        longlong2 x[UNROLL][2], y[UNROLL][2];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            i64 tile = tile0 + (i64)u * gridDim.x;
            if (tile < ntiles) {
                i64 row = tile * SA_TILE + threadIdx.x * SA_VEC;
                if (HAS_A) a[u] = ld_stream16(p.pa + row);
                if (HAS_B) b[u] = ld_stream16(p.pb + row);
                x[u][0] = ld_stream16_ll(p.fa + row);
                x[u][1] = ld_stream16_ll(p.fa + row + 2);
                y[u][0] = ld_stream16_ll(p.fb + row);
                y[u][1] = ld_stream16_ll(p.fb + row + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            i64 tile = tile0 + (i64)u * gridDim.x;
            if (tile < ntiles) {
                i64 row = tile * SA_TILE + threadIdx.x * SA_VEC;
                i64 rem = p.nrows - row;   // rows of this vector that exist (pad rows are masked)
                int av[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
                int bv[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
                i64 xv[4] = {x[u][0].x, x[u][0].y, x[u][1].x, x[u][1].y};
                i64 yv[4] = {y[u][0].x, y[u][0].y, y[u][1].x, y[u][1].y};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    bool ok = j < rem;
                    if (HAS_A) ok = ok && av[j] >= p.a_lo && av[j] <= p.a_hi;
                    if (HAS_B) ok = ok && bv[j] >= p.b_lo && bv[j] <= p.b_hi;
                    ok = ok && xv[j] >= p.fa_lo && xv[j] <= p.fa_hi && yv[j] >= p.fb_lo && yv[j] <= p.fb_hi;
                    sum += ok ? xv[j] * yv[j] : 0;
                    cnt += ok ? 1 : 0;
                }
            }
        }
    }
    __shared__ i64 s_sum[SA_THREADS / 32], s_cnt[SA_THREADS / 32];
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        i64 s = 0, c = 0;
#pragma unroll
        for (int w = 0; w < SA_THREADS / 32; w++) { s += s_sum[w]; c += s_cnt[w]; }
        partials[2 * blockIdx.x] = s;
        partials[2 * blockIdx.x + 1] = c;
    }
}

// ------------------------------------------------------------------------------
// Shape "lowcard chain": GROUP BY up to two byte-coded columns (<= LC_MAXG dense groups)
// with the accumulator set
//   [0] count(*)          [1] sum(q)  (int32 column)      [2] sum(A)
//   [3] sum(A*(c1+s1*B))  [4] sum(A*(c1+s1*B)*(c2+s2*C))   [5] sum(B)
// over rows passing an inclusive range on one int32/date column (TPC-H Q1).
//
// Group state: every thread owns a private [group][acc] table in shared memory laid out
// [slot][thread] so a warp's 64-bit accesses hit 32 distinct bank pairs whatever the
// lanes' groups are -- plain LDS/STS, no atomics, no inter-thread conflicts.  One block
// reduction at the end, one partial per CTA, 128-bit merge in finalize128_kernel.
// Algorithmic bytes per row: 4 + nkeys + 4 + 24 (= 34 for Q1).
// ------------------------------------------------------------------------------
#ifndef PG_LC_THREADS
#define PG_LC_THREADS 256
#endif
constexpr int LC_THREADS = PG_LC_THREADS;          // threads per CTA of the low-cardinality kernels
constexpr int LC_TILE = LC_THREADS * SA_VEC;       // rows per tile of the low-cardinality kernels
constexpr int LC_K = 6;
constexpr int LC_MAXG = 8;

struct LowcardParams {
    const int *pred; int lo, hi;
    const uint8_t *key0, *key1;          // key1 may be null (single key)
    const int *q;
    const i64 *A, *B, *C;
    i64 c1, s1, c2, s2;
    const uint8_t *luts;                 // [2][256] code -> dense id, device memory
    int n1;                              // gid = lut0[k0] * n1 + lut1[k1]
    int ngroups;
    i64 nrows;
    i64 row_base;                        // global row id of local row 0 (for first_row)
    int contig;                          // 1: each CTA owns a contiguous run of tiles, so the per-CTA
                                         // partials are ORDERED partial sums (needed by the sequential
                                         // rounding emulation); 0: tiles interleaved across CTAs
};

// tile iteration space of one CTA: tiles t0 + u*ustride, t0 = tbeg, tbeg+tstep, ... < tend
struct TileIter { i64 tbeg, tend, tstep, ustride; };
template <int UNROLL>
__device__ __forceinline__ TileIter tile_iter(i64 ntiles, int contig)
{
    TileIter it;
    if (contig) {
        i64 per = (ntiles + gridDim.x - 1) / gridDim.x;
        it.tbeg = (i64)blockIdx.x * per;
        it.tend = it.tbeg + per < ntiles ? it.tbeg + per : ntiles;
        it.tstep = UNROLL;
        it.ustride = 1;
    } else {
        it.tbeg = blockIdx.x;
        it.tend = ntiles;
        it.tstep = (i64)gridDim.x * UNROLL;
        it.ustride = gridDim.x;
    }
    return it;
}

template <bool HAS_KEY1, int UNROLL>
__global__ void __launch_bounds__(LC_THREADS)   // 2 CTAs/SM; forcing 3 (80 regs + max SMEM carve-out) measured 1.8x SLOWER
lowcard_chain_kernel(const LowcardParams p, i64 *__restrict__ partials /* [grid][G*K] */,
                     i64 *__restrict__ first_row /* [G], pre-set to INT64_MAX */)
{
    extern __shared__ i64 s_acc[];                 // [G*K][LC_THREADS]
    __shared__ uint8_t s_lut[2][256];
    __shared__ i64 s_first[LC_MAXG];
    const int G = p.ngroups;
    for (int i = threadIdx.x; i < G * LC_K * LC_THREADS; i += LC_THREADS) s_acc[i] = 0;
    for (int i = threadIdx.x; i < 512; i += LC_THREADS) s_lut[i >> 8][i & 255] = p.luts[i];
    if (threadIdx.x < LC_MAXG) s_first[threadIdx.x] = INT64_MAX;
    __syncthreads();

    const i64 ntiles = (p.nrows + LC_TILE - 1) / LC_TILE;
    const TileIter it = tile_iter<UNROLL>(ntiles, p.contig);
    i64 *my = s_acc + threadIdx.x;
    for (i64 tile0 = it.tbeg; tile0 < it.tend; tile0 += it.tstep) {
        int4 d[UNROLL], q[UNROLL];
        unsigned k0[UNROLL], k1[UNROLL];
        longlong2 a[UNROLL][2], b[UNROLL][2], c[UNROLL][2];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            i64 tile = tile0 + (i64)u * it.ustride;
            if (tile < it.tend) {
                i64 row = tile * LC_TILE + threadIdx.x * SA_VEC;
                d[u] = ld_stream16(p.pred + row);
                k0[u] = ld_stream4(p.key0 + row);
                if (HAS_KEY1) k1[u] = ld_stream4(p.key1 + row);
                q[u] = ld_stream16(p.q + row);
                a[u][0] = ld_stream16_ll(p.A + row); a[u][1] = ld_stream16_ll(p.A + row + 2);
                b[u][0] = ld_stream16_ll(p.B + row); b[u][1] = ld_stream16_ll(p.B + row + 2);
                c[u][0] = ld_stream16_ll(p.C + row); c[u][1] = ld_stream16_ll(p.C + row + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            i64 tile = tile0 + (i64)u * it.ustride;
            if (tile < it.tend) {
                i64 row = tile * LC_TILE + threadIdx.x * SA_VEC;
                i64 rem = p.nrows - row;
                int dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
                int qv[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
                i64 av[4] = {a[u][0].x, a[u][0].y, a[u][1].x, a[u][1].y};
                i64 bv[4] = {b[u][0].x, b[u][0].y, b[u][1].x, b[u][1].y};
                i64 cv[4] = {c[u][0].x, c[u][0].y, c[u][1].x, c[u][1].y};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    bool ok = j < rem && dv[j] >= p.lo && dv[j] <= p.hi;
                    if (ok) {
                        int g = s_lut[0][(k0[u] >> (8 * j)) & 255];
                        if (HAS_KEY1) g = g * p.n1 + s_lut[1][(k1[u] >> (8 * j)) & 255];
                        i64 *t = my + g * (LC_K * LC_THREADS);
                        i64 n = t[0];
                        if (n == 0) atomicMin((long long *)&s_first[g], (long long)(p.row_base + row + j));
                        i64 t2 = av[j] * (p.c1 + p.s1 * bv[j]);
                        i64 t3 = t2 * (p.c2 + p.s2 * cv[j]);
                        t[0] = n + 1;
                        t[1 * LC_THREADS] += qv[j];
                        t[2 * LC_THREADS] += av[j];
                        t[3 * LC_THREADS] += t2;
                        t[4 * LC_THREADS] += t3;
                        t[5 * LC_THREADS] += bv[j];
                    }
                }
            }
        }
    }
    __syncthreads();
    // block reduction: warp w sums slots w, w+8, ... across the 256 private copies
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = warp; v < G * LC_K; v += LC_THREADS / 32) {
        i64 s = 0;
#pragma unroll
        for (int j = 0; j < LC_THREADS / 32; j++) s += s_acc[v * LC_THREADS + lane + 32 * j];
        s = warp_sum(s);
        if (lane == 0) partials[(i64)blockIdx.x * (G * LC_K) + v] = s;
    }
    if (threadIdx.x < G && s_first[threadIdx.x] != INT64_MAX)
        atomicMin((long long *)&first_row[threadIdx.x], (long long)s_first[threadIdx.x]);
}

// ------------------------------------------------------------------------------
// Sequential-rounding emulation (govalues >19-digit regime).
//
// The reference folds sum(DECIMAL) with Decimal.Add in scan order
// (function_aggr.go:684-689).  Once a running sum needs 20 digits the library keeps 19 and
// rounds EVERY further addition half-to-even, so the result depends on row order:
//     S' = S + floor(x/10) + carry,  carry = [d>5] or ([d==5] and (S + floor(x/10)) odd),
// with d = x mod 10.  Each row is therefore a map parity(S) -> (parity(S'), carry) plus an
// exact floor(x/10); maps compose associatively, so a tile is summarised in parallel and
// tiles are composed in order on the host.  Only the rows after the crossing point are
// re-read (about 8% of lineitem for Q1's (N,O) sum_charge at SF100).
// ------------------------------------------------------------------------------
struct OrdParams {
    LowcardParams base;
    int group;       // dense group id to follow
    int slot;        // accumulator slot whose value sequence is followed (2..5)
    i64 tile_begin, tile_end;
    int chunk;       // consecutive tiles composed (in order) into ONE output summary by a CTA
};

struct OrdSummary {        // one per tile
    i64 sum_q;             // sum of floor(x/10)            (transducer kernel)
    i64 sum_x;             // exact sum of x                (both kernels)
    unsigned c0, c1;       // carries produced when entering with even / odd parity
    unsigned p0, p1;       // parity on exit
};

__device__ __forceinline__ i64 ord_value(const LowcardParams &p, int slot, i64 a, i64 b, i64 c)
{
    i64 t2 = a * (p.c1 + p.s1 * b);
    switch (slot) {
    case 2: return a;
    case 3: return t2;
    case 4: return t2 * (p.c2 + p.s2 * c);
    default: return b;
    }
}

struct OrdState { i64 sq, sx; unsigned c0, c1, p0, p1; };
__device__ __forceinline__ OrdState ord_compose(const OrdState &l, const OrdState &r)
{
    OrdState o;
    o.sq = l.sq + r.sq;
    o.sx = l.sx + r.sx;
    o.p0 = l.p0 ? r.p1 : r.p0;
    o.c0 = l.c0 + (l.p0 ? r.c1 : r.c0);
    o.p1 = l.p1 ? r.p1 : r.p0;
    o.c1 = l.c1 + (l.p1 ? r.c1 : r.c0);
    return o;
}

template <bool HAS_KEY1>
__global__ void __launch_bounds__(LC_THREADS)
ord_tile_kernel(const OrdParams op, OrdSummary *__restrict__ out /* [ceil((tile_end - tile_begin) / chunk)] */)
{
    const LowcardParams &p = op.base;
    __shared__ uint8_t s_lut[2][256];
    __shared__ OrdState s_w[LC_THREADS / 32];
    for (int i = threadIdx.x; i < 512; i += LC_THREADS) s_lut[i >> 8][i & 255] = p.luts[i];
    __syncthreads();
    const i64 nchunks = (op.tile_end - op.tile_begin + op.chunk - 1) / op.chunk;
    for (i64 ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        OrdState run = {0, 0, 0, 0, 0, 1};      // thread 0: ordered composition of the chunk's tiles
        const i64 t0 = op.tile_begin + ch * op.chunk, t1 = t0 + op.chunk < op.tile_end ? t0 + op.chunk : op.tile_end;
        for (i64 tile = t0; tile < t1; tile++) {
            i64 row = tile * LC_TILE + threadIdx.x * SA_VEC;
            i64 rem = p.nrows - row;
            int4 d = ld_stream16(p.pred + row);
            unsigned k0 = ld_stream4(p.key0 + row), k1 = HAS_KEY1 ? ld_stream4(p.key1 + row) : 0;
            longlong2 a0 = ld_stream16_ll(p.A + row), a1 = ld_stream16_ll(p.A + row + 2);
            longlong2 b0 = ld_stream16_ll(p.B + row), b1 = ld_stream16_ll(p.B + row + 2);
            longlong2 c0 = ld_stream16_ll(p.C + row), c1 = ld_stream16_ll(p.C + row + 2);
            int dv[4] = {d.x, d.y, d.z, d.w};
            i64 av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y}, cv[4] = {c0.x, c0.y, c1.x, c1.y};
            OrdState st = {0, 0, 0, 0, 0, 1};   // identity: parity preserved, no carries
#pragma unroll
            for (int j = 0; j < 4; j++) {
                bool ok = j < rem && dv[j] >= p.lo && dv[j] <= p.hi;
                int g = s_lut[0][(k0 >> (8 * j)) & 255];
                if (HAS_KEY1) g = g * p.n1 + s_lut[1][(k1 >> (8 * j)) & 255];
                if (ok && g == op.group) {
                    i64 x = ord_value(p, op.slot, av[j], bv[j], cv[j]);
                    i64 q = x / 10;
                    unsigned dgt = (unsigned)(x - q * 10), qb = (unsigned)(q & 1);
                    OrdState r;
                    r.sq = q;
                    r.sx = x;
                    unsigned t0b = qb, t1b = qb ^ 1u;     // parity of S + q when entering even / odd
                    r.c0 = (dgt > 5u) | ((dgt == 5u) & t0b);
                    r.c1 = (dgt > 5u) | ((dgt == 5u) & t1b);
                    r.p0 = t0b ^ r.c0;
                    r.p1 = t1b ^ r.c1;
                    st = ord_compose(st, r);
                }
            }
            // ordered composition across the warp, then across the 8 warps
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                OrdState r;
                r.sq = __shfl_down_sync(0xffffffffu, st.sq, o);
                r.sx = __shfl_down_sync(0xffffffffu, st.sx, o);
                r.c0 = __shfl_down_sync(0xffffffffu, st.c0, o);
                r.c1 = __shfl_down_sync(0xffffffffu, st.c1, o);
                r.p0 = __shfl_down_sync(0xffffffffu, st.p0, o);
                r.p1 = __shfl_down_sync(0xffffffffu, st.p1, o);
                if ((threadIdx.x & 31) + o < 32) st = ord_compose(st, r);
            }
            if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = st;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 0; w < LC_THREADS / 32; w++) run = ord_compose(run, s_w[w]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            OrdSummary o;
            o.sum_q = run.sq; o.sum_x = run.sx; o.c0 = run.c0; o.c1 = run.c1; o.p0 = run.p0; o.p1 = run.p1;
            out[ch] = o;
        }
    }
}

// ------------------------------------------------------------------------------
// Shape-agnostic scan-aggregate: any conjunction of range predicates on typed columns, up to
// GEN_MAXACC aggregates (sum / min / max of a product of <= 3 affine factors, or count), grouped
// by 0..2 byte-coded keys.  Same thread-private shared-memory group tables as the chain kernel,
// but columns, predicates and factors are runtime descriptors (one row per thread per step,
// typed scalar loads), so it is the fallback for shapes without a specialised kernel:
// correct for everything the lowering accepts, slower than the specialised kernels.
// ------------------------------------------------------------------------------
constexpr int GEN_MAXPRED = 8, GEN_MAXACC = 8, GEN_MAXFAC = 3;
enum { GEN_SUM = 0, GEN_MIN = 1, GEN_MAX = 2, GEN_COUNTV = 3 };   // COUNTV: count(column) = rows where it is not NULL

struct GenCol { const void *p; int width; const uint8_t *valid; };   // valid: packed bits, 1 = not NULL; null = no NULLs
__device__ __forceinline__ bool gen_valid(const GenCol &c, i64 row)
{
    return c.valid == nullptr || ((__ldg(c.valid + (row >> 3)) >> (row & 7)) & 1);
}
__device__ __forceinline__ i64 gen_load(const GenCol &c, i64 row)
{
    switch (c.width) {
    case 8: return __ldg((const i64 *)c.p + row);
    case 4: return (i64)__ldg((const int *)c.p + row);
    default: return (i64)__ldg((const uint8_t *)c.p + row);
    }
}

struct GenAcc {
    int kind, nfac;
    GenCol fac[GEN_MAXFAC];
    i64 c[GEN_MAXFAC];
    int s[GEN_MAXFAC];
};

// string predicate on a VARCHAR column (device copy: bytes + offsets): kind 1 LIKE, 2 NOT LIKE, 3 =, 4 <>
constexpr int GEN_MAXLIKE = 2, GEN_PATMAX = 48;
struct GenLike {
    const char *bytes;
    const i64 *off;
    int kind, plen;
    char pat[GEN_PATMAX];
};
// wildcardMatch of the reference (function_operator_boolean.go:336-377), byte for byte
__device__ __forceinline__ bool gen_wildcard_match(const char *pat, int plen, const char *tgt, i64 tlen)
{
    i64 p = 0, t = 0, star_p = -1, star_t = -1;
    while (t < tlen) {
        if (p < plen && pat[p] == '%') {
            p++;
            star_p = p;
            if (p >= plen) return true;
            star_t = t;
        } else if (p < plen && (pat[p] == '_' || pat[p] == tgt[t])) {
            p++;
            t++;
        } else {
            if (star_p == -1 || star_t == -1) return false;
            p = star_p;
            star_t++;
            t = star_t;
        }
    }
    while (p < plen && pat[p] == '%') p++;
    return p >= plen;
}
// `%lit%` with a wildcard-free literal of 1..8 bytes (the shape of `p_name like '%pink%'`): word-at-a-time
// search -- candidate positions of the literal's first byte come from the zero-byte trick on aligned 8-byte
// words, only those are verified against the (masked) literal.  The byte buffer is 8-byte aligned and padded.
__device__ __forceinline__ bool gen_contains(const char *bytes, i64 b, i64 e, u64 lit, int L)
{
    if (e - b < L) return false;
    const u64 ONES = 0x0101010101010101ULL, HIGH = 0x8080808080808080ULL;
    const u64 mask = L >= 8 ? ~0ULL : ((1ULL << (8 * L)) - 1);
    const u64 first = ONES * (lit & 0xff);
    const u64 *base = (const u64 *)bytes;
    const i64 wb = b >> 3, we = (e - 1) >> 3;
    u64 cur = __ldg(base + wb);
    for (i64 w = wb; w <= we; w++) {
        const u64 nxt = __ldg(base + w + 1);
        const u64 x = cur ^ first;
        u64 t = (x - ONES) & ~x & HIGH;          // bytes equal to the first literal byte (plus rare false positives, verified below)
        while (t) {
            const int j = (__ffsll((long long)t) - 1) >> 3;
            t &= t - 1;
            const i64 p = (w << 3) + j;
            if (p < b || p + L > e) continue;
            const int sh = j * 8;
            const u64 win = sh == 0 ? cur : (cur >> sh) | (nxt << (64 - sh));
            if ((win & mask) == lit) return true;
        }
        cur = nxt;
    }
    return false;
}

__device__ __forceinline__ bool gen_like_pass(const GenLike &l, i64 row)
{
    const i64 b = __ldg(l.off + row), e = __ldg(l.off + row + 1);
    if (l.kind >= 5) {                    // 5: contains, 6: does not contain (fast path of LIKE / NOT LIKE '%lit%')
        u64 lit = 0;
        for (int i = 0; i < l.plen; i++) lit |= (u64)(unsigned char)l.pat[i] << (8 * i);
        return gen_contains(l.bytes, b, e, lit, l.plen) == (l.kind == 5);
    }
    const char *tgt = l.bytes + b;
    bool m;
    if (l.kind <= 2) {
        m = gen_wildcard_match(l.pat, l.plen, tgt, e - b);
    } else {
        m = (e - b) == l.plen;
        for (int i = 0; m && i < l.plen; i++) m = tgt[i] == l.pat[i];
    }
    return m == (l.kind == 1 || l.kind == 3);
}

struct GenParams {
    i64 nrows, row_base;
    int nlike;
    GenLike like[GEN_MAXLIKE];
    int npred;
    GenCol pcol[GEN_MAXPRED];
    i64 plo[GEN_MAXPRED], phi[GEN_MAXPRED];
    int pset[GEN_MAXPRED];                 // 1: the predicate is a code set on a byte column (IN, <>, OR of =)
    unsigned pmask[GEN_MAXPRED][8];
    int nkeys;
    const uint8_t *key0, *key1;
    const uint8_t *luts;
    int n1, ngroups;
    int nacc;                       // plane 0 is always the row count; planes 1..nacc the aggregates;
                                    // with NULLS, planes nacc+1..2*nacc count the non-NULL inputs of each aggregate
    GenAcc acc[GEN_MAXACC];
};
__host__ __device__ __forceinline__ int gen_plane_kind(const GenParams &p, int plane)
{
    if (plane == 0 || plane > p.nacc) return GEN_SUM;
    int k = p.acc[plane - 1].kind;
    return k == GEN_COUNTV ? GEN_SUM : k;
}

// NULL semantics of the reference (masks AND-ed through expressions, function_operator_binary.go:267-481;
// selection skips NULL operands, function_operator_boolean.go:780-868; aggregates ignore NULL inputs,
// function_aggr.go IgnoreNull): a row with a NULL predicate operand is not selected; an aggregate
// skips rows where any column of its argument is NULL.
template <int NT, bool NULLS>
__global__ void __launch_bounds__(NT)
generic_scanagg_kernel(const GenParams p, i64 *__restrict__ partials /* [grid][G*P] */,
                       i64 *__restrict__ first_row /* [G] preset to 0x7f.. */)
{
    extern __shared__ i64 s_acc[];                 // [G*P][NT]
    __shared__ uint8_t s_lut[2][256];
    __shared__ i64 s_first[64];
    const int G = p.ngroups, P = NULLS ? 1 + 2 * p.nacc : 1 + p.nacc;
    for (int i = threadIdx.x; i < G * P * NT; i += NT) {
        int kind = gen_plane_kind(p, (i / NT) % P);
        s_acc[i] = kind == GEN_MIN ? INT64_MAX : kind == GEN_MAX ? INT64_MIN : 0;
    }
    if (p.nkeys > 0) for (int i = threadIdx.x; i < 512; i += NT) s_lut[i >> 8][i & 255] = p.luts[i];
    if (threadIdx.x < 64) s_first[threadIdx.x] = INT64_MAX;
    __syncthreads();
    i64 *my = s_acc + threadIdx.x;
    for (i64 row = (i64)blockIdx.x * NT + threadIdx.x; row < p.nrows; row += (i64)gridDim.x * NT) {
        bool ok = true;
        for (int k = 0; k < p.npred && ok; k++) {
            if (NULLS && !gen_valid(p.pcol[k], row)) { ok = false; break; }
            i64 v = gen_load(p.pcol[k], row);
            ok = p.pset[k] ? ((p.pmask[k][(v >> 5) & 7] >> (v & 31)) & 1u) != 0 : (v >= p.plo[k] && v <= p.phi[k]);
        }
        for (int k = 0; k < p.nlike && ok; k++) ok = gen_like_pass(p.like[k], row);
        if (!ok) continue;
        int g = 0;
        if (p.nkeys > 0) g = s_lut[0][__ldg(p.key0 + row)];
        if (p.nkeys > 1) g = g * p.n1 + s_lut[1][__ldg(p.key1 + row)];
        i64 *t = my + (i64)g * P * NT;
        i64 n = t[0];
        if (n == 0) atomicMin((long long *)&s_first[g], (long long)(p.row_base + row));
        t[0] = n + 1;
        for (int a = 0; a < p.nacc; a++) {
            const GenAcc &A = p.acc[a];
            i64 x = 1;
            bool vok = true;
            for (int f = 0; f < A.nfac; f++) {
                if (NULLS && !gen_valid(A.fac[f], row)) { vok = false; break; }
                x *= A.c[f] + A.s[f] * gen_load(A.fac[f], row);
            }
            if (!vok) continue;
            i64 *slot = t + (i64)(a + 1) * NT;
            i64 cur = *slot;
            *slot = A.kind == GEN_SUM ? cur + x : A.kind == GEN_MIN ? (x < cur ? x : cur) : A.kind == GEN_MAX ? (x > cur ? x : cur) : cur + 1;
            if (NULLS) t[(i64)(p.nacc + 1 + a) * NT] += 1;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = warp; v < G * P; v += NT / 32) {
        int kind = gen_plane_kind(p, v % P);
        i64 r = kind == GEN_SUM ? 0 : kind == GEN_MIN ? INT64_MAX : INT64_MIN;
        for (int j = 0; j < NT / 32; j++) {
            i64 x = s_acc[v * NT + lane + 32 * j];
            r = kind == GEN_SUM ? r + x : kind == GEN_MIN ? (x < r ? x : r) : (x > r ? x : r);
        }
        for (int o = 16; o > 0; o >>= 1) {
            i64 x = __shfl_xor_sync(0xffffffffu, r, o);
            r = kind == GEN_SUM ? r + x : kind == GEN_MIN ? (x < r ? x : r) : (x > r ? x : r);
        }
        if (lane == 0) partials[(i64)blockIdx.x * (G * P) + v] = r;
    }
    if (threadIdx.x < G && s_first[threadIdx.x] != INT64_MAX)
        atomicMin((long long *)&first_row[threadIdx.x], (long long)s_first[threadIdx.x]);
}

// per-CTA partials -> totals: sums exact in 128 bits, min/max by comparison.  kinds[v] per value.
static __global__ void finalize_generic_kernel(const i64 *__restrict__ partials, int nblocks, int nvals,
                                               const int *__restrict__ kinds, u64 *__restrict__ out /* [nvals][2] */)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvals) return;
    int kind = kinds[v];
    u64 lo = 0;
    i64 hi = 0;
    i64 m = kind == GEN_MIN ? INT64_MAX : INT64_MIN;
    for (int b = 0; b < nblocks; b++) {
        i64 x = partials[(i64)b * nvals + v];
        if (kind == GEN_SUM) {
            u64 nlo = lo + (u64)x;
            hi += (x < 0 ? -1 : 0) + (nlo < lo ? 1 : 0);
            lo = nlo;
        } else if (kind == GEN_MIN) m = x < m ? x : m;
        else m = x > m ? x : m;
    }
    if (kind == GEN_SUM) { out[2 * v] = lo; out[2 * v + 1] = (u64)hi; }
    else { out[2 * v] = (u64)m; out[2 * v + 1] = m < 0 ? ~0ULL : 0ULL; }
}

// ------------------------------------------------------------------------------
// Merge per-CTA int64 partials into exact 128-bit totals: out[v] = {lo, hi}.
// The reference accumulates in a 128-bit Hugeint (function_aggr.go:620-630) or a
// 19-digit Decimal (:684-689); per-CTA sums are proven < 2^63 at plan time from the
// column statistics, the cross-CTA total is carried in 128 bits.
// ------------------------------------------------------------------------------
static __global__ void finalize128_kernel(const i64 *__restrict__ partials, int nblocks, int nvals,
                                   u64 *__restrict__ out /* [nvals][2] */)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvals) return;
    u64 lo = 0;
    i64 hi = 0;
    for (int b = 0; b < nblocks; b++) {
        i64 x = partials[(i64)b * nvals + v];
        u64 nlo = lo + (u64)x;
        hi += (x < 0 ? -1 : 0) + (nlo < lo ? 1 : 0);
        lo = nlo;
    }
    out[2 * v] = lo;
    out[2 * v + 1] = (u64)hi;
}

}  // namespace pg
