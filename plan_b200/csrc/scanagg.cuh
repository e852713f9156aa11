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
#include <type_traits>

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

__device__ __forceinline__ uint2 ld_stream8(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------
// Physical column access.  pg_table_seal stores every integer-family column at the narrowest
// width its min/max statistics allow (frame of reference: logical = base + stored; NCol in
// common.cuh), so the scan kernels read 1/2/4/8-byte arrays.  A thread always handles SA_VEC = 4
// consecutive rows: one 4/8/16/32-byte streaming load per column, unpacked in registers.  The
// width is a kernel parameter (warp-uniform branch), not a template parameter: the table decides
// it at run time and the combinations would not be enumerable.
//   WIDE = false: every column of the kernel is stored in <= 4 bytes, raw vector = 16 bytes;
//   WIDE = true : 8-byte columns may occur, raw vector = 32 bytes, values are handled as i64.
// ------------------------------------------------------------------------------
template <bool WIDE> struct Raw4;
template <> struct Raw4<false> { int4 a; };
template <> struct Raw4<true> { int4 a, b; };

__device__ __forceinline__ void ld_raw4(const NCol &c, i64 row, Raw4<false> &r)
{
    const char *p = (const char *)c.p + row * c.pw;
    if (c.pw == 1) r.a.x = (int)ld_stream4(p);
    else if (c.pw == 2) { const uint2 t = ld_stream8(p); r.a.x = (int)t.x; r.a.y = (int)t.y; }
    else r.a = ld_stream16(p);
}
__device__ __forceinline__ void ld_raw4(const NCol &c, i64 row, Raw4<true> &r)
{
    const char *p = (const char *)c.p + row * c.pw;
    if (c.pw == 1) r.a.x = (int)ld_stream4(p);
    else if (c.pw == 2) { const uint2 t = ld_stream8(p); r.a.x = (int)t.x; r.a.y = (int)t.y; }
    else if (c.pw == 4) r.a = ld_stream16(p);
    else { r.a = ld_stream16(p); r.b = ld_stream16(p + 16); }
}
// STORED values (base not added).  32-bit form: pw <= 4 only.
__device__ __forceinline__ void unpack4(const NCol &c, const Raw4<false> &r, int (&v)[4])
{
    if (c.pw == 1) {
        const unsigned x = (unsigned)r.a.x;
        v[0] = (int)(x & 255u); v[1] = (int)((x >> 8) & 255u); v[2] = (int)((x >> 16) & 255u); v[3] = (int)(x >> 24);
    } else if (c.pw == 2) {
        const unsigned x = (unsigned)r.a.x, y = (unsigned)r.a.y;
        v[0] = (int)(x & 0xffffu); v[1] = (int)(x >> 16); v[2] = (int)(y & 0xffffu); v[3] = (int)(y >> 16);
    } else {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
    }
}
__device__ __forceinline__ void unpack4(const NCol &c, const Raw4<true> &r, i64 (&v)[4])
{
    if (c.pw == 8) {
        v[0] = (i64)(((u64)(unsigned)r.a.y << 32) | (unsigned)r.a.x); v[1] = (i64)(((u64)(unsigned)r.a.w << 32) | (unsigned)r.a.z);
        v[2] = (i64)(((u64)(unsigned)r.b.y << 32) | (unsigned)r.b.x); v[3] = (i64)(((u64)(unsigned)r.b.w << 32) | (unsigned)r.b.z);
    } else {
        Raw4<false> n;
        n.a = r.a;
        int t[4];
        unpack4(c, n, t);
        v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; v[3] = t[3];
    }
}
// one row by id (gathers, tails): LOGICAL value
__device__ __forceinline__ i64 ncol_load(const NCol &c, i64 row)
{
    switch (c.pw) {
    case 8: return __ldg((const i64 *)c.p + row);
    case 4: return (i64)__ldg((const int *)c.p + row) + c.base;
    case 2: return (i64)__ldg((const unsigned short *)c.p + row) + c.base;
    default: return (i64)__ldg((const uint8_t *)c.p + row) + c.base;
    }
}

// ------------------------------------------------------------------------------
// Shape "sumprod": ungrouped  sum(fa * fb)  with inclusive range predicates on up to two further
// columns and on the two factor columns (TPC-H Q6).  Bounds are in each column's STORED domain
// (the host subtracts the base and clamps), so a predicate is two integer compares on the raw
// value.  Algorithmic bytes per row = the physical widths of the referenced columns
// (Q6 at SF100: shipdate 2 + quantity 1 + discount 1 + extendedprice 4 = 8; 24 at native widths).
// ------------------------------------------------------------------------------
struct SumProdParams {
    NCol pa, pb;              // predicate-only columns
    NCol fa, fb;              // factor columns
    i64 a_lo, a_hi, b_lo, b_hi, fa_lo, fa_hi, fb_lo, fb_hi;     // stored-domain bounds
    i64 nrows;
};

template <bool WIDE, bool HAS_A, bool HAS_B, int UNROLL>
__global__ void __launch_bounds__(SA_THREADS)
sumprod_kernel(const SumProdParams p, i64 *__restrict__ partials /* [grid][2] = {sum, count} */)
{
    typedef typename std::conditional<WIDE, i64, int>::type V;
    const i64 ntiles = (p.nrows + SA_TILE - 1) / SA_TILE;
    const V alo = (V)p.a_lo, ahi = (V)p.a_hi, blo = (V)p.b_lo, bhi = (V)p.b_hi;
    const V xlo = (V)p.fa_lo, xhi = (V)p.fa_hi, ylo = (V)p.fb_lo, yhi = (V)p.fb_hi;
    const V xbase = (V)p.fa.base, ybase = (V)p.fb.base;
    i64 sum = 0;
    unsigned cnt = 0;
    i64 cnt64 = 0;
    for (i64 tile0 = blockIdx.x; tile0 < ntiles; tile0 += (i64)gridDim.x * UNROLL) {
        Raw4<WIDE> a[UNROLL], b[UNROLL], x[UNROLL], y[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const i64 tile = tile0 + (i64)u * gridDim.x;
            if (tile < ntiles) {
                const i64 row = tile * SA_TILE + threadIdx.x * SA_VEC;
                if (HAS_A) ld_raw4(p.pa, row, a[u]);
                if (HAS_B) ld_raw4(p.pb, row, b[u]);
                ld_raw4(p.fa, row, x[u]);
                ld_raw4(p.fb, row, y[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const i64 tile = tile0 + (i64)u * gridDim.x;
            if (tile < ntiles) {
                const i64 row = tile * SA_TILE + threadIdx.x * SA_VEC;
                const i64 rem = p.nrows - row;   // rows of this vector that exist (pad rows are masked)
                V av[4], bv[4], xv[4], yv[4];
                if (HAS_A) unpack4(p.pa, a[u], av);
                if (HAS_B) unpack4(p.pb, b[u], bv);
                unpack4(p.fa, x[u], xv);
                unpack4(p.fb, y[u], yv);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    bool ok = j < rem;
                    if (HAS_A) ok = ok && av[j] >= alo && av[j] <= ahi;
                    if (HAS_B) ok = ok && bv[j] >= blo && bv[j] <= bhi;
                    ok = ok && xv[j] >= xlo && xv[j] <= xhi && yv[j] >= ylo && yv[j] <= yhi;
                    const i64 prod = (i64)(V)(xv[j] + xbase) * (i64)(V)(yv[j] + ybase);   // NARROW: one 32x32->64 multiply
                    sum += ok ? prod : 0;
                    cnt += ok ? 1u : 0u;
                }
            }
        }
        cnt64 += cnt;      // a thread's per-iteration count is tiny; the running total is 64-bit
        cnt = 0;
    }
    __shared__ i64 s_sum[SA_THREADS / 32], s_cnt[SA_THREADS / 32];
    sum = warp_sum(sum);
    cnt64 = warp_sum(cnt64);
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt64; }
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
//   [0] count(*)          [1] sum(q)                      [2] sum(A)
//   [3] sum(A*(c1+s1*B))  [4] sum(A*(c1+s1*B)*(c2+s2*C))   [5] sum(B)
// over rows passing an inclusive range on one column (TPC-H Q1).
//
// Group state: every thread owns a private [group][acc] table in shared memory laid out
// [slot][thread], so a warp's accesses hit 32 distinct banks whatever the lanes' groups are --
// plain LDS/STS, no atomics.  At 11 stored bytes per row (Q1 at SF100: 2+1+1+1+4+1+1) the HBM
// roofline allows ~2 rows/clk/SM, which twelve 64-bit shared-memory accesses per row (96 B against
// 128 B/clk/SM) cannot feed.  Two measures bring the table traffic down:
//   * RUNS: a thread keeps the sums of its current group in registers and only touches its table
//     when the group changes (fact tables are clustered: half of lineitem is one (N,O) run);
//   * ACC32: the three small accumulators (count, sum of stored q, sum of stored B) use 32-bit
//     table slots when the statistics prove a thread's total fits (host: rows per thread x max
//     stored value < 2^32).
// sums of q, A and B are kept over the STORED values; the block reduction adds count x base.
// One partial per CTA, 128-bit merge in finalize128_kernel.
// ------------------------------------------------------------------------------
#ifndef PG_LC_THREADS
#define PG_LC_THREADS 256
#endif
constexpr int LC_THREADS = PG_LC_THREADS;          // threads per CTA of the low-cardinality kernels
constexpr int LC_TILE = LC_THREADS * SA_VEC;       // rows per tile of the low-cardinality kernels
constexpr int LC_K = 6;
constexpr int LC_MAXG = 8;

struct LowcardParams {
    NCol pred; i64 lo, hi;               // stored-domain bounds of the predicate column
    NCol key0, key1;                     // byte-coded keys (pw == 1); key1.p may be null (single key)
    NCol q, A, B, C;
    i64 c1, s1, c2, s2;                  // factors on LOGICAL values: (c1 + s1*B), (c2 + s2*C)
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

// bytes of one thread's private table: 3 (ACC32) or 0 32-bit planes + the 64-bit ones
__host__ __device__ constexpr int lc_smem_per_thread(int G, bool acc32) { return acc32 ? G * (3 * 8 + 3 * 4) : G * LC_K * 8; }

template <bool WIDE, bool ACC32, bool HAS_KEY1, int UNROLL>
__global__ void __launch_bounds__(LC_THREADS)
lowcard_chain_kernel(const LowcardParams p, i64 *__restrict__ partials /* [grid][G*K] */,
                     i64 *__restrict__ first_row /* [G], pre-set to INT64_MAX */)
{
    static_assert(!(WIDE && ACC32), "32-bit table slots are only used by the narrow kernel");
    typedef typename std::conditional<WIDE, i64, int>::type V;
    typedef typename std::conditional<ACC32, unsigned, u64>::type S;      // small accumulators
    extern __shared__ i64 s_dyn[];
    // 64-bit planes: slots {A, t2, t3} (ACC32) or all six; then the 32-bit planes {count, q, B}
    __shared__ uint8_t s_lut[2][256];
    __shared__ i64 s_first[LC_MAXG];
    __shared__ i64 s_tot[LC_MAXG * LC_K];
    const int G = p.ngroups;
    const int N64 = ACC32 ? 3 : LC_K;
    u64 *s64 = (u64 *)s_dyn;                                   // [G*N64][LC_THREADS]
    unsigned *s32 = (unsigned *)(s64 + (size_t)G * N64 * LC_THREADS);   // [G*3][LC_THREADS]   (ACC32 only)
    for (int i = threadIdx.x; i < G * N64 * LC_THREADS; i += LC_THREADS) s64[i] = 0;
    if (ACC32) for (int i = threadIdx.x; i < G * 3 * LC_THREADS; i += LC_THREADS) s32[i] = 0;
    for (int i = threadIdx.x; i < 512; i += LC_THREADS) s_lut[i >> 8][i & 255] = p.luts[i];
    if (threadIdx.x < LC_MAXG) s_first[threadIdx.x] = INT64_MAX;
    __syncthreads();

    const i64 ntiles = (p.nrows + LC_TILE - 1) / LC_TILE;
    const TileIter it = tile_iter<UNROLL>(ntiles, p.contig);
    const V plo = (V)p.lo, phi = (V)p.hi;
    const V abase = (V)p.A.base;
    const V f1c = (V)(p.c1 + p.s1 * p.B.base), f1s = (V)p.s1;      // (c1 + s1*B) on the stored value of B
    const V f2c = (V)(p.c2 + p.s2 * p.C.base), f2s = (V)p.s2;
    // the thread's current run: group cg, sums in registers
    int cg = -1;
    S rn = 0, rq = 0, rb = 0;
    i64 ra = 0, r2 = 0, r3 = 0, rfirst = 0;
    auto flush = [&]() {
        if (cg < 0) return;
        if (ACC32) {
            u64 *t = s64 + (size_t)cg * 3 * LC_THREADS + threadIdx.x;
            unsigned *w = s32 + (size_t)cg * 3 * LC_THREADS + threadIdx.x;
            const unsigned n = w[0];
            if (n == 0) atomicMin((long long *)&s_first[cg], (long long)rfirst);
            w[0] = n + (unsigned)rn;
            w[1 * LC_THREADS] += (unsigned)rq;
            w[2 * LC_THREADS] += (unsigned)rb;
            t[0] += (u64)ra;
            t[1 * LC_THREADS] += (u64)r2;
            t[2 * LC_THREADS] += (u64)r3;
        } else {
            u64 *t = s64 + (size_t)cg * LC_K * LC_THREADS + threadIdx.x;
            const u64 n = t[0];
            if (n == 0) atomicMin((long long *)&s_first[cg], (long long)rfirst);
            t[0] = n + (u64)rn;
            t[1 * LC_THREADS] += (u64)rq;
            t[2 * LC_THREADS] += (u64)ra;
            t[3 * LC_THREADS] += (u64)r2;
            t[4 * LC_THREADS] += (u64)r3;
            t[5 * LC_THREADS] += (u64)rb;
        }
    };
    for (i64 tile0 = it.tbeg; tile0 < it.tend; tile0 += it.tstep) {
        Raw4<WIDE> d[UNROLL], q[UNROLL], a[UNROLL], b[UNROLL], c[UNROLL];
        unsigned k0[UNROLL], k1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const i64 tile = tile0 + (i64)u * it.ustride;
            if (tile < it.tend) {
                const i64 row = tile * LC_TILE + threadIdx.x * SA_VEC;
                ld_raw4(p.pred, row, d[u]);
                k0[u] = ld_stream4((const uint8_t *)p.key0.p + row);
                if (HAS_KEY1) k1[u] = ld_stream4((const uint8_t *)p.key1.p + row);
                ld_raw4(p.q, row, q[u]);
                ld_raw4(p.A, row, a[u]);
                ld_raw4(p.B, row, b[u]);
                ld_raw4(p.C, row, c[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const i64 tile = tile0 + (i64)u * it.ustride;
            if (tile < it.tend) {
                const i64 row = tile * LC_TILE + threadIdx.x * SA_VEC;
                const i64 rem = p.nrows - row;
                V dv[4], qv[4], av[4], bv[4], cv[4];
                unpack4(p.pred, d[u], dv);
                unpack4(p.q, q[u], qv);
                unpack4(p.A, a[u], av);
                unpack4(p.B, b[u], bv);
                unpack4(p.C, c[u], cv);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool ok = j < rem && dv[j] >= plo && dv[j] <= phi;
                    if (ok) {
                        int g = s_lut[0][(k0[u] >> (8 * j)) & 255];
                        if (HAS_KEY1) g = g * p.n1 + s_lut[1][(k1[u] >> (8 * j)) & 255];
                        if (g != cg) {
                            flush();
                            cg = g;
                            rn = 0; rq = 0; rb = 0; ra = 0; r2 = 0; r3 = 0;
                            rfirst = p.row_base + row + j;
                        }
                        const V al = av[j] + abase;                           // logical A
                        const i64 t2 = (i64)al * (i64)(V)(f1c + f1s * bv[j]);   // NARROW: 32x32 -> 64
                        const i64 t3 = t2 * (i64)(V)(f2c + f2s * cv[j]);
                        rn += 1;
                        rq += (S)qv[j];
                        rb += (S)bv[j];
                        ra += (i64)av[j];
                        r2 += t2;
                        r3 += t3;
                    }
                }
            }
        }
    }
    flush();
    __syncthreads();
    // block reduction: warp w sums slots w, w+8, ... across the private copies
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = warp; v < G * LC_K; v += LC_THREADS / 32) {
        const int g = v / LC_K, k = v % LC_K;
        i64 s = 0;
        if (ACC32) {
            // slot k lives in a 32-bit plane (0 -> 0, 1 -> 1, 5 -> 2) or a 64-bit plane (2 -> 0, 3 -> 1, 4 -> 2)
            const bool small = k == 0 || k == 1 || k == 5;
            const int pl = k == 0 ? 0 : k == 1 ? 1 : k == 5 ? 2 : k - 2;
#pragma unroll
            for (int j = 0; j < LC_THREADS / 32; j++)
                s += small ? (i64)s32[(size_t)(g * 3 + pl) * LC_THREADS + lane + 32 * j] : (i64)s64[(size_t)(g * 3 + pl) * LC_THREADS + lane + 32 * j];
        } else {
#pragma unroll
            for (int j = 0; j < LC_THREADS / 32; j++) s += (i64)s64[(size_t)v * LC_THREADS + lane + 32 * j];
        }
        s = warp_sum(s);
        if (lane == 0) s_tot[v] = s;
    }
    __syncthreads();
    if (threadIdx.x < G * LC_K) {
        const int g = threadIdx.x / LC_K, k = threadIdx.x % LC_K;
        i64 s = s_tot[threadIdx.x];
        const i64 n = s_tot[g * LC_K];
        if (k == 1) s += n * p.q.base;            // stored -> logical sums
        else if (k == 2) s += n * p.A.base;
        else if (k == 5) s += n * p.B.base;
        partials[(i64)blockIdx.x * (G * LC_K) + threadIdx.x] = s;
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
// tiles are composed in order.  Only the rows after the crossing point are re-read (about 8%
// of lineitem for Q1's (N,O) sum_charge at SF100).
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
// the transducer of ONE addend x >= 0
__device__ __forceinline__ OrdState ord_of(i64 x)
{
    const i64 q = x / 10;
    const unsigned dgt = (unsigned)(x - q * 10), qb = (unsigned)(q & 1);
    OrdState r;
    r.sq = q;
    r.sx = x;
    const unsigned t0b = qb, t1b = qb ^ 1u;     // parity of S + q when entering even / odd
    r.c0 = (dgt > 5u) | ((dgt == 5u) & t0b);
    r.c1 = (dgt > 5u) | ((dgt == 5u) & t1b);
    r.p0 = t0b ^ r.c0;
    r.p1 = t1b ^ r.c1;
    return r;
}
__device__ __forceinline__ OrdState ord_shfl_down(const OrdState &st, int o)
{
    OrdState r;
    r.sq = __shfl_down_sync(0xffffffffu, st.sq, o);
    r.sx = __shfl_down_sync(0xffffffffu, st.sx, o);
    r.c0 = __shfl_down_sync(0xffffffffu, st.c0, o);
    r.c1 = __shfl_down_sync(0xffffffffu, st.c1, o);
    r.p0 = __shfl_down_sync(0xffffffffu, st.p0, o);
    r.p1 = __shfl_down_sync(0xffffffffu, st.p1, o);
    return r;
}

// 4 consecutive rows starting at `row` -> their ordered composite for (group, slot).  WIDE = false: every column is
// stored in <= 4 bytes (raw vectors of 16 bytes, 32-bit unpacking).
template <bool HAS_KEY1, bool WIDE = true>
__device__ __forceinline__ OrdState ord_quad_rows(const OrdParams &op, const uint8_t (*s_lut)[256], i64 row)
{
    typedef typename std::conditional<WIDE, i64, int>::type V;
    const LowcardParams &p = op.base;
    const i64 rem = p.nrows - row;
    Raw4<WIDE> d, a, b, c;
    ld_raw4(p.pred, row, d);
    const unsigned k0 = ld_stream4((const uint8_t *)p.key0.p + row), k1 = HAS_KEY1 ? ld_stream4((const uint8_t *)p.key1.p + row) : 0;
    ld_raw4(p.A, row, a);
    ld_raw4(p.B, row, b);
    ld_raw4(p.C, row, c);
    V dv[4], av[4], bv[4], cv[4];
    unpack4(p.pred, d, dv);
    unpack4(p.A, a, av);
    unpack4(p.B, b, bv);
    unpack4(p.C, c, cv);
    // Both entry parities are simulated directly (no per-row transducer objects): a row with digit d != 5 carries
    // [d > 5] whatever the state and flips both parities alike; a row with d == 5 carries parity(S + q) and leaves the
    // sum EVEN in both cases (half-to-even), after which the two simulations coincide.
    OrdState st = {0, 0, 0, 0, 0, 1};   // identity: parity preserved, no carries
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool ok = j < rem && dv[j] >= p.lo && dv[j] <= p.hi;
        int g = s_lut[0][(k0 >> (8 * j)) & 255];
        if (HAS_KEY1) g = g * p.n1 + s_lut[1][(k1 >> (8 * j)) & 255];
        if (ok && g == op.group) {
            const u64 x = (u64)ord_value(p, op.slot, av[j] + p.A.base, bv[j] + p.B.base, cv[j] + p.C.base);
            const u64 q = x / 10;
            const unsigned d = (unsigned)(x - q * 10), t = (unsigned)q & 1u;
            st.sq += (i64)q;
            st.sx += (i64)x;
            if (d == 5u) {
                st.c0 += st.p0 ^ t;
                st.c1 += st.p1 ^ t;
                st.p0 = 0;
                st.p1 = 0;
            } else {
                const unsigned c = d > 5u ? 1u : 0u;
                st.c0 += c;
                st.c1 += c;
                st.p0 ^= t ^ c;
                st.p1 ^= t ^ c;
            }
        }
    }
    return st;
}
// the 4 rows of one thread of one tile
template <bool HAS_KEY1>
__device__ __forceinline__ OrdState ord_thread_rows(const OrdParams &op, const uint8_t (*s_lut)[256], i64 tile)
{
    return ord_quad_rows<HAS_KEY1>(op, s_lut, tile * LC_TILE + threadIdx.x * SA_VEC);
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
            OrdState st = ord_thread_rows<HAS_KEY1>(op, s_lut, tile);
            // ordered composition across the warp, then across the 8 warps
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const OrdState r = ord_shfl_down(st, o);
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

// ---- the whole emulation on the device: no host round trip between the scan and the answer ----
// After the scan kernel every rank holds the per-CTA ORDERED partials (contig tiles) and, after one
// all-gather, every rank's exact totals.  Three small launches, identical on every rank:
//   ord_plan_kernel   which (group, slot) totals need 20 digits, which rank / CTA crosses 10^19
//   ord_jobs_kernel   transducer summaries of the tiles after the crossing (per tile inside the
//                     crossing CTA's range, per 16-tile chunk after it; later ranks: their whole shard)
//   ord_fold_kernel   crossing rank: finds the crossing tile and row, rounds there exactly as the
//                     reference's Add would, composes the rest in order; later ranks: compose their
//                     summaries into one transducer.  One 64-byte contribution per (job, rank).
// The contributions are all-gathered and applied in rank order on the host (a few dozen integer ops).
constexpr int ORD_MAXJOBS = 4;
constexpr int ORD_CHUNK = 4;            // tiles composed per summary after the crossing CTA's range
struct OrdJob {
    int g, s;
    int role;            // 0: this rank lies before the crossing, 1: the crossing rank, 2: a later rank
    int rstar;
    i64 tb, te;          // role 1: tile range of the crossing CTA
    u64 P;               // role 1: exact sum of everything before tile tb (< 10^19)
};
struct OrdContrib { u64 kind, s_lo, s_hi, q_lo, q_hi, c0, c1, p0p1; };   // kind 1: absolute state S; 2: transducer; 9: internal error
static_assert(sizeof(OrdContrib) == 64, "OrdContrib is exchanged as 64 bytes");

__device__ __forceinline__ u128 ord_thr() { return (u128)10000000000000000000ULL; }

// launched as <<<1, ORD_PLAN_THREADS>>>: the block stages the gathered totals and, for a crossing job, the CTAs' partials
// in shared memory (coalesced, all loads in flight at once); thread 0 then walks them -- the walks used to be chains of
// dependent global loads (135 us at SF100 for 296 CTAs).
constexpr int ORD_PLAN_THREADS = 256;
constexpr int ORD_PLAN_MAXG = 1024, ORD_PLAN_MAXCTA = 2048;
static __global__ void __launch_bounds__(ORD_PLAN_THREADS)
ord_plan_kernel(const u64 *__restrict__ gathered, i64 rank_words, int nranks, int myrank, int G, unsigned emu_mask,
                const i64 *__restrict__ part, int grid, i64 per, i64 ntiles, OrdJob *__restrict__ jobs, int *__restrict__ njobs)
{
    __shared__ u64 s_g[ORD_PLAN_MAXG];
    __shared__ i64 s_x[ORD_PLAN_MAXCTA];
    __shared__ OrdJob s_jobs[ORD_MAXJOBS];
    __shared__ int s_n;
    const u128 THR = ord_thr();
    const int GK = G * LC_K;
    const bool g_in_smem = (i64)nranks * 2 * GK <= ORD_PLAN_MAXG;
    if (g_in_smem)
        for (int i = threadIdx.x; i < nranks * 2 * GK; i += ORD_PLAN_THREADS) s_g[i] = gathered[(i64)(i / (2 * GK)) * rank_words + i % (2 * GK)];
    __syncthreads();
    auto total_of = [&](int r, int v) -> u128 {
        if (g_in_smem) return ((u128)s_g[r * 2 * GK + 2 * v + 1] << 64) | s_g[r * 2 * GK + 2 * v];
        return ((u128)gathered[r * rank_words + 2 * v + 1] << 64) | gathered[r * rank_words + 2 * v];
    };
    if (threadIdx.x == 0) {
        int n = 0;
        for (int g = 0; g < G; g++)
            for (int s = 2; s < LC_K; s++) {
                if (!((emu_mask >> s) & 1u)) continue;
                const int v = g * LC_K + s;
                u128 tot = 0;
                for (int r = 0; r < nranks; r++) tot += total_of(r, v);
                if (tot < THR) continue;
                if (n >= ORD_MAXJOBS) { n++; continue; }
                OrdJob j;
                j.g = g; j.s = s; j.tb = 0; j.te = 0; j.P = 0;
                u128 P = 0;
                int r = 0;
                for (; r < nranks; r++) {
                    const u128 t = total_of(r, v);
                    if (P + t >= THR) break;
                    P += t;
                }
                j.rstar = r;
                j.role = myrank == r ? 1 : myrank > r ? 2 : 0;
                j.P = (u64)P;                     // role 1: completed below with the CTAs before the crossing one
                s_jobs[n++] = j;
            }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    for (int e = 0; e < n && e < ORD_MAXJOBS; e++) {
        if (s_jobs[e].role != 1) continue;                       // block-uniform
        const int v = s_jobs[e].g * LC_K + s_jobs[e].s;
        const bool x_in_smem = grid <= ORD_PLAN_MAXCTA;
        if (x_in_smem) for (int c = threadIdx.x; c < grid; c += ORD_PLAN_THREADS) s_x[c] = part[(i64)c * GK + v];
        __syncthreads();
        if (threadIdx.x == 0) {
            u128 P = s_jobs[e].P;
            int c = 0;
            for (; c < grid; c++) {
                const u128 x = (u128)(u64)(x_in_smem ? s_x[c] : part[(i64)c * GK + v]);
                if (P + x >= THR) break;
                P += x;
            }
            if (c == grid) c = grid - 1;        // cannot happen (the totals say the crossing is here); keeps indices sane
            s_jobs[e].tb = (i64)c * per;
            s_jobs[e].te = s_jobs[e].tb + per < ntiles ? s_jobs[e].tb + per : ntiles;
            s_jobs[e].P = (u64)P;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int e = 0; e < n && e < ORD_MAXJOBS; e++) jobs[e] = s_jobs[e];
        *njobs = n;
    }
}

// ordered block-wide composition of one OrdState per thread (thread order); result valid in thread 0
__device__ __forceinline__ OrdState ord_block_compose(OrdState st, OrdState *s_w /* [LC_THREADS/32] */)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const OrdState r = ord_shfl_down(st, o);
        if ((threadIdx.x & 31) + o < 32) st = ord_compose(st, r);
    }
    __syncthreads();                       // s_w may still be read from a previous call
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = st;
    __syncthreads();
    OrdState run = {0, 0, 0, 0, 0, 1};
    if (threadIdx.x == 0)
        for (int w = 0; w < LC_THREADS / 32; w++) run = ord_compose(run, s_w[w]);
    return run;
}
__device__ __forceinline__ OrdState ord_from_summary(const OrdSummary &o)
{
    OrdState r;
    r.sq = o.sum_q; r.sx = o.sum_x; r.c0 = o.c0; r.c1 = o.c1; r.p0 = o.p0; r.p1 = o.p1;
    return r;
}

// summaries of every job's tiles: work item w < ntile1 -> tile tb + w alone; w >= ntile1 -> chunk of ORD_CHUNK tiles.
// One WARP per work item: its lanes compose their 4 rows, a shuffle ladder composes the 128-row group in lane
// order, lane 0 carries the running composite over the item's groups -- no block-wide synchronisation.
template <bool HAS_KEY1, bool WIDE>
__global__ void __launch_bounds__(LC_THREADS, 2)
ord_jobs_kernel(const LowcardParams p, const OrdJob *__restrict__ jobs, const int *__restrict__ njobs, OrdSummary *__restrict__ out,
                i64 job_stride, i64 ntiles)
{
    __shared__ uint8_t s_lut[2][256];
    const int nj = *njobs < ORD_MAXJOBS ? *njobs : ORD_MAXJOBS;
    if (nj == 0) return;
    for (int i = threadIdx.x; i < 512; i += LC_THREADS) s_lut[i >> 8][i & 255] = p.luts[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = LC_THREADS / 32;
    for (int e = 0; e < nj; e++) {
        const OrdJob job = jobs[e];
        if (job.role == 0) continue;
        OrdParams op;
        op.base = p;
        op.group = job.g;
        op.slot = job.s;
        const i64 ntile1 = job.te - job.tb;
        const i64 nch = (ntiles - job.te + ORD_CHUNK - 1) / ORD_CHUNK;
        for (i64 w = (i64)blockIdx.x * nw + warp; w < ntile1 + nch; w += (i64)gridDim.x * nw) {
            i64 t0, t1;
            if (w < ntile1) { t0 = job.tb + w; t1 = t0 + 1; }
            else { t0 = job.te + (w - ntile1) * ORD_CHUNK; t1 = t0 + ORD_CHUNK < ntiles ? t0 + ORD_CHUNK : ntiles; }
            OrdState run = {0, 0, 0, 0, 0, 1};
            // 256 rows per step: a lane takes 8 CONSECUTIVE rows (two quads, their loads in flight together) and
            // composes them in order before the ladder
            for (i64 row = t0 * LC_TILE; row < t1 * LC_TILE; row += 256) {
                const i64 r0 = row + lane * 8;
                const OrdState q0 = ord_quad_rows<HAS_KEY1, WIDE>(op, s_lut, r0), q1 = ord_quad_rows<HAS_KEY1, WIDE>(op, s_lut, r0 + 4);
                OrdState st = ord_compose(q0, q1);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const OrdState r = ord_shfl_down(st, o);
                    if (lane + o < 32) st = ord_compose(st, r);
                }
                if (lane == 0) run = ord_compose(run, st);
            }
            if (lane == 0) {
                OrdSummary o;
                o.sum_q = run.sq; o.sum_x = run.sx; o.c0 = run.c0; o.c1 = run.c1; o.p0 = run.p0; o.p1 = run.p1;
                out[(i64)e * job_stride + w] = o;
            }
        }
    }
}

// The same summaries for the common narrow case (every column stored in <= 4 bytes, non-negative values,
// A*(c1+s1*B) < 2^32 -- what the staged scan kernel requires too): 32-bit arithmetic up to the 64-bit addend, and NO
// shuffle ladder.  A lane simulates its 8 consecutive rows for both entering parities; the 32 lane results are
// composed in lane order with two ballots and bit arithmetic: a lane that saw a digit-5 row ends EVEN-or-odd
// whatever it entered with (a "reset"), any other lane just flips or keeps the parity, so the parity entering
// lane l is the exit parity of the last reset lane below l XOR the flips in between.  The order-independent parts
// (sum of floor(x/10), of x, of the carries) stay in lane-local accumulators until the work item ends.
struct OrdFastParams {
    NCol pred, key0, key1, A, B, C;
    unsigned p_lo, p_span;
    int abase, bbase, f1c, f1s, f2c, f2s;      // f1 = f1c + f1s * stored B, f2 = f2c + f2s * stored C
    int has_key1, n1;
    unsigned char code0[LC_MAXG], code1[LC_MAXG];   // dense id -> byte code, per key
    i64 nrows;
};

struct OrdFastRaw { Raw4<false> d, a, b, c; unsigned k0, k1; };
__device__ __forceinline__ void ord_fast_load(const OrdFastParams &p, i64 row, OrdFastRaw &r)
{
    ld_raw4(p.pred, row, r.d);
    r.k0 = ld_stream4((const uint8_t *)p.key0.p + row);
    r.k1 = p.has_key1 ? ld_stream4((const uint8_t *)p.key1.p + row) : 0u;
    ld_raw4(p.A, row, r.a);
    ld_raw4(p.B, row, r.b);
    ld_raw4(p.C, row, r.c);
}
__device__ __forceinline__ void ord_fast_quad(const OrdFastParams &p, int slot, unsigned target, i64 row, const OrdFastRaw &r, u64 &sq, u64 &sx,
                                              unsigned &c0, unsigned &c1, unsigned &p0, unsigned &p1)
{
    const i64 rem = p.nrows - row;
    int dv[4], av[4], bv[4], cv[4];
    unpack4(p.pred, r.d, dv);
    unpack4(p.A, r.a, av);
    unpack4(p.B, r.b, bv);
    unpack4(p.C, r.c, cv);
    const unsigned k0 = r.k0, k1 = r.k1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const unsigned kk = ((k0 >> (8 * j)) & 255u) | (((k1 >> (8 * j)) & 255u) << 8);
        const bool ok = j < rem && ((unsigned)dv[j] - p.p_lo) <= p.p_span && kk == target;
        if (ok) {
            const unsigned al = (unsigned)(av[j] + p.abase);
            const unsigned t2 = al * (unsigned)(p.f1c + p.f1s * bv[j]);
            // x = 10 q + dg.  No 64-bit division: the digit comes from residues (x = t2 * f2: dg = (t2 mod 10)(f2 mod 10) mod 10),
            // the parity of q from the low word ((x - dg) / 2 = 5 q has q's parity), and sum(q) = (sum(x) - sum(dg)) / 10 once per item.
            u64 x;
            unsigned xlo, dg;
            if (slot == 4) {
                const unsigned f2 = (unsigned)(p.f2c + p.f2s * cv[j]);
                x = (u64)t2 * f2;
                xlo = t2 * f2;
                dg = ((t2 % 10u) * (f2 % 10u)) % 10u;
            } else {
                xlo = slot == 3 ? t2 : slot == 2 ? al : (unsigned)(bv[j] + p.bbase);
                x = xlo;
                dg = xlo % 10u;
            }
            const unsigned t = ((xlo - dg) >> 1) & 1u;
            sq += dg;                 // the caller turns (sum x, sum digits) into sum q
            sx += x;
            if (dg == 5u) {
                c0 += p0 ^ t;
                c1 += p1 ^ t;
                p0 = 0;
                p1 = 0;
            } else {
                const unsigned cy = dg > 5u ? 1u : 0u;
                c0 += cy;
                c1 += cy;
                p0 ^= t ^ cy;
                p1 ^= t ^ cy;
            }
        }
    }
}

static __global__ void __launch_bounds__(LC_THREADS, 3)
ord_jobs_fast_kernel(const OrdFastParams p, const OrdJob *__restrict__ jobs, const int *__restrict__ njobs, OrdSummary *__restrict__ out,
                     i64 job_stride, i64 ntiles)
{
    const int nj = *njobs < ORD_MAXJOBS ? *njobs : ORD_MAXJOBS;
    if (nj == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = LC_THREADS / 32;
    const unsigned lt = (1u << lane) - 1u;
    for (int e = 0; e < nj; e++) {
        const OrdJob job = jobs[e];
        if (job.role == 0) continue;
        const int i0 = p.has_key1 ? job.g / p.n1 : job.g, i1 = p.has_key1 ? job.g % p.n1 : 0;
        const unsigned target = (unsigned)p.code0[i0] | (p.has_key1 ? (unsigned)p.code1[i1] << 8 : 0u);
        const i64 ntile1 = job.te - job.tb;
        const i64 nch = (ntiles - job.te + ORD_CHUNK - 1) / ORD_CHUNK;
        for (i64 w = (i64)blockIdx.x * nw + warp; w < ntile1 + nch; w += (i64)gridDim.x * nw) {
            i64 t0, t1;
            if (w < ntile1) { t0 = job.tb + w; t1 = t0 + 1; }
            else { t0 = job.te + (w - ntile1) * ORD_CHUNK; t1 = t0 + ORD_CHUNK < ntiles ? t0 + ORD_CHUNK : ntiles; }
            u64 sq = 0, sx = 0;
            unsigned lc0 = 0, lc1 = 0;          // this lane's carries given the ITEM was entered even / odd
            unsigned P0 = 0, P1 = 1;            // running exit parity of the item for both entries (warp-uniform)
            for (i64 row = t0 * LC_TILE; row < t1 * LC_TILE; row += 256) {
                unsigned c0 = 0, c1 = 0, q0 = 0, q1 = 1;
                const i64 r0 = row + lane * 8;
                OrdFastRaw ra, rb;
                ord_fast_load(p, r0, ra);
                ord_fast_load(p, r0 + 4, rb);
                ord_fast_quad(p, job.s, target, r0, ra, sq, sx, c0, c1, q0, q1);
                ord_fast_quad(p, job.s, target, r0 + 4, rb, sq, sx, c0, c1, q0, q1);
                const bool reset = q0 == q1;
                const unsigned R = __ballot_sync(0xffffffffu, reset), F = __ballot_sync(0xffffffffu, q0 != 0u);
                // parity entering this lane, for the step entered even (E0) / odd (E1)
                unsigned E0, E1;
                const unsigned belowR = R & lt;
                if (belowR) {
                    const int m = 31 - __clz(belowR);
                    const unsigned between = F & ~R & lt & ~((2u << m) - 1u);
                    E0 = E1 = ((F >> m) & 1u) ^ (__popc(between) & 1u);
                } else {
                    E0 = __popc(F & lt) & 1u;
                    E1 = E0 ^ 1u;
                }
                const unsigned cE0 = E0 ? c1 : c0, cE1 = E1 ? c1 : c0;
                lc0 += P0 ? cE1 : cE0;
                lc1 += P1 ? cE1 : cE0;
                // exit parity of the step
                unsigned G0, G1;
                if (R) {
                    const int m = 31 - __clz(R);
                    G0 = G1 = ((F >> m) & 1u) ^ (__popc(F & ~R & ~((2u << m) - 1u)) & 1u);
                } else {
                    G0 = __popc(F) & 1u;
                    G1 = G0 ^ 1u;
                }
                P0 = P0 ? G1 : G0;
                P1 = P1 ? G1 : G0;
            }
            const i64 td = warp_sum((i64)sq), tx = warp_sum((i64)sx);
            const i64 tc0 = warp_sum((i64)lc0), tc1 = warp_sum((i64)lc1);
            if (lane == 0) {
                OrdSummary o;
                o.sum_q = (tx - td) / 10; o.sum_x = tx; o.c0 = (unsigned)tc0; o.c1 = (unsigned)tc1; o.p0 = P0; o.p1 = P1;
                out[(i64)e * job_stride + w] = o;
            }
        }
    }
}

// one CTA per job: this rank's 64-byte contribution
template <bool HAS_KEY1>
__global__ void __launch_bounds__(LC_THREADS)
ord_fold_kernel(const LowcardParams p, const OrdJob *__restrict__ jobs, const int *__restrict__ njobs, const OrdSummary *__restrict__ sums,
                i64 job_stride, i64 ntiles, OrdContrib *__restrict__ contrib /* [ORD_MAXJOBS] */)
{
    __shared__ uint8_t s_lut[2][256];
    __shared__ OrdState s_w[LC_THREADS / 32];
    __shared__ u64 s_slice[LC_THREADS];
    __shared__ i64 s_tstar;
    __shared__ u64 s_P, s_S0;
    __shared__ int s_cross_thread, s_cross_j;
    const int e = blockIdx.x;
    const int nj = *njobs < ORD_MAXJOBS ? *njobs : ORD_MAXJOBS;
    OrdContrib mine;
    mine.kind = 0; mine.s_lo = mine.s_hi = mine.q_lo = mine.q_hi = mine.c0 = mine.c1 = mine.p0p1 = 0;
    if (e >= nj || jobs[e].role == 0) {
        if (threadIdx.x == 0) contrib[e] = mine;
        return;
    }
    const OrdJob job = jobs[e];
    const OrdSummary *sm = sums + (i64)e * job_stride;
    const u128 THR = ord_thr();
    const i64 ntile1 = job.te - job.tb;
    const i64 nsum = ntile1 + (ntiles - job.te + ORD_CHUNK - 1) / ORD_CHUNK;
    // ordered composition of summaries [from, nsum): contiguous slices per thread, then across the block
    auto compose_from = [&](i64 from) {
        const i64 n = nsum - from;
        const i64 L = (n + LC_THREADS - 1) / LC_THREADS;
        OrdState st = {0, 0, 0, 0, 0, 1};
        const i64 i0 = from + (i64)threadIdx.x * L, i1 = i0 + L < nsum ? i0 + L : nsum;
        i64 i = i0;
        for (; i + 4 <= i1; i += 4) {           // four independent 32-byte loads in flight, composed in order
            const OrdSummary a = sm[i], b = sm[i + 1], c = sm[i + 2], d = sm[i + 3];
            st = ord_compose(st, ord_compose(ord_compose(ord_from_summary(a), ord_from_summary(b)), ord_compose(ord_from_summary(c), ord_from_summary(d))));
        }
        for (; i < i1; i++) st = ord_compose(st, ord_from_summary(sm[i]));
        return ord_block_compose(st, s_w);
    };
    if (job.role == 2) {
        const OrdState all = compose_from(0);
        if (threadIdx.x == 0) {
            mine.kind = 2;
            mine.q_lo = (u64)all.sq;                 // sum of floor(x/10) over a shard: < 2^63
            mine.q_hi = 0;
            mine.c0 = all.c0; mine.c1 = all.c1;
            mine.p0p1 = (u64)all.p0 | ((u64)all.p1 << 1);
            contrib[e] = mine;
        }
        return;
    }
    for (int i = threadIdx.x; i < 512; i += LC_THREADS) s_lut[i >> 8][i & 255] = p.luts[i];
    // 1. the crossing tile: slice sums of the per-tile exact sums, then a walk inside the crossing slice
    const i64 L1 = (ntile1 + LC_THREADS - 1) / LC_THREADS;
    u64 mysum = 0;
    for (i64 i = (i64)threadIdx.x * L1; i < ((i64)threadIdx.x + 1) * L1 && i < ntile1; i++) mysum += (u64)sm[i].sum_x;
    s_slice[threadIdx.x] = mysum;
    if (threadIdx.x == 0) s_tstar = -1;
    __syncthreads();
    if (threadIdx.x == 0) {
        u128 P = job.P;
        int t = 0;
        for (; t < LC_THREADS; t++) {
            if (P + s_slice[t] >= THR) break;
            P += s_slice[t];
        }
        if (t < LC_THREADS) {
            for (i64 i = (i64)t * L1; i < ((i64)t + 1) * L1 && i < ntile1; i++) {
                if (P + (u64)sm[i].sum_x >= THR) { s_tstar = i; break; }
                P += (u64)sm[i].sum_x;
            }
        }
        s_P = (u64)P;
        s_cross_thread = -1;
    }
    __syncthreads();
    const i64 tstar = s_tstar;
    if (tstar < 0) {
        if (threadIdx.x == 0) { mine.kind = 9; contrib[e] = mine; }
        return;
    }
    // 2. that tile row by row: exact prefix up to the crossing row, rounded there exactly as the reference's Add
    //    would; the rows after it compose as transducers
    OrdParams op;
    op.base = p;
    op.group = job.g;
    op.slot = job.s;
    const i64 tile = job.tb + tstar;
    i64 xs[4];
    bool in[4];
    {
        const i64 row = tile * LC_TILE + threadIdx.x * SA_VEC;
        const i64 rem = p.nrows - row;
        Raw4<true> d, a, b, c;
        ld_raw4(p.pred, row, d);
        const unsigned k0 = ld_stream4((const uint8_t *)p.key0.p + row), k1 = HAS_KEY1 ? ld_stream4((const uint8_t *)p.key1.p + row) : 0;
        ld_raw4(p.A, row, a);
        ld_raw4(p.B, row, b);
        ld_raw4(p.C, row, c);
        i64 dv[4], av[4], bv[4], cv[4];
        unpack4(p.pred, d, dv);
        unpack4(p.A, a, av);
        unpack4(p.B, b, bv);
        unpack4(p.C, c, cv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int g = s_lut[0][(k0 >> (8 * j)) & 255];
            if (HAS_KEY1) g = g * p.n1 + s_lut[1][(k1 >> (8 * j)) & 255];
            in[j] = j < rem && dv[j] >= p.lo && dv[j] <= p.hi && g == job.g;
            xs[j] = in[j] ? ord_value(p, job.s, av[j] + p.A.base, bv[j] + p.B.base, cv[j] + p.C.base) : 0;
        }
    }
    s_slice[threadIdx.x] = (u64)(xs[0] + xs[1] + xs[2] + xs[3]);
    __syncthreads();
    if (threadIdx.x == 0) {                       // exclusive scan of 256 thread sums (in place)
        u64 run = 0;
        for (int t = 0; t < LC_THREADS; t++) { const u64 v = s_slice[t]; s_slice[t] = run; run += v; }
    }
    __syncthreads();
    {
        u128 P = (u128)s_P + s_slice[threadIdx.x];
        if (P < THR) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!in[j]) continue;
                P += (u64)xs[j];
                if (P >= THR) {                    // exactly one thread / row gets here
                    const u128 q = P / 10;
                    const unsigned r = (unsigned)(P - q * 10);
                    s_S0 = (u64)(q + ((r > 5 || (r == 5 && (q & 1))) ? 1 : 0));
                    s_cross_thread = (int)threadIdx.x;
                    s_cross_j = j;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (s_cross_thread < 0) {
        if (threadIdx.x == 0) { mine.kind = 9; contrib[e] = mine; }
        return;
    }
    OrdState st = {0, 0, 0, 0, 0, 1};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool after = (int)threadIdx.x > s_cross_thread || ((int)threadIdx.x == s_cross_thread && j > s_cross_j);
        if (in[j] && after) st = ord_compose(st, ord_of(xs[j]));
    }
    const OrdState t_tile = ord_block_compose(st, s_w);
    // 3. everything after that tile, in order
    const OrdState t_rest = compose_from(tstar + 1);
    if (threadIdx.x == 0) {
        u64 S = s_S0;
        S += (u64)t_tile.sq + ((S & 1) ? t_tile.c1 : t_tile.c0);
        S += (u64)t_rest.sq + ((S & 1) ? t_rest.c1 : t_rest.c0);
        mine.kind = 1;
        mine.s_lo = S;
        mine.s_hi = 0;
        contrib[e] = mine;
    }
}

// ------------------------------------------------------------------------------
// Shape-agnostic scan-aggregate: any conjunction of range / code-set / string predicates on typed
// columns, up to GEN_MAXACC aggregates (sum / min / max of a product of <= 3 affine factors, or
// count), grouped by 0..2 byte-coded keys, NULL-aware.  Same thread-private shared-memory group
// tables as the chain kernel.  Columns, predicates and factors are runtime descriptors, but a thread
// still handles 4 consecutive rows with ONE vector load per column (the descriptor loops run once per
// 4 rows, the row loops are unrolled), so every shape the lowering accepts streams its columns the
// way the specialised kernels do.
// ------------------------------------------------------------------------------
constexpr int GEN_MAXPRED = 8, GEN_MAXACC = 8, GEN_MAXFAC = 3;
enum { GEN_SUM = 0, GEN_MIN = 1, GEN_MAX = 2, GEN_COUNTV = 3 };   // COUNTV: count(column) = rows where it is not NULL

struct GenCol { NCol c; const uint8_t *valid; };   // valid: packed bits, 1 = not NULL; null = no NULLs
// validity of the 4 rows starting at row (a multiple of 4): low nibble, bit j = row + j is not NULL
__device__ __forceinline__ unsigned gen_valid4(const GenCol &c, i64 row)
{
    if (c.valid == nullptr) return 0xfu;
    return ((unsigned)__ldg(c.valid + (row >> 3)) >> (row & 4)) & 0xfu;
}

struct GenAcc {
    int kind, nfac;
    GenCol fac[GEN_MAXFAC];
    i64 c[GEN_MAXFAC];           // on the STORED value: factor = c + s * stored (the host folded the column base in)
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
    i64 plo[GEN_MAXPRED], phi[GEN_MAXPRED];     // STORED-domain bounds
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
    constexpr i64 TILE = (i64)NT * SA_VEC;
    const i64 ntiles = (p.nrows + TILE - 1) / TILE;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 row = tile * TILE + threadIdx.x * SA_VEC;
        const i64 rem = p.nrows - row;
        unsigned ok = rem >= 4 ? 0xfu : rem <= 0 ? 0u : (1u << rem) - 1u;      // bit j: row + j is selected so far
        for (int k = 0; k < p.npred && ok; k++) {
            if (NULLS) ok &= gen_valid4(p.pcol[k], row);
            Raw4<true> r;
            ld_raw4(p.pcol[k].c, row, r);
            i64 v[4];
            unpack4(p.pcol[k].c, r, v);
            if (p.pset[k]) {
#pragma unroll
                for (int j = 0; j < 4; j++) if (!((p.pmask[k][(v[j] >> 5) & 7] >> (v[j] & 31)) & 1u)) ok &= ~(1u << j);
            } else {
                const i64 lo = p.plo[k], hi = p.phi[k];
#pragma unroll
                for (int j = 0; j < 4; j++) if (v[j] < lo || v[j] > hi) ok &= ~(1u << j);
            }
        }
        for (int k = 0; k < p.nlike && ok; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) if (((ok >> j) & 1u) && !gen_like_pass(p.like[k], row + j)) ok &= ~(1u << j);
        if (!ok) continue;
        unsigned k0 = 0, k1 = 0;
        if (p.nkeys > 0) k0 = ld_stream4(p.key0 + row);
        if (p.nkeys > 1) k1 = ld_stream4(p.key1 + row);
        i64 *t[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int g = 0;
            if (p.nkeys > 0) g = s_lut[0][(k0 >> (8 * j)) & 255];
            if (p.nkeys > 1) g = g * p.n1 + s_lut[1][(k1 >> (8 * j)) & 255];
            t[j] = my + (i64)g * P * NT;
            if ((ok >> j) & 1u) {
                const i64 n = t[j][0];
                if (n == 0) atomicMin((long long *)&s_first[g], (long long)(p.row_base + row + j));
                t[j][0] = n + 1;
            }
        }
        for (int a = 0; a < p.nacc; a++) {
            const GenAcc &A = p.acc[a];
            i64 x[4] = {1, 1, 1, 1};
            unsigned vok = ok;
            for (int f = 0; f < A.nfac; f++) {
                if (NULLS) vok &= gen_valid4(A.fac[f], row);
                Raw4<true> r;
                ld_raw4(A.fac[f].c, row, r);
                i64 v[4];
                unpack4(A.fac[f].c, r, v);
#pragma unroll
                for (int j = 0; j < 4; j++) x[j] *= A.c[f] + A.s[f] * v[j];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!((vok >> j) & 1u)) continue;
                i64 *slot = t[j] + (i64)(a + 1) * NT;
                const i64 cur = *slot;
                *slot = A.kind == GEN_SUM ? cur + x[j] : A.kind == GEN_MIN ? (x[j] < cur ? x[j] : cur) : A.kind == GEN_MAX ? (x[j] > cur ? x[j] : cur) : cur + 1;
                if (NULLS) t[j][(i64)(p.nacc + 1 + a) * NT] += 1;
            }
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
// launched as <<<nvals, 32>>>: one warp per value, lanes stride over the CTAs' partials
static __global__ void finalize128_kernel(const i64 *__restrict__ partials, int nblocks, int nvals,
                                   u64 *__restrict__ out /* [nvals][2] */)
{
    const int v = blockIdx.x, lane = threadIdx.x;
    if (v >= nvals) return;
    u64 lo = 0;
    i64 hi = 0;
    for (int b = lane; b < nblocks; b += 32) {
        const i64 x = partials[(i64)b * nvals + v];
        const u64 nlo = lo + (u64)x;
        hi += (x < 0 ? -1 : 0) + (nlo < lo ? 1 : 0);
        lo = nlo;
    }
    // 128-bit warp reduction: add (lo, hi) pairs with carry
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 olo = __shfl_xor_sync(0xffffffffu, lo, o);
        const i64 ohi = __shfl_xor_sync(0xffffffffu, hi, o);
        const u64 nlo = lo + olo;
        hi += ohi + (nlo < lo ? 1 : 0);
        lo = nlo;
    }
    if (lane == 0) { out[2 * v] = lo; out[2 * v + 1] = (u64)hi; }
}

}  // namespace pg

#include "scan_staged.cuh"
