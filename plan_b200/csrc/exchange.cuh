// exchange.cuh -- all-to-all hash-partitioned ROW shuffle for a join whose two sides are sharded on
// different keys (BASELINE config 5: TPC-H Q9's lineitem x partsupp on (partkey, suppkey)).
//
// Reference semantics kept: JoinHashTable.Build / Finalize (/root/reference/pkg/compute/join_table.go:85-288)
// and Scan.Next / InnerJoin (join_scan.go:182-299) -- an INNER equi-join emits one row per (probe row, matching
// build row) pair.  Which GPU forms a pair is irrelevant to the result, so both sides travel to the OWNER of the
// key, rank mix64(key) % world, over NVLink (grouped ncclSend/ncclRecv), and the join runs there:
//
//   build rows   [key, x_0 .. x_{nx-1}]            x_i = the build columns the aggregate reads
//   probe rows   [key, group, a_0 .. a_{nterm-1}]  a_t = mul_t * product of term t's factors that do NOT come
//                                                  from the exchanged build side (evaluated where the fact row lives,
//                                                  together with every other join of the star and the dense group id)
//
// The owner inserts the received build rows into a bucketized table (payload = index into the receive buffer),
// probes it with the received probe rows, multiplies the deferred factors in and adds into a block-private
// shared-memory table of the dense groups -- the same sink as hits_star_kernel.  Records are made once
// (x_build_rows_kernel / star_pre_kernel), counted per destination in the same pass, and moved into destination
// order by x_scatter_kernel (warp-aggregated cursors) so the exchange itself is W contiguous sends.
#pragma once
#include "join.cuh"

namespace pg {

constexpr int X_MAXCOL = 3;            // build columns carried to the owner
constexpr int X_MAXWORLD = 64;
constexpr unsigned char X_DROP = 0xff;

__device__ __forceinline__ int x_dest(i64 key, int world) { return (int)(mix64((u64)key ^ 0xA24BAED4963EE407ULL) % (u64)world); }

// per-block destination histogram -> global counts
__device__ __forceinline__ void x_flush_counts(unsigned *s_c, int world, unsigned long long *cnt)
{
    __syncthreads();
    if ((int)threadIdx.x < world && s_c[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}

struct XBuildParams {
    const unsigned *hits;                 // build rows that pass the build side's own filters (filter pass)
    const unsigned long long *hit_count;
    int nkey;
    TypedCol key[2];
    int nx;
    TypedCol x[X_MAXCOL];
    int world;
    i64 *rec;                             // [hits][1 + nx]
    unsigned char *dest;                  // [hits]
    unsigned long long *cnt;              // [world]
};

static __global__ void __launch_bounds__(256)
x_build_rows_kernel(const XBuildParams p)
{
    __shared__ unsigned s_c[X_MAXWORLD];
    if (threadIdx.x < X_MAXWORLD) s_c[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long n = *p.hit_count;
    const int RW = 1 + p.nx;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const i64 row = (i64)p.hits[i];
        i64 key = load_typed(p.key[0], row);
        if (p.nkey == 2) key = (key << 32) | (load_typed(p.key[1], row) & 0xffffffffLL);
        i64 *r = p.rec + i * RW;
        r[0] = key;
        for (int c = 0; c < p.nx; c++) r[1 + c] = load_typed(p.x[c], row);
        const int d = x_dest(key, p.world);
        p.dest[i] = (unsigned char)d;
        atomicAdd(&s_c[d], 1u);
    }
    x_flush_counts(s_c, p.world, p.cnt);
}

// move records into destination order: cursor[d] starts at the d-th send offset (rows).  Lanes of a warp heading
// for the same rank take consecutive places with ONE atomic (match.any), so the cursors see <= world updates per warp.
static __global__ void __launch_bounds__(256)
x_scatter_kernel(const i64 *__restrict__ rec, const unsigned char *__restrict__ dest, const unsigned long long *n_ptr, int RW,
                 unsigned long long *cursor, i64 *__restrict__ send)
{
    const unsigned long long n = *n_ptr;
    const unsigned lane = threadIdx.x & 31;
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < n; i0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + lane;
        const int d = i < n ? (int)dest[i] : (int)X_DROP;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d == (int)X_DROP) continue;
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(&cursor[d], (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const unsigned long long pos = base + __popc(peers & ((1u << lane) - 1u));
        const i64 *src = rec + i * RW;
        i64 *dst = send + pos * RW;
        if ((RW & 1) == 0) {        // records of an even number of words are 16-byte aligned on both sides: 16-byte moves
            for (int w = 0; w < RW; w += 2) *(longlong2 *)(dst + w) = __ldg((const longlong2 *)(src + w));
        } else {
            for (int w = 0; w < RW; w++) dst[w] = src[w];
        }
    }
}

static __global__ void __launch_bounds__(256)
x_insert_kernel(const JoinTable jt, const i64 *__restrict__ rows, i64 n, int RW)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) jt_insert(jt, rows[i * RW], (u64)i);
}

// ---- probe side, where the fact rows live ----
struct XPreParams {
    int xl;                       // the exchanged lookup: skipped here, its key is the record's key
    int world;
    int term_x[STAR_MAXTERM];     // bit f set: factor f of the term reads the exchanged build side (deferred to the owner)
    i64 *rec;                     // [hits][2 + nterm]
    unsigned char *dest;          // [hits], X_DROP when another join of the star rejects the row
    unsigned long long *cnt;      // [world]
};

static __global__ void __launch_bounds__(256)
star_pre_kernel(const StarParams sp, const XPreParams xp)
{
    __shared__ unsigned s_c[X_MAXWORLD];
    if (threadIdx.x < X_MAXWORLD) s_c[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long n = *sp.hit_count;
    const int RW = 2 + sp.nterm;
    unsigned long long n_multi = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        i64 rows[STAR_MAXLOOKUP + 1];
        rows[0] = (i64)sp.hits[i];
        auto val = [&](const ValRef &r) { return load_typed(r.col, rows[r.from_build]); };
        bool ok = true;
        i64 xkey = 0;
#pragma unroll
        for (int l = 0; l < STAR_MAXLOOKUP; l++) {
            if (l >= sp.nlookup || !ok) continue;
            const StarLookup &L = sp.lk[l];
            i64 key = val(L.key[0]);
            if (L.nkey == 2) key = (key << 32) | (val(L.key[1]) & 0xffffffffLL);
            rows[l + 1] = -1;
            if (l == xp.xl) { xkey = key; continue; }
            if (L.existence) {
                ok = bitmap_test(L.jt, key);
            } else {
                int matches = 0;
                jt_probe(L.jt, key, [&](u64 r) { rows[l + 1] = (i64)r; matches++; });
                ok = matches > 0;
                n_multi += matches > 1 ? 1 : 0;
            }
        }
        if (!ok) { xp.dest[i] = X_DROP; continue; }
        int g = 0;
#pragma unroll
        for (int k = 0; k < STAR_MAXPART; k++) {
            if (k >= sp.nparts) continue;
            i64 v = val(sp.part[k].v);
            if (sp.part[k].fn == 1) v = year_of_days(v);
            g = g * sp.part[k].n + (int)(v - sp.part[k].lo);
        }
        i64 *r = xp.rec + i * RW;
        r[0] = xkey;
        r[1] = (i64)g;
#pragma unroll
        for (int t = 0; t < STAR_MAXTERM; t++) {
            if (t >= sp.nterm) continue;
            i64 x = sp.term[t].mul;
            for (int f = 0; f < sp.term[t].nfac; f++)
                if (!((xp.term_x[t] >> f) & 1)) x *= sp.term[t].fc[f] + sp.term[t].fs[f] * val(sp.term[t].fac[f]);
            r[2 + t] = x;
        }
        const int d = x_dest(xkey, xp.world);
        xp.dest[i] = (unsigned char)d;
        atomicAdd(&s_c[d], 1u);
    }
    x_flush_counts(s_c, xp.world, xp.cnt);
    n_multi = (unsigned long long)warp_sum((i64)n_multi);
    if ((threadIdx.x & 31) == 0 && n_multi) atomicAdd(&sp.counters[2], n_multi);
}

// ---- owner side: received probe rows x received build rows -> dense groups ----
struct XPostParams {
    const i64 *prow;              // [np][2 + nterm]
    i64 np;
    const i64 *brow;              // [nb][1 + nx]
    int nx;
    JoinTable jt;                 // key -> index into brow
    int nterm;
    int nxf[STAR_MAXTERM];        // deferred factors of each term: (xfc + xfs * brow[xcol])
    int xcol[STAR_MAXTERM][3];
    i64 xfc[STAR_MAXTERM][3];
    int xfs[STAR_MAXTERM][3];
    int ngroups;
    unsigned long long *gsum, *gsum_hi, *gcnt;
    unsigned long long *counters; // [1] joined rows, [2] probe rows with more than one match
};

static __global__ void __launch_bounds__(256)
star_post_kernel(const XPostParams p)
{
    extern __shared__ unsigned long long s_xstar[];
    unsigned long long *s_sum = s_xstar, *s_cnt = s_xstar + p.ngroups;
    for (int i = threadIdx.x; i < 2 * p.ngroups; i += blockDim.x) s_xstar[i] = 0;
    __syncthreads();
    const int RWp = 2 + p.nterm, RWb = 1 + p.nx;
    unsigned long long n_join = 0, n_multi = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < p.np; i += (i64)gridDim.x * blockDim.x) {
        const i64 *r = p.prow + i * RWp;
        const i64 key = __ldg(r);
        i64 b = -1;
        int matches = 0;
        jt_probe(p.jt, key, [&](u64 x) { b = (i64)x; matches++; });
        if (matches == 0) continue;
        n_multi += matches > 1 ? 1 : 0;
        n_join++;
        const int g = (int)__ldg(r + 1);
        const i64 *br = p.brow + b * RWb;
        i64 amount = 0;
#pragma unroll
        for (int t = 0; t < STAR_MAXTERM; t++) {
            if (t >= p.nterm) continue;
            i64 x = __ldg(r + 2 + t);
            for (int f = 0; f < p.nxf[t]; f++) x *= p.xfc[t][f] + p.xfs[t][f] * __ldg(br + 1 + p.xcol[t][f]);
            amount += x;
        }
        atomicAdd(&s_sum[g], (unsigned long long)amount);
        atomicAdd(&s_cnt[g], 1ULL);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.ngroups; i += blockDim.x) {
        if (s_cnt[i]) {
            const unsigned long long x = s_sum[i];
            const unsigned long long old = atomicAdd(&p.gsum[i], x);
            const long long hi = (long long)(old + x < old ? 1 : 0) - ((long long)x < 0 ? 1 : 0);
            if (hi) atomicAdd(&p.gsum_hi[i], (unsigned long long)hi);
            atomicAdd(&p.gcnt[i], s_cnt[i]);
        }
    }
    n_join = (unsigned long long)warp_sum((i64)n_join);
    n_multi = (unsigned long long)warp_sum((i64)n_multi);
    if ((threadIdx.x & 31) == 0) {
        if (n_join) atomicAdd(&p.counters[1], n_join);
        if (n_multi) atomicAdd(&p.counters[2], n_multi);
    }
}

}  // namespace pg
