// stage.cuh -- column tiles staged through shared memory by the bulk-copy engine (sm_100a).
//
// The scan kernels over narrow physical encodings (1/2/4-byte frame-of-reference columns, table.cu) are
// issue-bound when every thread fetches its own few bytes: a byte column gives a 4-byte load per 4 rows.
// Here ONE elected producer lane per CTA moves whole column tiles (tile_rows x width bytes, contiguous in
// HBM) into a ring of shared-memory stages with `cp.async.bulk` (the 1-D form of TMA: no tensor map), each
// stage guarded by a full / empty mbarrier pair; the consumer warps read their rows back with conflict-free
// LDS.32/64/128 and never compute a global address.  Bytes in flight per SM = stages x stage bytes,
// independent of the register file.
//
// Row ownership inside a tile: consumer warp w, quad q (0 <= q < QPT), lane l owns the 4 consecutive rows
//     (w * QPT + q) * 128 + 4 * l + {0,1,2,3}
// so a warp's quad of a width-pw column is one contiguous 128*pw-byte span: LDS.32 (pw 1), LDS.64 (pw 2),
// LDS.128 (pw 4) with lanes at stride 4*pw -- no bank conflicts; 8-byte columns take two LDS.128.
#pragma once
#include "common.cuh"

namespace pg {

constexpr int ST_MAXCOL = 8;
constexpr int ST_CONS_WARPS = 8;                          // consumer warps per CTA
constexpr int ST_CONS_THREADS = ST_CONS_WARPS * 32;
constexpr int ST_THREADS = ST_CONS_THREADS + 32;          // + the producer warp (one active lane)
constexpr int ST_MAXSTAGE = 8;
constexpr int ST_HDR = 128;                               // mbarriers: full[8], empty[8]

struct StageDesc {
    const char *src[ST_MAXCOL];      // column base addresses (16-byte aligned, capacity padded to ROW_PAD rows)
    int pw[ST_MAXCOL];               // bytes per value
    int off[ST_MAXCOL];              // byte offset of the column's tile inside a stage (16-byte aligned)
    int ncol;
    int stage_bytes;                 // multiple of 128
    int nstage;
    int tile_rows;                   // ST_CONS_WARPS * 128 * QPT
};

// host: lay the columns out in a stage; returns stage_bytes
inline int stage_layout(StageDesc *d, int tile_rows)
{
    int off = 0;
    for (int c = 0; c < d->ncol; c++) {
        d->off[c] = off;
        off += tile_rows * d->pw[c];
        off = (off + 127) & ~127;
    }
    d->tile_rows = tile_rows;
    d->stage_bytes = off;
    return off;
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// The tiles of one CTA: iteration k handles tile first + k * step, 0 <= k < count.
struct TileSeq { i64 first, step, count; };
__device__ __forceinline__ TileSeq tile_seq(i64 ntiles, int contig)
{
    TileSeq s;
    if (contig) {                    // a contiguous run per CTA: the per-CTA partials are ORDERED partial sums
        const i64 per = (ntiles + gridDim.x - 1) / gridDim.x;
        s.first = (i64)blockIdx.x * per;
        s.step = 1;
        const i64 left = ntiles - s.first;
        s.count = left <= 0 ? 0 : left < per ? left : per;
    } else {
        s.first = blockIdx.x;
        s.step = gridDim.x;
        s.count = s.first >= ntiles ? 0 : (ntiles - s.first + gridDim.x - 1) / gridDim.x;
    }
    return s;
}

struct StageRing {
    unsigned full0, empty0, data0;       // shared-space addresses
    char *data;                          // generic address of stage 0
};

// all threads; ends with a __syncthreads
__device__ __forceinline__ StageRing stage_ring_init(unsigned char *smem, const StageDesc &d)
{
    StageRing r;
    r.full0 = smem_u32(smem);
    r.empty0 = r.full0 + 8 * ST_MAXSTAGE;
    r.data = (char *)smem + ST_HDR;
    r.data0 = r.full0 + ST_HDR;
    if (threadIdx.x == 0) {
        for (int s = 0; s < d.nstage; s++) {
            mbar_init(r.full0 + 8 * s, 1);
            mbar_init(r.empty0 + 8 * s, ST_CONS_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    return r;
}

// the producer warp's whole job (call with every lane of warp ST_CONS_WARPS; lane 0 works)
__device__ __forceinline__ void stage_produce(const StageRing &r, const StageDesc &d, const TileSeq &seq)
{
    if ((threadIdx.x & 31) != 0) return;
    int s = 0;
    unsigned ph = 0;
    unsigned tx = 0;
    for (int c = 0; c < d.ncol; c++) tx += (unsigned)(d.tile_rows * d.pw[c]);
    for (i64 k = 0; k < seq.count; k++) {
        const i64 tile = seq.first + k * seq.step;
        mbar_wait(r.empty0 + 8 * s, ph ^ 1u);               // passes at once on a fresh barrier
        const unsigned bar = r.full0 + 8 * s;
        mbar_expect_tx(bar, tx);
        const unsigned dst = r.data0 + (unsigned)s * (unsigned)d.stage_bytes;
        for (int c = 0; c < d.ncol; c++) {
            const unsigned bytes = (unsigned)(d.tile_rows * d.pw[c]);
            bulk_g2s(dst + (unsigned)d.off[c], d.src[c] + tile * (i64)bytes, bytes, bar);
        }
        if (++s == d.nstage) { s = 0; ph ^= 1u; }
    }
}

// consumer side cursor
struct StageCursor {
    int s;
    unsigned ph;
};
__device__ __forceinline__ const char *stage_acquire(const StageRing &r, const StageDesc &d, const StageCursor &c)
{
    mbar_wait(r.full0 + 8 * c.s, c.ph);
    return r.data + (size_t)c.s * (size_t)d.stage_bytes;
}
// Every lane of the warp has issued its reads of the stage.  The refill is written by the ASYNC proxy, the reads
// went through the generic proxy: without a proxy fence between a lane's LDS and the arrive that frees the stage,
// the bulk copy of the next tile was observed to land before the reads were performed (Q9's filter pass at SF1
// lost a few of 326137 hits in every second run; with the fence: 0 of 60 runs) -- mbarrier release semantics
// alone do not order generic-proxy reads before async-proxy writes.
__device__ __forceinline__ void stage_release(const StageRing &r, const StageDesc &d, StageCursor &c)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(r.empty0 + 8 * c.s);
    if (++c.s == d.nstage) { c.s = 0; c.ph ^= 1u; }
}

// One quad (4 consecutive rows) of a staged column, held raw in registers; get<j>() yields the STORED value of
// row j as a 32-bit word.  The width is a template parameter: 1 / 2 / 4 bytes, 0 = column absent (reads as 0),
// -1 = decided at run time (warp-uniform switch; the fallback for width combinations that are not instantiated).
template <int PW> struct Quad;
template <> struct Quad<0> {
    __device__ __forceinline__ void load(const char *, int, int) {}
    template <int J> __device__ __forceinline__ unsigned get() const { return 0u; }
};
template <> struct Quad<1> {
    unsigned r;
    __device__ __forceinline__ void load(const char *col, int qrow, int) { r = *(const unsigned *)(col + qrow); }
    template <int J> __device__ __forceinline__ unsigned get() const { return __byte_perm(r, 0u, 0x4440u + J); }
};
template <> struct Quad<2> {
    uint2 r;
    __device__ __forceinline__ void load(const char *col, int qrow, int) { r = *(const uint2 *)(col + 2 * qrow); }
    template <int J> __device__ __forceinline__ unsigned get() const
    {
        const unsigned w = J < 2 ? r.x : r.y;
        return (J & 1) ? w >> 16 : w & 0xffffu;
    }
};
template <> struct Quad<4> {
    uint4 r;
    __device__ __forceinline__ void load(const char *col, int qrow, int) { r = *(const uint4 *)(col + 4 * qrow); }
    template <int J> __device__ __forceinline__ unsigned get() const { return J == 0 ? r.x : J == 1 ? r.y : J == 2 ? r.z : r.w; }
};
template <> struct Quad<-1> {
    unsigned v[4];
    __device__ __forceinline__ void load(const char *col, int qrow, int pw)
    {
        const char *p = col + qrow * pw;
        if (pw == 1) {
            const unsigned x = *(const unsigned *)p;
            v[0] = __byte_perm(x, 0u, 0x4440u); v[1] = __byte_perm(x, 0u, 0x4441u); v[2] = __byte_perm(x, 0u, 0x4442u); v[3] = x >> 24;
        } else if (pw == 2) {
            const uint2 x = *(const uint2 *)p;
            v[0] = x.x & 0xffffu; v[1] = x.x >> 16; v[2] = x.y & 0xffffu; v[3] = x.y >> 16;
        } else if (pw == 4) {
            const uint4 x = *(const uint4 *)p;
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
            v[0] = v[1] = v[2] = v[3] = 0;
        }
    }
    template <int J> __device__ __forceinline__ unsigned get() const { return v[J]; }
};
#endif

}  // namespace pg
