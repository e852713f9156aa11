// table.cu -- library context and device-resident columnar tables.
//
// A pg_table is the HBM copy of what the reference's scanExecutor would hand the
// hot path chunk by chunk (/root/reference/pkg/compute/executor_scan.go:144-223):
// one contiguous array per column in the device-native encoding of plangpu.h,
// capacity padded to ROW_PAD rows (zero filled) so kernels can always issue full
// 16-byte vector loads.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace pg {

// ---------------------------------------------------------------- device memory cache --
namespace {
struct DevCache {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;          // size -> block
    std::unordered_map<void *, size_t> size_of;         // every block handed out or cached
    size_t cached_bytes = 0;
};
// one cache per CUDA device: a block is only ever handed back to the device it was allocated on
// (single-process multi-device mode, pg_init_devices: one host thread per device)
constexpr int DEV_MAXDEV = 64;
DevCache &dev_cache_of(int device) { static DevCache c[DEV_MAXDEV]; return c[device >= 0 && device < DEV_MAXDEV ? device : 0]; }
DevCache &dev_cache() { return dev_cache_of(ctx().device); }
constexpr size_t DEV_GRAIN = (size_t)2 << 20;
constexpr size_t DEV_CACHE_LIMIT = (size_t)96 << 30;    // keep at most this much parked (B200: 180 GB)
}  // namespace

cudaError_t dev_alloc(void **p, size_t bytes)
{
    DevCache &c = dev_cache();
    const size_t n = (std::max<size_t>(bytes, 1) + DEV_GRAIN - 1) / DEV_GRAIN * DEV_GRAIN;
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto it = c.free_blocks.find(n);
        if (it != c.free_blocks.end()) {
            *p = it->second;
            c.free_blocks.erase(it);
            c.cached_bytes -= n;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, n);
    if (e != cudaSuccess) {          // out of memory with blocks parked in the cache: release them and retry
        cudaGetLastError();
        dev_trim();
        e = cudaMalloc(p, n);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> g(c.mu);
        c.size_of[*p] = n;
    }
    return e;
}

void dev_free(void *p)
{
    if (!p) return;
    DevCache *own = &dev_cache();
    {   // freed from a thread bound to another device: the block goes back to the cache of the device that owns it
        std::lock_guard<std::mutex> g(own->mu);
        if (own->size_of.find(p) == own->size_of.end()) own = nullptr;
    }
    for (int d = 0; d < DEV_MAXDEV && !own; d++) {
        DevCache &c = dev_cache_of(d);
        std::lock_guard<std::mutex> g(c.mu);
        if (c.size_of.find(p) != c.size_of.end()) own = &c;
    }
    if (!own) { cudaFree(p); return; }
    DevCache &c = *own;
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.size_of.find(p);
    if (it == c.size_of.end()) { cudaFree(p); return; }
    if (c.cached_bytes + it->second > DEV_CACHE_LIMIT) {
        c.size_of.erase(it);
        cudaFree(p);
        return;
    }
    c.free_blocks.emplace(it->second, p);
    c.cached_bytes += it->second;
}

void dev_trim()
{
    DevCache &c = dev_cache();
    std::lock_guard<std::mutex> g(c.mu);
    for (auto &kv : c.free_blocks) { c.size_of.erase(kv.second); cudaFree(kv.second); }
    c.free_blocks.clear();
    c.cached_bytes = 0;
}


static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

// The context a call works on: the one this THREAD was bound to with pg_use_device (single-process multi-device
// mode: pg_init_devices creates one context per GPU and the host drives each from its own thread, which is what the
// NCCL collectives inside a plan need), else the process-wide context of pg_init (one process per GPU).
static Context g_default_ctx;
static std::vector<std::unique_ptr<Context>> g_dev_ctx;
static thread_local Context *tl_ctx = nullptr;
Context &ctx() { return tl_ctx ? *tl_ctx : g_default_ctx; }

// ---------------------------------------------------------------- statistics --

// min/max of an integer column; one partial per block merged with 64-bit atomics.
template <typename T>
__global__ void stats_minmax_kernel(const T *__restrict__ v, i64 n, i64 *out_min, i64 *out_max, unsigned long long *out_adj)
{
    i64 lo = INT64_MAX, hi = INT64_MIN;
    unsigned long long adj = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        i64 x = (i64)v[i];
        lo = x < lo ? x : lo;
        hi = x > hi ? x : hi;
        if (i + 1 < n) {                                // the neighbour load hits the same line
            i64 y = (i64)v[i + 1];
            if (y == x) adj += 1;
            if (y <= x) adj += 1ULL << 32;              // high half counts "not strictly increasing" steps
        }
    }
    for (int o = 16; o > 0; o >>= 1) adj += __shfl_xor_sync(0xffffffffu, adj, o);
    if ((threadIdx.x & 31) == 0 && adj) {
        if (adj & 0xffffffffULL) atomicAdd(out_adj, adj & 0xffffffffULL);
        if (adj >> 32) atomicAdd(out_adj + 1, adj >> 32);
    }
    for (int o = 16; o > 0; o >>= 1) {
        i64 l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out_min, lo);
        atomicMax(out_max, hi);
    }
}

// which byte codes occur in a CHAR1 / DICT8 column (256-bit presence set)
__global__ void stats_bytes_kernel(const uint8_t *__restrict__ v, i64 n, uint32_t *present)
{
    __shared__ uint32_t s[8];
    if (threadIdx.x < 8) s[threadIdx.x] = 0;
    __syncthreads();
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        uint8_t b = v[i];
#pragma unroll
        for (int w = 0; w < 8; w++)
            if ((b >> 5) == w) loc[w] |= 1u << (b & 31);
    }
#pragma unroll
    for (int w = 0; w < 8; w++)
        if (loc[w]) atomicOr(&s[w], loc[w]);
    __syncthreads();
    if (threadIdx.x < 8 && s[threadIdx.x]) atomicOr(&present[threadIdx.x], s[threadIdx.x]);
}

// any zero bit among the first n bits of a packed validity bitmap?
__global__ void stats_nulls_kernel(const uint8_t *__restrict__ bits, i64 n, int *has_null)
{
    i64 nbytes = (n + 7) / 8;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (i64)gridDim.x * blockDim.x) {
        uint8_t b = bits[i];
        uint8_t mask = 0xff;
        if (i == nbytes - 1 && (n & 7)) mask = (uint8_t)((1u << (n & 7)) - 1);
        if ((b & mask) != mask) *has_null = 1;
    }
}

// scratch that is released on every exit path
struct ScratchBuf {
    void *p = nullptr;
    ~ScratchBuf() { if (p) dev_free(p); }
    cudaError_t alloc(size_t n) { return dev_alloc(&p, n ? n : 16); }
    template <typename T> T *as() const { return (T *)p; }
};

static int compute_stats(pg_table *t)
{
    Context &c = ctx();
    size_t ncol = t->cols.size();
    ScratchBuf b_adj, b_mm, b_present, b_flag;
    PG_CUDA(b_adj.alloc(sizeof(unsigned long long) * 2 * ncol));
    PG_CUDA(b_mm.alloc(sizeof(i64) * 2 * ncol));
    PG_CUDA(b_present.alloc(sizeof(uint32_t) * 8 * ncol));
    PG_CUDA(b_flag.alloc(sizeof(int) * ncol));
    unsigned long long *d_adj = b_adj.as<unsigned long long>();
    i64 *d_mm = b_mm.as<i64>();
    uint32_t *d_present = b_present.as<uint32_t>();
    int *d_flag = b_flag.as<int>();
    PG_CUDA(cudaMemsetAsync(d_adj, 0, sizeof(unsigned long long) * 2 * ncol, c.stream));
    std::vector<i64> h_mm(2 * ncol);
    for (size_t i = 0; i < ncol; i++) { h_mm[2 * i] = INT64_MAX; h_mm[2 * i + 1] = INT64_MIN; }
    PG_CUDA(cudaMemcpyAsync(d_mm, h_mm.data(), sizeof(i64) * 2 * ncol, cudaMemcpyHostToDevice, c.stream));
    PG_CUDA(cudaMemsetAsync(d_present, 0, sizeof(uint32_t) * 8 * ncol, c.stream));
    PG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int) * ncol, c.stream));
    int grid = c.prop.multiProcessorCount * 8;
    for (size_t i = 0; i < ncol && t->nrows > 0; i++) {
        Column &col = t->cols[i];
        switch (col.type) {
        case PG_T_INT32: case PG_T_DATE32:
            stats_minmax_kernel<int32_t><<<grid, 256, 0, c.stream>>>((const int32_t *)col.d_data, t->nrows, d_mm + 2 * i, d_mm + 2 * i + 1, d_adj + 2 * i);
            break;
        case PG_T_INT64: case PG_T_DECIMAL64:
            stats_minmax_kernel<i64><<<grid, 256, 0, c.stream>>>((const i64 *)col.d_data, t->nrows, d_mm + 2 * i, d_mm + 2 * i + 1, d_adj + 2 * i);
            break;
        case PG_T_CHAR1: case PG_T_DICT8:
            stats_bytes_kernel<<<grid, 256, 0, c.stream>>>((const uint8_t *)col.d_data, t->nrows, d_present + 8 * i);
            break;
        default: break;
        }
        if (col.d_valid) stats_nulls_kernel<<<grid, 256, 0, c.stream>>>(col.d_valid, t->nrows, d_flag + i);
    }
    PG_CUDA(cudaGetLastError());
    std::vector<uint32_t> h_present(8 * ncol);
    std::vector<int> h_flag(ncol);
    std::vector<unsigned long long> h_adj(2 * ncol);
    PG_CUDA(cudaMemcpyAsync(h_adj.data(), d_adj, sizeof(unsigned long long) * 2 * ncol, cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaMemcpyAsync(h_mm.data(), d_mm, sizeof(i64) * 2 * ncol, cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaMemcpyAsync(h_present.data(), d_present, sizeof(uint32_t) * 8 * ncol, cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaMemcpyAsync(h_flag.data(), d_flag, sizeof(int) * ncol, cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaStreamSynchronize(c.stream));
    for (size_t i = 0; i < ncol; i++) {
        Column &col = t->cols[i];
        col.vmin = t->nrows > 0 ? h_mm[2 * i] : 0;
        col.vmax = t->nrows > 0 ? h_mm[2 * i + 1] : 0;
        memcpy(col.present, &h_present[8 * i], sizeof col.present);
        col.has_nulls = col.d_valid != nullptr && h_flag[i] != 0;
        col.adjacent_equal = (i64)h_adj[2 * i];
        col.adjacent_descents = (i64)h_adj[2 * i + 1];
        col.stats_ok = true;
    }
    return PG_OK;
}

// ---------------------------------------------------------------- physical encoding --
// Frame-of-reference narrowing at seal: an integer-family column whose value span fits 1, 2 or 4 bytes
// is re-encoded at that width (NCol: logical = base + stored).  TPC-H at SF100: l_quantity, l_discount,
// l_tax -> 1 byte, dates -> 2, l_extendedprice, keys -> 4; Q1 reads 11 bytes per row instead of 34, Q6 8
// instead of 24.  The scan kernels sit at the HBM roofline, so fewer stored bytes is the only speed left.
template <typename S, typename D>
__global__ void pack_kernel(const S *__restrict__ src, D *__restrict__ dst, i64 n, i64 base)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i] = (D)((i64)src[i] - base);
}
// value = base + stored, written at the column's NATIVE width (append of narrow host buffers, column export)
template <typename S, typename D>
__global__ void widen_kernel(const S *__restrict__ src, D *__restrict__ dst, i64 n, i64 base)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i] = (D)((i64)src[i] + base);
}

template <typename S>
static void launch_pack(const S *src, void *dst, int pw, i64 n, i64 base, cudaStream_t st)
{
    const int grid = ctx().prop.multiProcessorCount * 8;
    if (pw == 1) pack_kernel<S, uint8_t><<<grid, 256, 0, st>>>(src, (uint8_t *)dst, n, base);
    else if (pw == 2) pack_kernel<S, uint16_t><<<grid, 256, 0, st>>>(src, (uint16_t *)dst, n, base);
    else pack_kernel<S, int32_t><<<grid, 256, 0, st>>>(src, (int32_t *)dst, n, base);
}
template <typename D>
static void launch_widen(const void *src, int pw, D *dst, i64 n, i64 base, cudaStream_t st)
{
    const int grid = ctx().prop.multiProcessorCount * 8;
    if (pw == 1) widen_kernel<uint8_t, D><<<grid, 256, 0, st>>>((const uint8_t *)src, dst, n, base);
    else if (pw == 2) widen_kernel<uint16_t, D><<<grid, 256, 0, st>>>((const uint16_t *)src, dst, n, base);
    else if (pw == 4) widen_kernel<int32_t, D><<<grid, 256, 0, st>>>((const int32_t *)src, dst, n, base);
    else widen_kernel<i64, D><<<grid, 256, 0, st>>>((const i64 *)src, dst, n, base);
}

// narrowest width for [vmin, vmax]; base 0 whenever the values themselves fit the stored type
static void choose_encoding(i64 vmin, i64 vmax, int native, int *pw, i64 *base)
{
    const u64 span = (u64)vmax - (u64)vmin;
    *pw = native;
    *base = 0;
    if (span <= 0xffu) { *pw = 1; *base = (vmin >= 0 && vmax <= 0xff) ? 0 : vmin; }
    else if (span <= 0xffffu) { *pw = 2; *base = (vmin >= 0 && vmax <= 0xffff) ? 0 : vmin; }
    else if (span <= 0xffffffffu && native == 8) {
        *pw = 4;
        if (vmin >= INT32_MIN && vmax <= INT32_MAX) *base = 0;
        else *base = span <= (u64)INT32_MAX ? vmin : vmin + ((i64)1 << 31);      // stored stays a signed int32
    }
    if (*pw >= native) { *pw = native; *base = 0; }
}

static int pack_columns(pg_table *t)
{
    Context &c = ctx();
    const char *off = getenv("PG_NO_PACK");
    if (off && atoi(off)) return PG_OK;
    if (t->nrows == 0) return PG_OK;
    for (Column &col : t->cols) {
        if (!is_packable(col.type) || col.pw != 0) continue;
        const int native = type_size(col.type);
        int pw;
        i64 base;
        choose_encoding(col.vmin, col.vmax, native, &pw, &base);
        if (pw == native) continue;
        void *nd = nullptr;
        PG_CUDA(dev_alloc(&nd, (size_t)pw * (size_t)t->capacity));
        PG_CUDA(cudaMemsetAsync(nd, 0, (size_t)pw * (size_t)t->capacity, c.stream));
        if (native == 8) launch_pack<i64>((const i64 *)col.d_data, nd, pw, t->nrows, base, c.stream);
        else launch_pack<int32_t>((const int32_t *)col.d_data, nd, pw, t->nrows, base, c.stream);
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaStreamSynchronize(c.stream));
        dev_free(col.d_data);
        col.d_data = nd;
        col.pw = pw;
        col.base = base;
    }
    return PG_OK;
}

static int grow(pg_table *t, i64 need_rows)
{
    Context &c = ctx();
    if (need_rows <= t->capacity) return PG_OK;
    i64 cap = std::max<i64>(round_up(need_rows, ROW_PAD), t->capacity + t->capacity / 2);
    cap = round_up(cap, ROW_PAD);
    for (Column &col : t->cols) {
        if (col.type == PG_T_VARCHAR) continue;      // host-resident
        size_t esz = (size_t)type_size(col.type);
        void *nd = nullptr;
        PG_CUDA(dev_alloc(&nd, esz * (size_t)cap));
        PG_CUDA(cudaMemsetAsync(nd, 0, esz * (size_t)cap, c.stream));
        if (col.d_data && t->dev_rows > 0)
            PG_CUDA(cudaMemcpyAsync(nd, col.d_data, esz * (size_t)t->dev_rows, cudaMemcpyDeviceToDevice, c.stream));
        PG_CUDA(cudaStreamSynchronize(c.stream));
        if (col.d_data) dev_free(col.d_data);
        col.d_data = nd;
        if (col.d_valid) {
            uint8_t *nv = nullptr;
            PG_CUDA(dev_alloc((void **)&nv, (size_t)cap / 8));
            PG_CUDA(cudaMemsetAsync(nv, 0xff, (size_t)cap / 8, c.stream));
            PG_CUDA(cudaMemcpyAsync(nv, col.d_valid, (size_t)(t->capacity / 8), cudaMemcpyDeviceToDevice, c.stream));
            PG_CUDA(cudaStreamSynchronize(c.stream));
            dev_free(col.d_valid);
            col.d_valid = nv;
        }
    }
    t->capacity = cap;
    return PG_OK;
}

// Host -> device copy of one column slice.  Pinned (or registered) sources are DMAed
// directly (*direct = true: the caller's buffer is still being read when this returns);
// pageable ones are pipelined through two pinned staging buffers and are consumed on return.
static int h2d(void *dst, const void *src, size_t bytes, bool *direct)
{
    Context &c = ctx();
    if (bytes == 0) return PG_OK;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
        *direct = true;
        return PG_OK;
    }
    size_t off = 0;
    int k = 0;
    while (off < bytes) {
        size_t n = std::min(c.stage_bytes, bytes - off);
        PG_CUDA(cudaEventSynchronize(c.stage_ev[k]));
        memcpy(c.stage[k], (const char *)src + off, n);
        PG_CUDA(cudaMemcpyAsync((char *)dst + off, c.stage[k], n, cudaMemcpyHostToDevice, c.stream));
        PG_CUDA(cudaEventRecord(c.stage_ev[k], c.stream));
        off += n;
        k ^= 1;
    }
    return PG_OK;
}

// scatter `n` validity bits from a host bitmap into the device bitmap at bit offset `at`
__global__ void valid_scatter_kernel(uint8_t *dst, i64 at, const uint8_t *src, i64 n)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        int bit = (src[i >> 3] >> (i & 7)) & 1;
        if (!bit) {
            i64 j = at + i;
            atomicAnd((unsigned int *)(dst + ((j >> 3) & ~(i64)3)), ~(1u << (((j >> 3) & 3) * 8 + (j & 7))));
        }
    }
}

static int ensure_valid_bitmap(pg_table *t, Column &col)
{
    if (col.d_valid) return PG_OK;
    PG_CUDA(dev_alloc((void **)&col.d_valid, (size_t)t->capacity / 8));
    PG_CUDA(cudaMemsetAsync(col.d_valid, 0xff, (size_t)t->capacity / 8, ctx().stream));
    return PG_OK;
}

// ---------------------------------------------------------------- small-append staging --
// The Go shim appends what scanExecutor hands it: 2048-row chunks from pageable Go heap
// (executor_scan.go:158-223), ~293 k calls for SF100 lineitem.  Such calls are copied into pinned,
// column-wise host staging (the caller's buffers are consumed when the call returns -- the cgo rule)
// and reach the device as one asynchronous copy per column per STAGE_ROWS rows: no per-call
// synchronisation, no per-call device allocation.  Two staging sets alternate so the host keeps
// filling one while the other is in flight.
constexpr i64 STAGE_ROWS = (i64)1 << 18;          // rows per staging set
constexpr i64 STAGE_SMALL = (i64)1 << 16;         // appends of at most this many rows are staged

struct AppendStage {
    struct Set {
        std::vector<void *> col;                  // pinned, STAGE_ROWS x native width (null for VARCHAR)
        std::vector<uint8_t *> valid;             // pinned packed validity, allocated on first use
        std::vector<char> any_valid;
        i64 rows = 0;
        cudaEvent_t ev = nullptr;
    } set[2];
    int cur = 0;
    std::vector<uint8_t *> d_valid_tmp;           // device scratch for the validity scatter, per column
    ~AppendStage()
    {
        for (Set &s : set) {
            for (void *p : s.col) if (p) cudaFreeHost(p);
            for (uint8_t *p : s.valid) if (p) cudaFreeHost(p);
            if (s.ev) cudaEventDestroy(s.ev);
        }
        for (uint8_t *p : d_valid_tmp) if (p) dev_free(p);
    }
};

static int stage_create(pg_table *t)
{
    if (t->stage) return PG_OK;
    std::unique_ptr<AppendStage> st(new AppendStage());
    const size_t ncol = t->cols.size();
    st->d_valid_tmp.assign(ncol, nullptr);
    for (AppendStage::Set &s : st->set) {
        s.col.assign(ncol, nullptr);
        s.valid.assign(ncol, nullptr);
        s.any_valid.assign(ncol, 0);
        PG_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        for (size_t i = 0; i < ncol; i++) {
            if (t->cols[i].type == PG_T_VARCHAR) continue;
            PG_CUDA(cudaMallocHost(&s.col[i], (size_t)STAGE_ROWS * (size_t)type_size(t->cols[i].type)));
        }
    }
    t->stage = st.release();
    return PG_OK;
}

// copy the current staging set to the device (asynchronously) and switch to the other one
static int stage_flush(pg_table *t)
{
    Context &c = ctx();
    AppendStage *st = t->stage;
    if (!st) return PG_OK;
    AppendStage::Set &s = st->set[st->cur];
    if (s.rows == 0) return PG_OK;
    PG_TRY(grow(t, t->dev_rows + s.rows));
    for (size_t i = 0; i < t->cols.size(); i++) {
        Column &col = t->cols[i];
        if (col.type == PG_T_VARCHAR) continue;
        const size_t esz = (size_t)type_size(col.type);
        PG_CUDA(cudaMemcpyAsync((char *)col.d_data + esz * (size_t)t->dev_rows, s.col[i], esz * (size_t)s.rows, cudaMemcpyHostToDevice, c.stream));
        if (s.any_valid[i]) {
            PG_TRY(ensure_valid_bitmap(t, col));
            if (!st->d_valid_tmp[i]) PG_CUDA(dev_alloc((void **)&st->d_valid_tmp[i], (size_t)STAGE_ROWS / 8));
            PG_CUDA(cudaMemcpyAsync(st->d_valid_tmp[i], s.valid[i], (size_t)(s.rows + 7) / 8, cudaMemcpyHostToDevice, c.stream));
            valid_scatter_kernel<<<64, 256, 0, c.stream>>>(col.d_valid, t->dev_rows, st->d_valid_tmp[i], s.rows);
            PG_CUDA(cudaGetLastError());
        }
    }
    PG_CUDA(cudaEventRecord(s.ev, c.stream));
    t->dev_rows += s.rows;
    st->cur ^= 1;
    AppendStage::Set &n = st->set[st->cur];
    PG_CUDA(cudaEventSynchronize(n.ev));          // the set we are about to refill has left the host
    n.rows = 0;
    std::fill(n.any_valid.begin(), n.any_valid.end(), 0);
    return PG_OK;
}

// host-side widening of a narrow source into the native staging array
template <typename D>
static void widen_host(const pg_colbuf &b, int w, D *dst, i64 n)
{
    const i64 base = b.base;
    switch (w) {
    case 1: { const uint8_t *s = (const uint8_t *)b.data; for (i64 i = 0; i < n; i++) dst[i] = (D)(base + s[i]); break; }
    case 2: { const uint16_t *s = (const uint16_t *)b.data; for (i64 i = 0; i < n; i++) dst[i] = (D)(base + s[i]); break; }
    case 4: { const int32_t *s = (const int32_t *)b.data; for (i64 i = 0; i < n; i++) dst[i] = (D)(base + s[i]); break; }
    default: { const i64 *s = (const i64 *)b.data; for (i64 i = 0; i < n; i++) dst[i] = (D)(base + s[i]); break; }
    }
}

static void copy_bits(uint8_t *dst, i64 at, const uint8_t *src, i64 n)
{
    if ((at & 7) == 0) {
        memcpy(dst + (at >> 3), src, (size_t)(n + 7) / 8);
        if (n & 7) dst[(at + n) >> 3] |= (uint8_t)(0xffu << (n & 7));     // slots past the end stay "valid" until they are written
        return;
    }
    for (i64 k = 0; k < n; k++) {
        const i64 j = at + k;
        if ((src[k >> 3] >> (k & 7)) & 1) dst[j >> 3] |= (uint8_t)(1u << (j & 7));
        else dst[j >> 3] &= (uint8_t)~(1u << (j & 7));
    }
}

static int stage_append(pg_table *t, i64 nrows, const pg_colbuf *cols)
{
    PG_TRY(stage_create(t));
    AppendStage *st = t->stage;
    i64 done = 0;
    while (done < nrows) {
        AppendStage::Set &s = st->set[st->cur];
        const i64 n = std::min(nrows - done, STAGE_ROWS - s.rows);
        for (size_t i = 0; i < t->cols.size(); i++) {
            const Column &col = t->cols[i];
            if (col.type == PG_T_VARCHAR) continue;
            const int esz = type_size(col.type);
            const int w = cols[i].width ? cols[i].width : esz;
            char *dst = (char *)s.col[i] + (size_t)esz * (size_t)s.rows;
            if (w == esz && cols[i].base == 0) {
                memcpy(dst, (const char *)cols[i].data + (size_t)w * (size_t)done, (size_t)esz * (size_t)n);
            } else {
                pg_colbuf b = cols[i];
                b.data = (const char *)cols[i].data + (size_t)w * (size_t)done;
                if (esz == 8) widen_host<i64>(b, w, (i64 *)dst, n);
                else if (esz == 4) widen_host<int32_t>(b, w, (int32_t *)dst, n);
                else widen_host<uint8_t>(b, w, (uint8_t *)dst, n);
            }
            if (cols[i].valid) {
                if (!s.valid[i]) {
                    PG_CUDA(cudaMallocHost((void **)&s.valid[i], (size_t)STAGE_ROWS / 8));
                }
                if (!s.any_valid[i]) { memset(s.valid[i], 0xff, (size_t)STAGE_ROWS / 8); s.any_valid[i] = 1; }
                if ((done & 7) == 0) copy_bits(s.valid[i], s.rows, cols[i].valid + (done >> 3), n);
                else {
                    for (i64 k = 0; k < n; k++) {
                        const i64 j = s.rows + k, q = done + k;
                        if ((cols[i].valid[q >> 3] >> (q & 7)) & 1) s.valid[i][j >> 3] |= (uint8_t)(1u << (j & 7));
                        else s.valid[i][j >> 3] &= (uint8_t)~(1u << (j & 7));
                    }
                }
            }
        }
        s.rows += n;
        done += n;
        if (s.rows == STAGE_ROWS) PG_TRY(stage_flush(t));
    }
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_abi_version(void) { return PG_ABI_VERSION; }

int pg_trim(void)
{
    if (ctx().ready) cudaSetDevice(ctx().device);
    dev_trim();
    return PG_OK;
}

const char *pg_last_error(void) { return get_error(); }

int pg_init_devices(int ndev, const int *devices)
{
    if (ndev < 1 || ndev > DEV_MAXDEV || !devices) PG_FAIL(PG_EINVAL, "pg_init_devices: bad arguments");
    if (!g_dev_ctx.empty() || g_default_ctx.ready) PG_FAIL(PG_ESTATE, "pg_init_devices: the library is already initialised");
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++) if (devices[i] == devices[j]) PG_FAIL(PG_EINVAL, "pg_init_devices: device %d listed twice", devices[i]);
    std::vector<std::unique_ptr<Context>> made;
    for (int i = 0; i < ndev; i++) {
        made.emplace_back(new Context());
        tl_ctx = made.back().get();
        const int s = pg_init(devices[i]);
        if (s != PG_OK) { tl_ctx = nullptr; return s; }
    }
    g_dev_ctx = std::move(made);
    tl_ctx = g_dev_ctx[0].get();
    if (ndev > 1) {
        std::vector<Context *> cs;
        for (auto &c : g_dev_ctx) cs.push_back(c.get());
        const int s = comm_init_all(cs);
        if (s != PG_OK) return s;
    }
    PG_CUDA(cudaSetDevice(g_dev_ctx[0]->device));
    return PG_OK;
}

int pg_use_device(int index)
{
    if (index < 0 || index >= (int)g_dev_ctx.size()) PG_FAIL(PG_EINVAL, "pg_use_device: index %d out of range (%zu devices initialised)", index, g_dev_ctx.size());
    tl_ctx = g_dev_ctx[(size_t)index].get();
    PG_CUDA(cudaSetDevice(tl_ctx->device));
    return PG_OK;
}

int pg_num_devices(void) { return g_dev_ctx.empty() ? (g_default_ctx.ready ? 1 : 0) : (int)g_dev_ctx.size(); }

int pg_init(int device)
{
    Context &c = ctx();
    if (c.ready) {
        if (c.device == device) return PG_OK;
        PG_FAIL(PG_ESTATE, "pg_init: already bound to device %d", c.device);
    }
    int n = 0;
    PG_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) PG_FAIL(PG_EINVAL, "pg_init: device %d out of range (%d visible)", device, n);
    PG_CUDA(cudaSetDevice(device));
    PG_CUDA(cudaGetDeviceProperties(&c.prop, device));
    if (c.prop.major != 10)
        PG_FAIL(PG_EUNSUPPORTED, "pg_init: libplangpu is built for sm_100a only; device %d is sm_%d%d",
                device, c.prop.major, c.prop.minor);
    PG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    PG_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.stage_bytes = (size_t)64 << 20;
    for (int k = 0; k < 2; k++) {
        PG_CUDA(cudaMallocHost(&c.stage[k], c.stage_bytes));
        PG_CUDA(cudaEventCreateWithFlags(&c.stage_ev[k], cudaEventDisableTiming));
    }
    c.device = device;
    c.ready = true;
    return PG_OK;
}

static int shutdown_context(Context &c);

int pg_shutdown(void)
{
    if (!g_dev_ctx.empty()) {            // multi-device mode: every context, whichever thread calls
        for (auto &c : g_dev_ctx) {
            tl_ctx = c.get();
            pg_comm_destroy();
            shutdown_context(*c);
        }
        tl_ctx = nullptr;
        g_dev_ctx.clear();
        return PG_OK;
    }
    return shutdown_context(ctx());
}

static int shutdown_context(Context &c)
{
    if (!c.ready) return PG_OK;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (int k = 0; k < 2; k++) {
        if (c.stage[k]) cudaFreeHost(c.stage[k]);
        if (c.stage_ev[k]) cudaEventDestroy(c.stage_ev[k]);
        c.stage[k] = nullptr;
        c.stage_ev[k] = nullptr;
    }
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    c.stream = c.copy_stream = nullptr;
    c.ready = false;
    c.device = -1;
    return PG_OK;
}

int pg_device_info(pg_devinfo *out)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_device_info: call pg_init first");
    if (!out) PG_FAIL(PG_EINVAL, "pg_device_info: null out");
    out->device = c.device;
    out->sm_count = c.prop.multiProcessorCount;
    out->cc_major = c.prop.major;
    out->cc_minor = c.prop.minor;
    out->hbm_bytes = (int64_t)c.prop.totalGlobalMem;
    out->l2_bytes = (int64_t)c.prop.l2CacheSize;
    out->max_smem_per_block = (int32_t)c.prop.sharedMemPerBlockOptin;
    out->world_size = c.world;
    out->rank = c.rank;
    return PG_OK;
}

int pg_table_create(const char *name, int ncol, const pg_coldesc *cols, pg_table **out)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_table_create: call pg_init first");
    if (!out || !cols || ncol <= 0 || ncol > 256) PG_FAIL(PG_EINVAL, "pg_table_create: bad arguments");
    PG_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<pg_table> t(new pg_table());
    t->name = name ? name : "";
    for (int i = 0; i < ncol; i++) {
        Column col;
        col.name = cols[i].name ? cols[i].name : "";
        col.type = cols[i].type;
        col.width = cols[i].width;
        col.scale = cols[i].scale;
        if (type_size(col.type) == 0 || col.type == PG_T_HUGEINT || col.type == PG_T_DECIMAL128)
            PG_FAIL(PG_EINVAL, "pg_table_create: column %d (%s) has unsupported input type %d", i, col.name.c_str(), col.type);
        // a DECIMAL64 column is an unscaled int64: the coefficient has at most 19 digits (govalues), the scale 0..19
        if (col.type == PG_T_DECIMAL64 && (col.scale < 0 || col.scale > 19 || col.width < 0 || col.width > 19 || (col.width > 0 && col.scale > col.width)))
            PG_FAIL(PG_EINVAL, "pg_table_create: column %d (%s) DECIMAL(%d,%d) is outside DECIMAL64 (width <= 19, 0 <= scale <= width)", i,
                    col.name.c_str(), col.width, col.scale);
        if (col.type == PG_T_DICT8) {
            if (cols[i].dict_len < 0 || cols[i].dict_len > 256 || (cols[i].dict_len > 0 && !cols[i].dict))
                PG_FAIL(PG_EINVAL, "pg_table_create: column %d (%s) bad dictionary", i, col.name.c_str());
            for (int k = 0; k < cols[i].dict_len; k++) col.dict.push_back(cols[i].dict[k] ? cols[i].dict[k] : "");
        }
        t->cols.push_back(col);
    }
    *out = t.release();
    return PG_OK;
}

int pg_table_reserve(pg_table *t, int64_t nrows)
{
    if (!t || nrows < 0) PG_FAIL(PG_EINVAL, "pg_table_reserve: bad arguments");
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_reserve: table %s is sealed", t->name.c_str());
    PG_CUDA(cudaSetDevice(ctx().device));
    return grow(t, std::max<i64>(nrows, 1));
}

int pg_table_append_cols(pg_table *t, int64_t nrows, const pg_colbuf *cols)
{
    Context &c = ctx();
    if (!t || nrows < 0 || (nrows > 0 && !cols)) PG_FAIL(PG_EINVAL, "pg_table_append_cols: bad arguments");
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_append_cols: table %s is sealed", t->name.c_str());
    if (nrows == 0) return PG_OK;
    PG_CUDA(cudaSetDevice(c.device));
    for (size_t i = 0; i < t->cols.size(); i++) {
        const Column &col = t->cols[i];
        if (!cols[i].data) PG_FAIL(PG_EINVAL, "pg_table_append_cols: column %zu (%s) is NULL", i, col.name.c_str());
        const int esz = type_size(col.type), w = cols[i].width ? cols[i].width : esz;
        if (w != 1 && w != 2 && w != 4 && w != 8 && w != 16) PG_FAIL(PG_EINVAL, "pg_table_append_cols: column %s: width %d", col.name.c_str(), w);
        if ((w != esz || cols[i].base != 0) && (!is_packable(col.type) || w > esz))
            PG_FAIL(PG_EINVAL, "pg_table_append_cols: column %s: a narrow / based buffer needs an integer-family column at least as wide", col.name.c_str());
        if (col.type == PG_T_VARCHAR) {
            if (cols[i].valid) PG_FAIL(PG_EUNSUPPORTED, "pg_table_append_cols: NULLs in VARCHAR column %s", col.name.c_str());
            const pg_string *sv = (const pg_string *)cols[i].data;
            for (int64_t r = 0; r < nrows; r++)
                if (sv[r].len < 0 || (sv[r].len > 0 && !sv[r].data)) PG_FAIL(PG_EINVAL, "pg_table_append_cols: bad string in column %s", col.name.c_str());
        }
    }
    // host-resident VARCHAR payloads (validated above, so a failure below cannot leave them half appended)
    for (size_t i = 0; i < t->cols.size(); i++) {
        Column &col = t->cols[i];
        if (col.type != PG_T_VARCHAR) continue;
        const pg_string *sv = (const pg_string *)cols[i].data;
        if (col.h_off.empty()) col.h_off.push_back(0);
        for (int64_t r = 0; r < nrows; r++) {
            col.h_bytes.append(sv[r].data ? sv[r].data : "", (size_t)sv[r].len);
            col.h_off.push_back((int64_t)col.h_bytes.size());
        }
    }
    auto rollback_strings = [&]() {
        for (Column &col : t->cols) {
            if (col.type != PG_T_VARCHAR) continue;
            col.h_off.resize((size_t)t->nrows + 1);
            col.h_bytes.resize((size_t)col.h_off.back());
        }
    };
    int rc = PG_OK;
    if (nrows <= STAGE_SMALL) {
        rc = stage_append(t, nrows, cols);
    } else {
        bool direct = false;
        rc = stage_flush(t);
        if (rc == PG_OK) rc = grow(t, t->dev_rows + nrows);
        std::vector<std::unique_ptr<ScratchBuf>> scratch;
        for (size_t i = 0; i < t->cols.size() && rc == PG_OK; i++) {
            Column &col = t->cols[i];
            if (col.type == PG_T_VARCHAR) continue;
            const size_t esz = (size_t)type_size(col.type);
            const int w = cols[i].width ? cols[i].width : (int)esz;
            char *dst = (char *)col.d_data + esz * (size_t)t->dev_rows;
            if ((size_t)w == esz && cols[i].base == 0) {
                rc = h2d(dst, cols[i].data, esz * (size_t)nrows, &direct);
            } else {
                // narrow host buffer: the narrow bytes cross PCIe, the device widens them into the native array
                scratch.emplace_back(new ScratchBuf());
                if (scratch.back()->alloc((size_t)w * (size_t)nrows) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory for the append scratch"); rc = PG_ENOMEM; break; }
                rc = h2d(scratch.back()->p, cols[i].data, (size_t)w * (size_t)nrows, &direct);
                if (rc != PG_OK) break;
                if (esz == 8) launch_widen<i64>(scratch.back()->p, w, (i64 *)dst, nrows, cols[i].base, c.stream);
                else launch_widen<int32_t>(scratch.back()->p, w, (int32_t *)dst, nrows, cols[i].base, c.stream);
                if (cudaGetLastError() != cudaSuccess) { set_error("widen kernel launch failed"); rc = PG_ECUDA; }
            }
            if (rc == PG_OK && cols[i].valid) {
                rc = ensure_valid_bitmap(t, col);
                if (rc != PG_OK) break;
                const size_t vb = (size_t)(nrows + 7) / 8;
                scratch.emplace_back(new ScratchBuf());
                if (scratch.back()->alloc(vb) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory for the validity scratch"); rc = PG_ENOMEM; break; }
                rc = h2d(scratch.back()->p, cols[i].valid, vb, &direct);
                if (rc != PG_OK) break;
                valid_scatter_kernel<<<256, 256, 0, c.stream>>>(col.d_valid, t->dev_rows, scratch.back()->as<uint8_t>(), nrows);
                if (cudaGetLastError() != cudaSuccess) { set_error("validity scatter launch failed"); rc = PG_ECUDA; }
            }
        }
        // pinned caller buffers are read by the DMA engine until the stream drains (cgo rule: C must not keep
        // using Go memory after the call); the scratch blocks go back to the cache only once their kernels ran
        if (direct || !scratch.empty()) {
            if (cudaStreamSynchronize(c.stream) != cudaSuccess && rc == PG_OK) { set_error("append: stream synchronisation failed"); rc = PG_ECUDA; }
        }
        if (rc == PG_OK) t->dev_rows += nrows;
    }
    if (rc != PG_OK) { rollback_strings(); return rc; }
    t->nrows += nrows;
    t->version++;
    return PG_OK;
}

int pg_table_append(pg_table *t, int64_t nrows, const void *const *cols, const uint8_t *const *valid)
{
    if (!t || nrows < 0 || (nrows > 0 && !cols)) PG_FAIL(PG_EINVAL, "pg_table_append: bad arguments");
    std::vector<pg_colbuf> b(t->cols.size());
    for (size_t i = 0; i < b.size(); i++) {
        b[i].data = nrows > 0 ? cols[i] : nullptr;
        b[i].width = 0;
        b[i].reserved = 0;
        b[i].base = 0;
        b[i].valid = valid ? valid[i] : nullptr;
    }
    return pg_table_append_cols(t, nrows, b.data());
}

int pg_table_device_column(pg_table *t, int col, void **dev_ptr)
{
    if (!t || !dev_ptr || col < 0 || col >= (int)t->cols.size()) PG_FAIL(PG_EINVAL, "pg_table_device_column: bad arguments");
    if (t->capacity == 0) PG_FAIL(PG_ESTATE, "pg_table_device_column: reserve rows first");
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_device_column: table is sealed (columns are stored in their packed encoding)");
    if (t->cols[col].type == PG_T_VARCHAR) PG_FAIL(PG_EUNSUPPORTED, "pg_table_device_column: VARCHAR columns are host-resident");
    *dev_ptr = t->cols[col].d_data;
    return PG_OK;
}

int pg_table_set_rows(pg_table *t, int64_t nrows)
{
    if (!t || nrows < 0 || nrows > t->capacity) PG_FAIL(PG_EINVAL, "pg_table_set_rows: %lld rows exceed capacity", (long long)nrows);
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_set_rows: table is sealed");
    if (t->stage && t->stage->set[t->stage->cur].rows) PG_FAIL(PG_ESTATE, "pg_table_set_rows: appended rows are still staged");
    t->nrows = nrows;
    t->dev_rows = nrows;
    t->version++;
    return PG_OK;
}

int pg_table_seal(pg_table *t, int64_t global_row_offset)
{
    if (!t) PG_FAIL(PG_EINVAL, "pg_table_seal: null table");
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_seal: table %s is already sealed", t->name.c_str());
    PG_CUDA(cudaSetDevice(ctx().device));
    PG_TRY(stage_flush(t));
    if (t->stage) {                       // staging is only needed while the table grows
        PG_CUDA(cudaStreamSynchronize(ctx().stream));
        delete t->stage;
        t->stage = nullptr;
    }
    if (t->capacity == 0) PG_TRY(grow(t, 1));
    t->global_offset = global_row_offset;
    for (Column &col : t->cols) {
        if (col.type != PG_T_VARCHAR) continue;
        if (col.h_off.empty()) col.h_off.push_back(0);
        if ((int64_t)col.h_off.size() != t->nrows + 1)
            PG_FAIL(PG_ESTATE, "pg_table_seal: VARCHAR column %s holds %zu strings for %lld rows", col.name.c_str(), col.h_off.size() - 1, (long long)t->nrows);
    }
    PG_TRY(compute_stats(t));
    PG_TRY(pack_columns(t));
    for (Column &col : t->cols) {        // device copy of string columns for GPU-side predicates (LIKE, =, <>)
        if (col.type != PG_T_VARCHAR) continue;
        if (col.d_off) { dev_free(col.d_off); col.d_off = nullptr; }
        if (col.d_bytes) { dev_free(col.d_bytes); col.d_bytes = nullptr; }
        PG_CUDA(dev_alloc((void **)&col.d_off, col.h_off.size() * 8));
        PG_CUDA(dev_alloc((void **)&col.d_bytes, col.h_bytes.size() + 64));
        PG_CUDA(cudaMemcpyAsync(col.d_off, col.h_off.data(), col.h_off.size() * 8, cudaMemcpyHostToDevice, ctx().stream));
        PG_CUDA(cudaMemcpyAsync(col.d_bytes, col.h_bytes.data(), col.h_bytes.size(), cudaMemcpyHostToDevice, ctx().stream));
        PG_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    t->sealed = true;
    t->version++;
    return PG_OK;
}

int pg_table_set_distribution(pg_table *t, int dist)
{
    if (!t || (dist != PG_DIST_SHARDED && dist != PG_DIST_REPLICATED)) PG_FAIL(PG_EINVAL, "pg_table_set_distribution: bad arguments");
    t->dist = dist;
    t->version++;
    return PG_OK;
}

int pg_table_rows(const pg_table *t, int64_t *nrows)
{
    if (!t || !nrows) PG_FAIL(PG_EINVAL, "pg_table_rows: bad arguments");
    *nrows = t->nrows;
    return PG_OK;
}

int pg_table_column_encoding(const pg_table *t, int col, int32_t *stored_width, int64_t *base)
{
    if (!t || col < 0 || col >= (int)t->cols.size()) PG_FAIL(PG_EINVAL, "pg_table_column_encoding: bad arguments");
    if (stored_width) *stored_width = t->cols[(size_t)col].type == PG_T_VARCHAR ? 0 : t->cols[(size_t)col].phys_width();
    if (base) *base = t->cols[(size_t)col].base;
    return PG_OK;
}

// Bulk export of a column range in the NATIVE encoding (decodes packed columns on the device first).
int pg_table_read_column(pg_table *t, int col, int64_t row, int64_t nrows, void *host_out)
{
    if (!t || !host_out || col < 0 || col >= (int)t->cols.size() || row < 0 || nrows < 0 || row + nrows > t->nrows)
        PG_FAIL(PG_EINVAL, "pg_table_read_column: bad arguments");
    Context &c = ctx();
    PG_CUDA(cudaSetDevice(c.device));
    Column &cl = t->cols[(size_t)col];
    if (cl.type == PG_T_VARCHAR) PG_FAIL(PG_EUNSUPPORTED, "pg_table_read_column: VARCHAR columns are host-resident");
    if (nrows == 0) return PG_OK;
    PG_TRY(stage_flush(t));
    const size_t esz = (size_t)type_size(cl.type);
    const int pw = cl.phys_width();
    if ((size_t)pw == esz && cl.base == 0) {
        PG_CUDA(cudaMemcpyAsync(host_out, (const char *)cl.d_data + esz * (size_t)row, esz * (size_t)nrows, cudaMemcpyDeviceToHost, c.stream));
        PG_CUDA(cudaStreamSynchronize(c.stream));
        return PG_OK;
    }
    ScratchBuf tmp;
    PG_CUDA(tmp.alloc(esz * (size_t)nrows));
    const void *src = (const char *)cl.d_data + (size_t)pw * (size_t)row;
    if (esz == 8) launch_widen<i64>(src, pw, tmp.as<i64>(), nrows, cl.base, c.stream);
    else launch_widen<int32_t>(src, pw, tmp.as<int32_t>(), nrows, cl.base, c.stream);
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaMemcpyAsync(host_out, tmp.p, esz * (size_t)nrows, cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaStreamSynchronize(c.stream));
    return PG_OK;
}

int pg_table_read_column_stored(pg_table *t, int col, int64_t row, int64_t nrows, void *host_out)
{
    if (!t || !host_out || col < 0 || col >= (int)t->cols.size() || row < 0 || nrows < 0 || row + nrows > t->nrows)
        PG_FAIL(PG_EINVAL, "pg_table_read_column_stored: bad arguments");
    Context &c = ctx();
    PG_CUDA(cudaSetDevice(c.device));
    Column &cl = t->cols[(size_t)col];
    if (cl.type == PG_T_VARCHAR) PG_FAIL(PG_EUNSUPPORTED, "pg_table_read_column_stored: VARCHAR columns are host-resident");
    if (nrows == 0) return PG_OK;
    PG_TRY(stage_flush(t));
    const size_t pw = (size_t)cl.phys_width();
    PG_CUDA(cudaMemcpyAsync(host_out, (const char *)cl.d_data + pw * (size_t)row, pw * (size_t)nrows, cudaMemcpyDeviceToHost, c.stream));
    PG_CUDA(cudaStreamSynchronize(c.stream));
    return PG_OK;
}

void pg_table_free(pg_table *t)
{
    if (!t) return;
    if (ctx().ready) {
        cudaSetDevice(ctx().device);
        if (t->stage) cudaStreamSynchronize(ctx().stream);
    }
    delete t->stage;
    for (Column &col : t->cols) {
        if (col.d_data) dev_free(col.d_data);
        if (col.d_valid) dev_free(col.d_valid);
        if (col.d_off) dev_free(col.d_off);
        if (col.d_bytes) dev_free(col.d_bytes);
    }
    delete t;
}

}  // extern "C"
