// table.cu -- library context and device-resident columnar tables.
//
// A pg_table is the HBM copy of what the reference's scanExecutor would hand the
// hot path chunk by chunk (/root/reference/pkg/compute/executor_scan.go:144-223):
// one contiguous array per column in the device-native encoding of plangpu.h,
// capacity padded to ROW_PAD rows (zero filled) so kernels can always issue full
// 16-byte vector loads.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include <map>
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
DevCache &dev_cache() { static DevCache c; return c; }
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
    DevCache &c = dev_cache();
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

Context &ctx()
{
    static Context c;
    return c;
}

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

static int compute_stats(pg_table *t)
{
    Context &c = ctx();
    i64 *d_mm = nullptr;
    uint32_t *d_present = nullptr;
    int *d_flag = nullptr;
    unsigned long long *d_adj = nullptr;
    size_t ncol = t->cols.size();
    PG_CUDA(cudaMalloc(&d_adj, sizeof(unsigned long long) * 2 * ncol));
    PG_CUDA(cudaMemsetAsync(d_adj, 0, sizeof(unsigned long long) * 2 * ncol, c.stream));
    PG_CUDA(cudaMalloc(&d_mm, sizeof(i64) * 2 * ncol));
    PG_CUDA(cudaMalloc(&d_present, sizeof(uint32_t) * 8 * ncol));
    PG_CUDA(cudaMalloc(&d_flag, sizeof(int) * ncol));
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
    cudaFree(d_adj);
    cudaFree(d_mm);
    cudaFree(d_present);
    cudaFree(d_flag);
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
        if (col.d_data && t->nrows > 0)
            PG_CUDA(cudaMemcpyAsync(nd, col.d_data, esz * (size_t)t->nrows, cudaMemcpyDeviceToDevice, c.stream));
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
// directly; pageable ones are pipelined through two pinned staging buffers.
static int h2d(void *dst, const void *src, size_t bytes)
{
    Context &c = ctx();
    if (bytes == 0) return PG_OK;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
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

int pg_shutdown(void)
{
    Context &c = ctx();
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
    pg_table *t = new pg_table();
    t->name = name ? name : "";
    for (int i = 0; i < ncol; i++) {
        Column col;
        col.name = cols[i].name ? cols[i].name : "";
        col.type = cols[i].type;
        col.width = cols[i].width;
        col.scale = cols[i].scale;
        if (type_size(col.type) == 0 || col.type == PG_T_HUGEINT || col.type == PG_T_DECIMAL128) {
            delete t;
            PG_FAIL(PG_EINVAL, "pg_table_create: column %d (%s) has unsupported input type %d", i, col.name.c_str(), col.type);
        }
        if (col.type == PG_T_DICT8) {
            if (cols[i].dict_len < 0 || cols[i].dict_len > 256 || (cols[i].dict_len > 0 && !cols[i].dict)) {
                delete t;
                PG_FAIL(PG_EINVAL, "pg_table_create: column %d (%s) bad dictionary", i, col.name.c_str());
            }
            for (int k = 0; k < cols[i].dict_len; k++) col.dict.push_back(cols[i].dict[k] ? cols[i].dict[k] : "");
        }
        t->cols.push_back(col);
    }
    *out = t;
    return PG_OK;
}

int pg_table_reserve(pg_table *t, int64_t nrows)
{
    if (!t || nrows < 0) PG_FAIL(PG_EINVAL, "pg_table_reserve: bad arguments");
    PG_CUDA(cudaSetDevice(ctx().device));
    return grow(t, std::max<i64>(nrows, 1));
}

int pg_table_append(pg_table *t, int64_t nrows, const void *const *cols, const uint8_t *const *valid)
{
    Context &c = ctx();
    if (!t || nrows < 0 || (nrows > 0 && !cols)) PG_FAIL(PG_EINVAL, "pg_table_append: bad arguments");
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_append: table %s is sealed", t->name.c_str());
    if (nrows == 0) return PG_OK;
    PG_CUDA(cudaSetDevice(c.device));
    PG_TRY(grow(t, t->nrows + nrows));
    for (size_t i = 0; i < t->cols.size(); i++) {
        Column &col = t->cols[i];
        if (!cols[i]) PG_FAIL(PG_EINVAL, "pg_table_append: column %zu (%s) is NULL", i, col.name.c_str());
        if (col.type == PG_T_VARCHAR) {
            if (valid && valid[i]) PG_FAIL(PG_EUNSUPPORTED, "pg_table_append: NULLs in VARCHAR column %s", col.name.c_str());
            const pg_string *sv = (const pg_string *)cols[i];
            if (col.h_off.empty()) col.h_off.push_back(0);
            for (int64_t r = 0; r < nrows; r++) {
                if (sv[r].len < 0 || (sv[r].len > 0 && !sv[r].data)) PG_FAIL(PG_EINVAL, "pg_table_append: bad string in column %s", col.name.c_str());
                col.h_bytes.append(sv[r].data ? sv[r].data : "", (size_t)sv[r].len);
                col.h_off.push_back((int64_t)col.h_bytes.size());
            }
            continue;
        }
        size_t esz = (size_t)type_size(col.type);
        PG_TRY(h2d((char *)col.d_data + esz * (size_t)t->nrows, cols[i], esz * (size_t)nrows));
        if (valid && valid[i]) {
            if (!col.d_valid) {
                PG_CUDA(dev_alloc((void **)&col.d_valid, (size_t)t->capacity / 8));
                PG_CUDA(cudaMemsetAsync(col.d_valid, 0xff, (size_t)t->capacity / 8, c.stream));
            }
            size_t vb = (size_t)(nrows + 7) / 8;
            uint8_t *d_tmp = nullptr;
            PG_CUDA(cudaMalloc(&d_tmp, vb));
            PG_TRY(h2d(d_tmp, valid[i], vb));
            valid_scatter_kernel<<<256, 256, 0, c.stream>>>(col.d_valid, t->nrows, d_tmp, nrows);
            PG_CUDA(cudaGetLastError());
            PG_CUDA(cudaStreamSynchronize(c.stream));
            cudaFree(d_tmp);
        }
    }
    // the host buffers may be reused by the caller as soon as we return (cgo rule)
    PG_CUDA(cudaStreamSynchronize(c.stream));
    t->nrows += nrows;
    t->version++;
    return PG_OK;
}

int pg_table_device_column(pg_table *t, int col, void **dev_ptr)
{
    if (!t || !dev_ptr || col < 0 || col >= (int)t->cols.size()) PG_FAIL(PG_EINVAL, "pg_table_device_column: bad arguments");
    if (t->capacity == 0) PG_FAIL(PG_ESTATE, "pg_table_device_column: reserve rows first");
    if (t->cols[col].type == PG_T_VARCHAR) PG_FAIL(PG_EUNSUPPORTED, "pg_table_device_column: VARCHAR columns are host-resident");
    *dev_ptr = t->cols[col].d_data;
    return PG_OK;
}

int pg_table_set_rows(pg_table *t, int64_t nrows)
{
    if (!t || nrows < 0 || nrows > t->capacity) PG_FAIL(PG_EINVAL, "pg_table_set_rows: %lld rows exceed capacity", (long long)nrows);
    if (t->sealed) PG_FAIL(PG_ESTATE, "pg_table_set_rows: table is sealed");
    t->nrows = nrows;
    t->version++;
    return PG_OK;
}

int pg_table_seal(pg_table *t, int64_t global_row_offset)
{
    if (!t) PG_FAIL(PG_EINVAL, "pg_table_seal: null table");
    PG_CUDA(cudaSetDevice(ctx().device));
    if (t->capacity == 0) PG_TRY(grow(t, 1));
    t->global_offset = global_row_offset;
    for (Column &col : t->cols) {
        if (col.type != PG_T_VARCHAR) continue;
        if (col.h_off.empty()) col.h_off.push_back(0);
        if ((int64_t)col.h_off.size() != t->nrows + 1)
            PG_FAIL(PG_ESTATE, "pg_table_seal: VARCHAR column %s holds %zu strings for %lld rows", col.name.c_str(), col.h_off.size() - 1, (long long)t->nrows);
    }
    PG_TRY(compute_stats(t));
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

void pg_table_free(pg_table *t)
{
    if (!t) return;
    if (ctx().ready) cudaSetDevice(ctx().device);
    for (Column &col : t->cols) {
        if (col.d_data) dev_free(col.d_data);
        if (col.d_valid) dev_free(col.d_valid);
        if (col.d_off) dev_free(col.d_off);
        if (col.d_bytes) dev_free(col.d_bytes);
    }
    delete t;
}

}  // extern "C"
