// common.cuh -- shared host/device helpers of libplangpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/plangpu.h"
#include "../../include/plangpu_desc.h"

namespace pg {

typedef long long i64;
typedef unsigned long long u64;
typedef __int128 i128;
typedef unsigned __int128 u128;

// Rows are padded to this multiple so that every 16-byte vector load / bulk copy of
// a full tile stays inside the allocation; the pad is zero filled and masked out.
constexpr i64 ROW_PAD = 8192;

void set_error(const char *fmt, ...);
const char *get_error();

#define PG_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            pg::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,   \
                          __LINE__, cudaGetErrorString(_e));                              \
            return PG_ECUDA;                                                              \
        }                                                                                 \
    } while (0)

#define PG_TRY(call)                       \
    do {                                   \
        int _s = (call);                   \
        if (_s != PG_OK) return _s;        \
    } while (0)

#define PG_FAIL(code, ...)                 \
    do {                                   \
        pg::set_error(__VA_ARGS__);        \
        return (code);                     \
    } while (0)

struct Context {
    bool ready = false;
    int device = -1;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    // pinned staging for pageable host sources
    void *stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // communicator (comm.cu)
    int world = 1, rank = 0;
    void *nccl_comm = nullptr;
};
Context &ctx();
int comm_init_all(const std::vector<Context *> &ctxs);      // comm.cu: one NCCL communicator over the contexts of pg_init_devices

inline int type_size(int t)
{
    switch (t) {
    case PG_T_INT32: case PG_T_DATE32: return 4;
    case PG_T_INT64: case PG_T_DECIMAL64: case PG_T_FLOAT64: return 8;
    case PG_T_CHAR1: case PG_T_DICT8: case PG_T_BOOL: return 1;
    case PG_T_HUGEINT: case PG_T_DECIMAL128: case PG_T_VARCHAR: return 16;
    default: return 0;
    }
}

inline bool is_packable(int t) { return t == PG_T_INT32 || t == PG_T_INT64 || t == PG_T_DATE32 || t == PG_T_DECIMAL64; }

inline i64 round_up(i64 x, i64 m) { return (x + m - 1) / m * m; }

// Device memory comes from a size-keyed cache of cudaMalloc blocks (table.cu): column buffers and
// pipeline scratch of a re-ingested table / re-prepared plan are the same sizes as before, and a
// multi-GB cudaMalloc + cudaFree pair costs tens of milliseconds.  Blocks are rounded up to 2 MiB;
// when an allocation fails the cache is emptied and the request retried.
cudaError_t dev_alloc(void **p, size_t bytes);
void dev_free(void *p);
void dev_trim();          // give every cached block back to the driver

// A device column as the kernels see it.  pg_table_seal re-encodes every integer-family column at the
// narrowest physical width its min/max statistics allow (frame of reference):
//     logical value = base + stored,   stored: pw 8 -> int64 (base 0), 4 -> int32 (signed), 2 -> uint16, 1 -> uint8
// Native (unpacked) columns are the same thing with base 0 and pw = the type's width.
struct NCol {
    const void *p;
    int pw;
    int pad_;
    i64 base;
};

struct Column {
    std::string name;
    int type = 0, width = 0, scale = 0;
    std::vector<std::string> dict;
    void *d_data = nullptr;
    int pw = 0;                   // physical bytes per value of d_data (0 until the first allocation: the type's width)
    i64 base = 0;                 // frame of reference of d_data
    int phys_width() const { return pw ? pw : type_size(type); }
    NCol ncol() const { NCol c; c.p = d_data; c.pw = phys_width(); c.pad_ = 0; c.base = base; return c; }
    uint8_t *d_valid = nullptr;   // packed validity, only allocated once a NULL was seen
    bool has_nulls = false;
    // statistics computed at seal
    bool stats_ok = false;
    i64 vmin = 0, vmax = 0;
    i64 adjacent_equal = 0;                           // rows whose value equals the next row's (clustering)
    i64 adjacent_descents = 0;                        // rows whose value is >= the next row's; 0 => strictly increasing => unique
    uint32_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // byte columns: which codes occur
    // The same statistics over ALL ranks' shards of a SHARDED table, agreed with one all-gather when a plan is
    // prepared (agree_table_stats, plan.cu).  Every decision that shapes a collective -- which pipeline runs,
    // whether NULL counts are carried, dense group layouts, exactness proofs -- reads these, so every rank
    // takes the same path whatever its own shard holds.
    bool g_ok = false;
    bool g_has_nulls = false;
    i64 g_vmin = 0, g_vmax = 0;
    uint32_t g_present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool any_nulls() const { return g_ok ? g_has_nulls : has_nulls; }
    i64 gmin() const { return g_ok ? g_vmin : vmin; }
    i64 gmax() const { return g_ok ? g_vmax : vmax; }
    const uint32_t *gpresent() const { return g_ok ? g_present : present; }
    // PG_T_VARCHAR: host-resident payload, row r = h_bytes[h_off[r] .. h_off[r+1]) (h_off holds nrows+1 entries)
    std::vector<int64_t> h_off;
    std::string h_bytes;
    // ... and a device copy (uploaded at seal) for predicates evaluated on the GPU (LIKE, =, <>)
    char *d_bytes = nullptr;
    int64_t *d_off = nullptr;
};

}  // namespace pg

namespace pg { struct AppendStage; }

struct pg_table {
    std::string name;
    std::vector<pg::Column> cols;
    pg::i64 nrows = 0, capacity = 0;
    pg::i64 dev_rows = 0;            // rows already copied to the device (nrows - dev_rows sit in `stage`)
    pg::AppendStage *stage = nullptr;   // pinned host staging of small appends (table.cu)
    bool sealed = false;
    int dist = 0;              // PG_DIST_*
    pg::i64 global_offset = 0;
    // agreed across ranks (see Column::g_ok): the largest shard and the whole table
    pg::i64 g_max_rows = 0, g_total_rows = 0;
    uint64_t g_version = 0;   // version the agreement was made for (0 = never)
    int g_world = 0;
    pg::i64 max_rows() const { return g_version == version && g_max_rows > 0 ? g_max_rows : nrows; }
    pg::i64 total_rows() const { return g_version == version && g_total_rows > 0 ? g_total_rows : nrows; }
    uint64_t version = 0;   // bumped whenever contents change (plan caches key on it)
};
