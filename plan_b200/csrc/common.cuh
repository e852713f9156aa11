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

inline int type_size(int t)
{
    switch (t) {
    case PG_T_INT32: case PG_T_DATE32: return 4;
    case PG_T_INT64: case PG_T_DECIMAL64: case PG_T_FLOAT64: return 8;
    case PG_T_CHAR1: case PG_T_DICT8: return 1;
    case PG_T_HUGEINT: case PG_T_DECIMAL128: case PG_T_VARCHAR: return 16;
    default: return 0;
    }
}

inline i64 round_up(i64 x, i64 m) { return (x + m - 1) / m * m; }

// Device memory comes from a size-keyed cache of cudaMalloc blocks (table.cu): column buffers and
// pipeline scratch of a re-ingested table / re-prepared plan are the same sizes as before, and a
// multi-GB cudaMalloc + cudaFree pair costs tens of milliseconds.  Blocks are rounded up to 2 MiB;
// when an allocation fails the cache is emptied and the request retried.
cudaError_t dev_alloc(void **p, size_t bytes);
void dev_free(void *p);
void dev_trim();          // give every cached block back to the driver

struct Column {
    std::string name;
    int type = 0, width = 0, scale = 0;
    std::vector<std::string> dict;
    void *d_data = nullptr;
    uint8_t *d_valid = nullptr;   // packed validity, only allocated once a NULL was seen
    bool has_nulls = false;
    // statistics computed at seal
    bool stats_ok = false;
    i64 vmin = 0, vmax = 0;
    i64 adjacent_equal = 0;                           // rows whose value equals the next row's (clustering)
    i64 adjacent_descents = 0;                        // rows whose value is >= the next row's; 0 => strictly increasing => unique
    uint32_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // byte columns: which codes occur
    // PG_T_VARCHAR: host-resident payload, row r = h_bytes[h_off[r] .. h_off[r+1]) (h_off holds nrows+1 entries)
    std::vector<int64_t> h_off;
    std::string h_bytes;
    // ... and a device copy (uploaded at seal) for predicates evaluated on the GPU (LIKE, =, <>)
    char *d_bytes = nullptr;
    int64_t *d_off = nullptr;
};

}  // namespace pg

struct pg_table {
    std::string name;
    std::vector<pg::Column> cols;
    pg::i64 nrows = 0, capacity = 0;
    bool sealed = false;
    int dist = 0;              // PG_DIST_*
    pg::i64 global_offset = 0;
    uint64_t version = 0;   // bumped whenever contents change (plan caches key on it)
};
