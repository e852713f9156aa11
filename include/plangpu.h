/*
 * plangpu.h -- C ABI of libplangpu: the B200 (sm_100a) execution hot path behind
 * daviszhen/plan's physical-operator API.
 *
 * What this replaces in the reference (/root/reference, Go):
 *   - OperatorExec{Init,Execute,Close}            pkg/compute/executor_operator.go:52-56
 *     for the operators built by buildOperatorExec pkg/compute/executor.go:305-350:
 *     aggExecutor      pkg/compute/executor_aggr.go:12-272
 *     joinExecutor     pkg/compute/executor_join.go:12-274
 *     filterExecutor   pkg/compute/executor_filter.go:12-118
 *     scan-embedded filter pkg/compute/executor_scan.go:225-241
 *   - the data they exchange: chunk.Chunk / chunk.Vector pkg/chunk/chunk.go:16-20,
 *     pkg/chunk/vector.go:15-22, validity pkg/util/bitmap.go:3-77.
 *
 * The Go side (a `gpuPipelineExec` implementing OperatorExec)
 * binds exactly these entry points through cgo.  Nothing here exposes CUDA, NCCL,
 * torch or C++ types: plain pointers, sizes and opaque handles only.
 *
 * Conventions
 *   - every function returns a pg_status (0 = OK); the message of the last
 *     failure on the calling thread is available from pg_last_error();
 *   - arithmetic faults the reference turns into a Go panic
 *     (function_operator_binary.go:134-140) surface as PG_EOVERFLOW;
 *   - PG_EUNSUPPORTED from pg_plan_compile means "build the stock Go executors for
 *     this subtree" -- plan selection at plan-build time; there is NO CPU fallback
 *     inside this library;
 *   - host buffers passed in are copied before the call returns (cgo rule: C must
 *     not retain Go pointers); buffers handed out stay valid until the owning
 *     handle is freed.
 */
#ifndef PLANGPU_H
#define PLANGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 2

typedef enum {
    PG_OK = 0,
    PG_EINVAL = 1,
    PG_ENOMEM = 2,
    PG_ECUDA = 3,
    PG_ENCCL = 4,
    PG_EOVERFLOW = 5,
    PG_EUNSUPPORTED = 6,
    PG_ESTATE = 7
} pg_status;

/* ---- device-native column encodings (what crosses the boundary) ------------
 * The Go shim flattens each needed chunk.Vector (FLAT/CONST/DICT/SEQUENCE via
 * ToUnifiedFormat, pkg/chunk/vector_format.go:64-97) into one of these.        */
typedef enum {
    PG_T_INT32 = 1,      /* LTID_INTEGER                       4 B                  */
    PG_T_INT64 = 2,      /* LTID_BIGINT                        8 B                  */
    PG_T_DATE32 = 3,     /* LTID_DATE  days since 1970-01-01   4 B                  */
    PG_T_DECIMAL64 = 4,  /* LTID_DECIMAL(w,s) unscaled int64 at the column's scale  */
    PG_T_CHAR1 = 5,      /* VARCHAR(1): the byte itself        1 B                  */
    PG_T_DICT8 = 6,      /* low-cardinality VARCHAR: uint8 code into coldesc.dict   */
    PG_T_FLOAT64 = 7,    /* LTID_DOUBLE                        8 B                  */
    PG_T_HUGEINT = 8,    /* LTID_HUGEINT {uint64 lower; int64 upper}   16 B (results) */
    PG_T_DECIMAL128 = 9, /* result DECIMAL: pg_decimal, 16 B                        */
    PG_T_VARCHAR = 10,   /* LTID_VARCHAR of any cardinality: one pg_string per row (the layout of
                            common.String, pkg/common/string.go:10-13).  Kept in HOST memory by the
                            library: such a column can only be CARRIED to the result (a group key
                            that is functionally dependent on a unique join key, e.g. c_name in
                            TPC-H Q18); predicates, join keys and aggregates on it are refused
                            with PG_EUNSUPPORTED.                                            */
    PG_T_BOOL = 11       /* LTID_BOOLEAN, 1 B (results of row-emitting pipelines: MARK columns, projected predicates) */
} pg_type;

typedef struct {               /* field order of common.String {Len int; Data unsafe.Pointer} (string.go:10-13, amd64) */
    int64_t len;
    const char *data;          /* not NUL-terminated                                */
} pg_string;

/* A DECIMAL result value with govalues' own fields (value = (-1)^neg*coef*10^-scale),
 * so the Go side rebuilds it without struct punning (pkg/chunk/vector.go:256-263). */
typedef struct {
    uint64_t coef;
    int32_t scale;
    uint32_t neg;
} pg_decimal;

typedef struct {
    uint64_t lower;
    int64_t upper;
} pg_hugeint;

typedef struct {
    const char *name;          /* column name (diagnostics only)                   */
    int32_t type;              /* pg_type                                          */
    int32_t width;             /* DECIMAL precision                                */
    int32_t scale;             /* DECIMAL scale                                    */
    int32_t dict_len;          /* PG_T_DICT8: number of dictionary entries         */
    const char *const *dict;   /* PG_T_DICT8: entries, code i -> dict[i]           */
} pg_coldesc;

typedef struct pg_table pg_table;
typedef struct pg_plan pg_plan;
typedef struct pg_result pg_result;

/* ---- library / device ------------------------------------------------------ */
int pg_abi_version(void);
/* Device memory freed by pg_table_free / pg_plan_free is parked in a size-keyed cache for the next ingest or
 * plan (multi-GB cudaMalloc/cudaFree pairs cost tens of milliseconds); pg_trim gives all of it back to the driver. */
int pg_trim(void);
/* Bind the calling process to one GPU (one process per GPU). */
int pg_init(int device);
int pg_shutdown(void);
const char *pg_last_error(void);

typedef struct {
    int32_t device;
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int64_t hbm_bytes;
    int64_t l2_bytes;
    int32_t max_smem_per_block;
    int32_t world_size, rank;
} pg_devinfo;
int pg_device_info(pg_devinfo *out);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------------
 * Rank 0 creates the id, the host (Go: any side channel; Python: torch.distributed)
 * broadcasts the 128 bytes, every rank calls pg_comm_init.  Row-range sharded
 * tables then merge partial aggregates with an all-gather inside pg_plan_execute. */
/* Single-process multi-device mode (the reference is ONE process: pkg/compute runs a query on one goroutine).
 * pg_init_devices binds the library to `ndev` GPUs at once and sets up one NCCL communicator over them
 * (ncclCommInitAll's job); device i is rank i.  Every other call works on the device the CALLING THREAD selected with
 * pg_use_device(i) -- the Go shim runs one goroutine per device, locked to its OS thread (runtime.LockOSThread),
 * which ingests that device's row-range shard and calls pg_plan_execute; the collectives inside a plan (partial
 * aggregate merges, the all-to-all exchanges) match up across those threads exactly as across processes.  Use either
 * pg_init (+ pg_comm_init, one process per GPU) or pg_init_devices, not both.  pg_shutdown ends either mode. */
int pg_init_devices(int ndev, const int *devices);
int pg_use_device(int index);
int pg_num_devices(void);
/* The library's own cross-rank merge of partial aggregates, callable WITHOUT a GPU (host arithmetic only; the pipelines
 * call the same function after their all-gather).  `gathered`: nranks records of `rank_bytes` bytes, each
 * [nvals x {uint64 lo, uint64 hi}] two's-complement 128-bit totals followed by [ngroups] int64 first-row ids
 * (0x7f7f7f7f7f7f7f7f = the group has no row on that rank).  Sums are exact in 128 bits, first rows take the minimum. */
int pg_host_merge_partials(const void *gathered, int64_t rank_bytes, int nranks, int nvals, int ngroups,
                           uint64_t *out_totals, int64_t *out_first);
int pg_comm_unique_id(void *out128);
int pg_comm_init(int world_size, int rank, const void *id128);
int pg_comm_destroy(void);

/* ---- tables: device-resident columnar copies of scan output ------------------
 * Replaces draining scanExecutor chunk by chunk (executor_scan.go:144-223).      */
int pg_table_create(const char *name, int ncol, const pg_coldesc *cols, pg_table **out);
/* Reserve HBM for `nrows` rows up front (optional; append grows geometrically). */
int pg_table_reserve(pg_table *t, int64_t nrows);
/* Append `nrows` rows from HOST buffers (one per column, native encoding).
 * valid[i] may be NULL (all valid) or a packed bitmap, bit=1 valid, LSB first
 * (pkg/util/bitmap.go).  `valid` itself may be NULL.  Copies synchronously.     */
int pg_table_append(pg_table *t, int64_t nrows, const void *const *cols, const uint8_t *const *valid);
/* The same with NARROW host buffers: the shim knows each chunk's value range (it walks the vector anyway to
 * flatten govalues Decimals / Dates, chunk/vector_format.go:64-97) and may hand a column over at 1, 2 or 4 bytes
 * per value with a frame of reference:  value = base + data[i], data[i] read as uint8 / uint16 / int32 / int64
 * for width 1 / 2 / 4 / 8.  Only the narrow bytes cross PCIe; the device widens them.  width 0 = the column's
 * native width (base must be 0).  Narrow buffers are accepted for INT32 / INT64 / DATE32 / DECIMAL64 columns.
 * Appends of at most 65536 rows (the reference's 2048-row chunks) are gathered in pinned host staging and
 * copied asynchronously per 262144 rows: no per-call synchronisation or device allocation.                  */
typedef struct {
    const void *data;
    int32_t width;
    int32_t reserved;
    int64_t base;
    const uint8_t *valid;      /* packed validity for this column or NULL */
} pg_colbuf;
int pg_table_append_cols(pg_table *t, int64_t nrows, const pg_colbuf *cols);
/* Device pointer of column `col` (capacity = reserved rows) for producers that
 * already hold the data in HBM (the in-box generator); follow with set_rows.    */
int pg_table_device_column(pg_table *t, int col, void **dev_ptr);
int pg_table_set_rows(pg_table *t, int64_t nrows);
/* Finish ingest: computes column statistics (min/max, byte-code dictionaries) on
 * the device.  `global_row_offset` is this rank's first row in the unsharded table
 * (0 on a single GPU); shards are contiguous row ranges in rank order.            */
int pg_table_seal(pg_table *t, int64_t global_row_offset);
/* How the table relates to the other ranks when a communicator is up (default SHARDED):
 * SHARDED    = this rank holds one contiguous row range of the table (rank order = row order);
 * REPLICATED = every rank holds the whole table (small dimension tables).
 * Joins between sharded tables run shard-local when the column statistics prove the shards
 * are co-partitioned on the join key (disjoint key ranges); otherwise PG_EUNSUPPORTED.     */
#define PG_DIST_SHARDED 0
#define PG_DIST_REPLICATED 1
int pg_table_set_distribution(pg_table *t, int dist);
int pg_table_rows(const pg_table *t, int64_t *nrows);
/* Physical encoding pg_table_seal chose for a column from its min/max statistics (frame of reference:
 * value = base + stored): stored bytes per value (1, 2, 4 or 8; 0 for host-resident VARCHAR) and the base. */
int pg_table_column_encoding(const pg_table *t, int col, int32_t *stored_width, int64_t *base);
/* Bulk export: `nrows` rows of column `col` starting at `row`, in the column's NATIVE encoding, into a host
 * buffer (the inverse of pg_table_append; the device decodes packed columns first).  Not for VARCHAR.      */
int pg_table_read_column(pg_table *t, int col, int64_t row, int64_t nrows, void *host_out);
/* The same in the column's STORED encoding (pg_table_column_encoding: width bytes per value, value = base + stored):
 * a straight device-to-host copy, and exactly the buffers pg_table_append_cols takes back -- the export / re-ingest
 * pair of a (table, column, version) cache on the Go side (executor_scan.go:158-223 re-reads from storage instead). */
int pg_table_read_column_stored(pg_table *t, int col, int64_t row, int64_t nrows, void *host_out);
void pg_table_free(pg_table *t);

/* ---- plans ------------------------------------------------------------------
 * `desc` is the flat little-endian int64 plan descriptor the shim serialises from a
 * PhysicalOperator subtree (builder_physical_operator.go:49-66) -- see
 * plangpu_desc.h.  The C side never sees Go structs.                             */
int pg_plan_compile(const int64_t *desc, size_t nwords, pg_plan **out);
int pg_plan_bind(pg_plan *p, int slot, pg_table *t);
/* Select kernels for the bound tables (lowering, shape matching, scratch allocation).
 * PG_EUNSUPPORTED here is the signal for the shim's Init() to build the stock executors.
 * pg_plan_execute calls it implicitly when needed.                                    */
int pg_plan_prepare(pg_plan *p);
/* Human-readable description of the kernels chosen (EXPLAIN for the GPU part). */
const char *pg_plan_explain(pg_plan *p);
/* Runs the pipeline to completion (blocking), merges across ranks if a
 * communicator is up, and finalises aggregates in exact arithmetic.              */
int pg_plan_execute(pg_plan *p, pg_result **out);
void pg_plan_free(pg_plan *p);

/* ---- results: re-emitted by the shim as ordinary <=2048-row chunks ----------- */
int pg_result_num_columns(const pg_result *r, int *ncol);
int pg_result_column_type(const pg_result *r, int col, int32_t *type, int32_t *width, int32_t *scale);
int pg_result_rows(const pg_result *r, int64_t *nrows);
/* Next batch of at most max_rows rows: cols[i] points at library-owned host memory
 * in the native encoding; valid[i] is NULL when the batch has no NULLs in column i.
 * *nrows == 0 means Done (executor_operator.go:11-18).                            */
int pg_result_next(pg_result *r, int64_t max_rows, int64_t *nrows, const void **cols, const uint8_t **valid);
int pg_result_rewind(pg_result *r);
/* PG_T_DICT8 result columns: the dictionary the codes index (code i -> entries[i], NUL-terminated, owned by
 * the result).  *nentries == 0 for other columns.  The Go shim builds the VARCHAR vector from it
 * (chunk/vector.go:208-217), so a result needs no side channel to the table it came from. */
int pg_result_column_dict(const pg_result *r, int col, int32_t *nentries, const char *const **entries);

typedef struct {
    double exec_ms;            /* wall time of pg_plan_execute                    */
    double kernel_ms;          /* CUDA-event time of all kernels of the pipeline  */
    double main_kernel_ms;     /* CUDA-event time of the dominant (scan) kernel   */
    double comm_ms;            /* CUDA-event time of the all-to-all ROW exchanges of a join (build + probe side;
                                  exchange.cuh); 0 for pipelines without one                            */
    int64_t rows_scanned;      /* rows read from the probe/fact table             */
    int64_t algorithmic_bytes; /* bytes of referenced columns, each read once     */
    int64_t main_kernel_bytes; /* algorithmic bytes of the dominant kernel        */
    int32_t kernel_launches;
    int32_t reserved;
    int64_t aux[8];            /* pipeline specific counters: [0] rows passing the scan / probe filters, [1] joined rows
                                  (row pipelines: output rows), [2..5] rows passing / rows built of the first two build
                                  stages, [6] result groups (rows) before any LIMIT, [7] rows sent to OTHER ranks by a
                                  shuffle / row exchange ([5] = bytes sent, for a star join with a row exchange)        */
} pg_stats;
int pg_result_stats(const pg_result *r, pg_stats *out);
void pg_result_free(pg_result *r);

#ifdef __cplusplus
}
#endif
#endif /* PLANGPU_H */
