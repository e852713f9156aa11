// pipeline.hpp -- plan / result objects shared by the pipeline builders.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "plan_ir.hpp"

namespace pg {

struct ResCol {
    int type = 0, width = 0, scale = 0;
    std::vector<uint8_t> data;
    std::vector<uint8_t> valid;     // one byte per row (1 = not NULL); empty = no NULLs in this column
    std::vector<char> heap;         // PG_T_VARCHAR: the bytes the pg_string rows point into (sized once, never regrown)
    std::vector<std::string> dict;  // PG_T_DICT8 results that may be ORDER BY keys: code -> string (ordering is by the string)
    void push_null(size_t rows_before, size_t elem)
    {
        if (valid.empty()) valid.assign(rows_before, 1);
        valid.push_back(0);
        data.resize(data.size() + elem);
    }
    void mark_valid() { if (!valid.empty()) valid.push_back(1); }
    template <typename T> void push(const T &v)
    {
        size_t n = data.size();
        data.resize(n + sizeof(T));
        memcpy(data.data() + n, &v, sizeof(T));
    }
};

struct Pipeline {
    std::string explain;
    virtual ~Pipeline();
    virtual int run(pg_result *res) = 0;
};

// RAII device / pinned buffers
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n)
    {
        release();
        if (n == 0) n = 16;
        cudaError_t e = dev_alloc(&p, n);
        if (e != cudaSuccess) { p = nullptr; set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e)); return PG_ENOMEM; }
        bytes = n;
        return PG_OK;
    }
    void release() { if (p) dev_free(p); p = nullptr; bytes = 0; }
    ~DevBuf() { release(); }
    template <typename T> T *as() const { return (T *)p; }
};
struct PinBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n)
    {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e != cudaSuccess) { p = nullptr; set_error("cudaMallocHost(%zu) failed: %s", n, cudaGetErrorString(e)); return PG_ENOMEM; }
        bytes = n;
        return PG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    ~PinBuf() { release(); }
    template <typename T> T *as() const { return (T *)p; }
};

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    int init()
    {
        if (a) return PG_OK;
        PG_CUDA(cudaEventCreate(&a));
        PG_CUDA(cudaEventCreate(&b));
        return PG_OK;
    }
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    float ms() const { float m = 0; cudaEventElapsedTime(&m, a, b); return m; }
};

// PG_TRACE=1: per-phase wall-clock breakdown of pg_plan_execute on stderr (syncs the stream at
// every mark, so use it for diagnosis only -- never while benchmarking)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char *who;
    explicit Trace(const char *w) : who(w)
    {
        const char *e = getenv("PG_TRACE");
        on = e && atoi(e) != 0;
        if (on) { cudaStreamSynchronize(ctx().stream); t0 = std::chrono::steady_clock::now(); }
    }
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(ctx().stream);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[pg trace r%d] %s: %-28s %8.3f ms\n", ctx().rank, who, what,
                std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// all-gather `bytes` from every rank into recv[world][bytes] (comm.cu); world==1 copies.
int comm_allgather(const void *d_send, void *d_recv, size_t bytes, cudaStream_t stream);
// variable all-to-all of fixed-size rows (counts / offsets in rows) over grouped ncclSend/ncclRecv
int comm_alltoallv(const void *d_send, const i64 *send_cnt, const i64 *send_off, void *d_recv, const i64 *recv_cnt,
                   const i64 *recv_off, size_t row_bytes, cudaStream_t stream);

}  // namespace pg

struct pg_result {
    std::vector<pg::ResCol> cols;
    std::vector<std::vector<uint8_t>> valid_scratch;   // packed validity bitmaps handed out by pg_result_next
    std::vector<std::vector<const char *>> dict_ptrs;  // pointer tables handed out by pg_result_column_dict
    pg::i64 nrows = 0, cursor = 0;
    pg_stats stats{};
};

struct pg_plan {
    pg::Node root;
    const pg::Node *topk = nullptr;     // root when it is a PG_OP_TOPK, else null
    pg::Node scan_copy;                 // scan with merged filters (high-cardinality Agg <- Scan)
    pg::Node mark_copy;                 // Filter(mark = b) <- MARK join rewritten as a SEMI / ANTI join
    const pg::Node &agg_root() const { return topk ? root.children[0] : root; }
    std::vector<pg_table *> slots;
    std::vector<uint64_t> bound_versions;
    int pipe_world = 1;
    std::unique_ptr<pg::Pipeline> pipe;
};
