// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process and single-threaded (no goroutine in pkg/compute,
// SURVEY.md 2); sharding is this build's addition.  lineitem/orders are row-range
// sharded; partial aggregates (a few hundred bytes) are merged with ncclAllGather and an
// identical, rank-ordered exact merge on every rank -- NCCL has no 128-bit sum, and a
// rank-ordered merge keeps the result independent of the reduction tree.
//
// NCCL is resolved at run time (dlopen) so that a single-GPU process needs no NCCL at
// all and a multi-GPU one shares the libnccl its host (torch) already loaded.
#include <dlfcn.h>

#include "common.cuh"
#include "pipeline.hpp"

namespace pg {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt64 = 4, ncclUint64 = 5 };

struct Nccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static Nccl &nccl()
{
    static Nccl n;
    return n;
}

static int load_nccl()
{
    Nccl &n = nccl();
    if (n.handle) return PG_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
    }
    if (!n.handle) PG_FAIL(PG_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define PG_SYM(field, name)                                                   \
    *(void **)(&n.field) = dlsym(n.handle, name);                             \
    if (!n.field) PG_FAIL(PG_ENCCL, "libnccl lacks symbol %s", name)
    PG_SYM(GetUniqueId, "ncclGetUniqueId");
    PG_SYM(CommInitRank, "ncclCommInitRank");
    PG_SYM(CommDestroy, "ncclCommDestroy");
    PG_SYM(AllGather, "ncclAllGather");
    PG_SYM(Send, "ncclSend");
    PG_SYM(Recv, "ncclRecv");
    PG_SYM(GroupStart, "ncclGroupStart");
    PG_SYM(GroupEnd, "ncclGroupEnd");
    PG_SYM(GetErrorString, "ncclGetErrorString");
#undef PG_SYM
    return PG_OK;
}

#define PG_NCCL(call)                                                                         \
    do {                                                                                      \
        ncclResult_t _r = (call);                                                             \
        if (_r != 0) PG_FAIL(PG_ENCCL, "NCCL error %d (%s) at %s:%d", _r,                     \
                             nccl().GetErrorString ? nccl().GetErrorString(_r) : "?", __FILE__, __LINE__); \
    } while (0)

int comm_allgather(const void *d_send, void *d_recv, size_t bytes, cudaStream_t stream)
{
    Context &c = ctx();
    if (c.world <= 1) {
        PG_CUDA(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, stream));
        return PG_OK;
    }
    PG_NCCL(nccl().AllGather(d_send, d_recv, bytes, ncclInt8, (ncclComm_t)c.nccl_comm, stream));
    return PG_OK;
}

// variable all-to-all of fixed-size rows: counts/offsets in rows, grouped ncclSend/ncclRecv (one pair
// per peer; NVSwitch gives every pair full bandwidth, so no ordering tricks are needed)
int comm_alltoallv(const void *d_send, const i64 *send_cnt, const i64 *send_off, void *d_recv, const i64 *recv_cnt,
                   const i64 *recv_off, size_t row_bytes, cudaStream_t stream)
{
    Context &c = ctx();
    if (c.world <= 1) {
        PG_CUDA(cudaMemcpyAsync(d_recv, d_send, (size_t)send_cnt[0] * row_bytes, cudaMemcpyDeviceToDevice, stream));
        return PG_OK;
    }
    PG_NCCL(nccl().GroupStart());
    for (int r = 0; r < c.world; r++) {
        if (send_cnt[r] > 0)
            PG_NCCL(nccl().Send((const char *)d_send + (size_t)send_off[r] * row_bytes, (size_t)send_cnt[r] * row_bytes, ncclInt8, r,
                                (ncclComm_t)c.nccl_comm, stream));
        if (recv_cnt[r] > 0)
            PG_NCCL(nccl().Recv((char *)d_recv + (size_t)recv_off[r] * row_bytes, (size_t)recv_cnt[r] * row_bytes, ncclInt8, r,
                                (ncclComm_t)c.nccl_comm, stream));
    }
    PG_NCCL(nccl().GroupEnd());
    return PG_OK;
}

// ncclCommInitAll's job for contexts this process owns: one rank per context, initialised inside one group from the
// calling thread (the documented single-thread multi-device form); rank i = contexts[i].
int comm_init_all(const std::vector<Context *> &ctxs)
{
    const int n = (int)ctxs.size();
    PG_TRY(load_nccl());
    ncclUniqueId id;
    PG_NCCL(nccl().GetUniqueId(&id));
    std::vector<ncclComm_t> comms((size_t)n, nullptr);
    PG_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; i++) {
        PG_CUDA(cudaSetDevice(ctxs[(size_t)i]->device));
        PG_NCCL(nccl().CommInitRank(&comms[(size_t)i], n, id, i));
    }
    PG_NCCL(nccl().GroupEnd());
    for (int i = 0; i < n; i++) {
        ctxs[(size_t)i]->nccl_comm = comms[(size_t)i];
        ctxs[(size_t)i]->world = n;
        ctxs[(size_t)i]->rank = i;
    }
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_comm_unique_id(void *out128)
{
    if (!out128) PG_FAIL(PG_EINVAL, "pg_comm_unique_id: null");
    PG_TRY(load_nccl());
    ncclUniqueId id;
    PG_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return PG_OK;
}

int pg_comm_init(int world_size, int rank, const void *id128)
{
    Context &c = ctx();
    if (!c.ready) PG_FAIL(PG_ESTATE, "pg_comm_init: call pg_init first");
    if (world_size < 1 || rank < 0 || rank >= world_size) PG_FAIL(PG_EINVAL, "pg_comm_init: bad world/rank");
    if (c.nccl_comm) PG_FAIL(PG_ESTATE, "pg_comm_init: communicator already up");
    if (world_size == 1) { c.world = 1; c.rank = 0; return PG_OK; }
    if (!id128) PG_FAIL(PG_EINVAL, "pg_comm_init: null id");
    PG_TRY(load_nccl());
    PG_CUDA(cudaSetDevice(c.device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    PG_NCCL(nccl().CommInitRank(&comm, world_size, id, rank));
    c.nccl_comm = comm;
    c.world = world_size;
    c.rank = rank;
    return PG_OK;
}

int pg_comm_destroy(void)
{
    Context &c = ctx();
    if (c.nccl_comm) {
        cudaSetDevice(c.device);
        cudaStreamSynchronize(c.stream);
        nccl().CommDestroy((ncclComm_t)c.nccl_comm);
        c.nccl_comm = nullptr;
    }
    c.world = 1;
    c.rank = 0;
    return PG_OK;
}

}  // extern "C"
