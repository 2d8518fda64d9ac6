// comm.cu — the two collectives of the path over NCCL, enqueued on the caller's stream through the C ABI.
//
// Reference: none (the reference is one process); SURVEY.md §8b/§8e: games and trees are independent, the only exchanges are a sum
// all-reduce of the fp32 vector [gradient | loss numerator | position count] once per REINFORCE update (src/train_rl.py:55-66 on R
// ranks) and of int64 result counters at report time.  An `iago_comm` either adopts an ncclComm_t the host created
// (iago_comm_from_nccl) or creates one from a unique id that the host distributes (iago_comm_unique_id on rank 0 ->
// iago_comm_create on every rank).  NCCL is bound at run time (dlopen of libnccl.so.2): the library has no link-time dependency on
// it and loads on machines without NCCL; the entry points then return IAGO_E_STATE.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace iago {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi *nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    // the copy the process already holds (torch's bundled libnccl.so.2 when torch is imported) is found by its soname
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("NCCL is not available: %s", dlerror());
        return nullptr;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    if (!api.ok) set_error("libnccl.so.2 lacks an expected symbol");
    return api.ok ? &api : nullptr;
}

}  // namespace iago

using namespace iago;

struct iago_comm {
    iago_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool owned = false;
};

#define IAGO_NCCL(expr)                                                                                   \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess) {                                                                          \
            set_error("%s failed: %s", #expr, api->GetErrorString(_r));                                   \
            return IAGO_E_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)

extern "C" {

int iago_comm_unique_id(char *id, int64_t bytes) {
    IAGO_REQUIRE(id && bytes >= (int64_t)sizeof(ncclUniqueId), "id buffer of at least 128 bytes");
    NcclApi *api = nccl();
    if (!api) return IAGO_E_STATE;
    ncclUniqueId u;
    IAGO_NCCL(api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return IAGO_OK;
}

int iago_comm_create(iago_ctx *ctx, const char *id, int rank, int world, iago_comm **out) {
    IAGO_REQUIRE(ctx && id && out, "NULL argument");
    IAGO_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world");
    NcclApi *api = nccl();
    if (!api) return IAGO_E_STATE;
    DeviceGuard guard(ctx->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    iago_comm *c = new iago_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    c->owned = true;
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return IAGO_E_CUDA;
    }
    *out = c;
    return IAGO_OK;
}

int iago_comm_from_nccl(iago_ctx *ctx, void *nccl_comm, int rank, int world, iago_comm **out) {
    IAGO_REQUIRE(ctx && nccl_comm && out, "NULL argument");
    if (!nccl()) return IAGO_E_STATE;
    iago_comm *c = new iago_comm();
    c->ctx = ctx;
    c->comm = (ncclComm_t)nccl_comm;
    c->rank = rank;
    c->world = world;
    c->owned = false;   // the caller destroys its communicator
    *out = c;
    return IAGO_OK;
}

int iago_comm_destroy(iago_comm *c) {
    if (!c) return IAGO_OK;
    NcclApi *api = nccl();
    if (api && c->owned && c->comm) {
        DeviceGuard guard(c->ctx->device);
        api->CommDestroy(c->comm);
    }
    delete c;
    return IAGO_OK;
}

int iago_comm_allreduce_sum_f32(iago_comm *c, float *buf, int64_t count, void *stream) {
    IAGO_REQUIRE(c && buf && count >= 0, "NULL argument");
    NcclApi *api = nccl();
    if (!api) return IAGO_E_STATE;
    if (c->world == 1 || count == 0) return IAGO_OK;
    DeviceGuard guard(c->ctx->device);
    IAGO_NCCL(api->AllReduce(buf, buf, (size_t)count, ncclFloat, ncclSum, c->comm, (cudaStream_t)stream));
    return IAGO_OK;
}

int iago_comm_allreduce_sum_i64(iago_comm *c, int64_t *buf, int64_t count, void *stream) {
    IAGO_REQUIRE(c && buf && count >= 0, "NULL argument");
    NcclApi *api = nccl();
    if (!api) return IAGO_E_STATE;
    if (c->world == 1 || count == 0) return IAGO_OK;
    DeviceGuard guard(c->ctx->device);
    IAGO_NCCL(api->AllReduce(buf, buf, (size_t)count, ncclInt64, ncclSum, c->comm, (cudaStream_t)stream));
    return IAGO_OK;
}

int iago_comm_rank(iago_comm *c, int *rank, int *world) {
    IAGO_REQUIRE(c, "NULL argument");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return IAGO_OK;
}

}  // extern "C"
