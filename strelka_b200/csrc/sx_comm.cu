// sx_comm.cu -- the one collective of the multi-GPU path: a gather of fixed-size call records to one rank over NCCL
// (NVLink 5 / NVSwitch).  Regions shard across ranks with no data-path communication; this is the in-memory
// analogue of the reference's file-level concatIndexVcf (/root/reference/src/python/lib/strelkaSharedWorkflow.py:126-136).
//
// NCCL is resolved at run time with dlopen so that the library loads on hosts without it and can share the libnccl the host
// process (e.g. torch.distributed) has already mapped.
#include "sx_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

namespace
{
struct nccl_api
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

nccl_api* load_nccl(std::string* err)
{
    static nccl_api api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
    {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h)
    {
        if (err) *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return nullptr;
    }
#define SX_SYM(field, name)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));        \
    if (!api.field)                                                           \
    {                                                                         \
        if (err) *err = std::string("NCCL symbol missing: ") + name;          \
        return nullptr;                                                       \
    }
    SX_SYM(GetUniqueId, "ncclGetUniqueId")
    SX_SYM(CommInitRank, "ncclCommInitRank")
    SX_SYM(CommDestroy, "ncclCommDestroy")
    SX_SYM(Send, "ncclSend")
    SX_SYM(Recv, "ncclRecv")
    SX_SYM(GroupStart, "ncclGroupStart")
    SX_SYM(GroupEnd, "ncclGroupEnd")
    SX_SYM(GetErrorString, "ncclGetErrorString")
#undef SX_SYM
    api.lib = h;
    return &api;
}
} // namespace

static_assert(sizeof(ncclUniqueId) == SX_NCCL_ID_BYTES, "SX_NCCL_ID_BYTES must match ncclUniqueId");

extern "C" int sx_comm_get_unique_id(void* id_out)
{
    if (!id_out) return SX_ERR_ARG;
    std::string err;
    nccl_api* api = load_nccl(&err);
    if (!api)
    {
        sx_fail(nullptr, SX_ERR_NCCL, "%s", err.c_str());
        return SX_ERR_NCCL;
    }
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return SX_ERR_NCCL;
    memcpy(id_out, &id, sizeof(id));
    return SX_OK;
}

extern "C" int sx_comm_init(sx_ctx* ctx, const void* id, int rank, int world_size)
{
    if (!ctx || !id || rank < 0 || rank >= world_size) return sx_fail(ctx, SX_ERR_ARG, "sx_comm_init: bad argument");
    std::string err;
    nccl_api* api = load_nccl(&err);
    if (!api) return sx_fail(ctx, SX_ERR_NCCL, "%s", err.c_str());
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    const ncclResult_t r = api->CommInitRank(&comm, world_size, uid, rank);
    if (r != ncclSuccess) return sx_fail(ctx, SX_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString(r));
    ctx->nccl = comm;
    ctx->nccl_lib = api;
    ctx->rank = rank;
    ctx->world = world_size;
    return SX_OK;
}

extern "C" int sx_gather_records(sx_ctx* ctx, const void* local_dev, size_t bytes, void* all_dev, int root)
{
    if (!ctx) return SX_ERR_ARG;
    if (ctx->world == 1)
    {
        if (all_dev && all_dev != local_dev) SX_CUDA(ctx, cudaMemcpyAsync(all_dev, local_dev, bytes, cudaMemcpyDeviceToDevice, ctx->s_compute));
        SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
        return SX_OK;
    }
    if (!ctx->nccl) return sx_fail(ctx, SX_ERR_NCCL, "sx_gather_records: sx_comm_init has not been called");
    nccl_api* api = static_cast<nccl_api*>(ctx->nccl_lib);
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl);
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclResult_t r = api->GroupStart();
    if (r == ncclSuccess) r = api->Send(local_dev, bytes, ncclChar, root, comm, ctx->s_compute);
    if (ctx->rank == root)
        for (int p = 0; p < ctx->world && r == ncclSuccess; ++p) r = api->Recv(static_cast<char*>(all_dev) + (size_t)p * bytes, bytes, ncclChar, p, comm, ctx->s_compute);
    if (r == ncclSuccess) r = api->GroupEnd();
    if (r != ncclSuccess) return sx_fail(ctx, SX_ERR_NCCL, "sx_gather_records: %s", api->GetErrorString(r));
    SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    return SX_OK;
}
