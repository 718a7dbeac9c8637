// sx_comm.cu -- the one collective of the multi-GPU path: a gather of call records to one rank over NCCL
// (NVLink 5 / NVSwitch).  Regions shard across ranks with no data-path communication; this is the in-memory
// analogue of the reference's file-level concatIndexVcf (/root/reference/src/python/lib/strelkaSharedWorkflow.py:126-136).
//
// NCCL is resolved at run time with dlopen so that the library loads on hosts without it and shares the libnccl the host
// process (e.g. torch.distributed) has already mapped; nothing of NCCL is needed at compile time -- the handful of types the
// seven entry points use are declared here (they are part of NCCL's stable C ABI: ncclUniqueId is 128 opaque bytes,
// ncclComm_t an opaque pointer, ncclSuccess 0, ncclChar 0).
//
// The gather runs on a stream of its own (s_comm), ordered after the compute stream by an event, so the next step's kernels
// never wait for rank 0's receives; sx_comm_wait() is the only host-blocking call.
#include "sx_internal.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace
{
typedef struct sx_ncclComm* ncclComm_t;
typedef struct
{
    char internal[SX_NCCL_ID_BYTES];
} ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
constexpr ncclResult_t ncclSuccess = 0;
constexpr ncclDataType_t ncclChar = 0;

struct nccl_api
{
    void* lib = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// one resolution per process (std::call_once); a failure keeps its message for every later caller
nccl_api* load_nccl(std::string* err)
{
    static nccl_api api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* n : {"libnccl.so.2", "libnccl.so"})
            if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h)
        {
            const char* d = dlerror();
            api.err = std::string("dlopen(libnccl.so.2) failed: ") + (d ? d : "?");
            return;
        }
        bool ok = true;
        auto sym = [&](auto& field, const char* name) {
            field = reinterpret_cast<std::remove_reference_t<decltype(field)>>(dlsym(h, name));
            if (!field && ok)
            {
                api.err = std::string("NCCL symbol missing: ") + name;
                ok = false;
            }
        };
        sym(api.GetUniqueId, "ncclGetUniqueId");
        sym(api.CommInitRank, "ncclCommInitRank");
        sym(api.CommDestroy, "ncclCommDestroy");
        sym(api.Send, "ncclSend");
        sym(api.Recv, "ncclRecv");
        sym(api.GroupStart, "ncclGroupStart");
        sym(api.GroupEnd, "ncclGroupEnd");
        sym(api.GetErrorString, "ncclGetErrorString");
        if (ok) api.lib = h;
    });
    if (!api.lib && err) *err = api.err;
    return api.lib ? &api : nullptr;
}

int order_after_compute(sx_ctx* ctx)
{
    SX_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->s_compute));
    SX_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comm, ctx->ev_comm, 0));
    return SX_OK;
}

// one grouped exchange on s_comm: every rank sends `bytes` to root, root receives counts[p] bytes from rank p at offs[p]
int grouped_gather(sx_ctx* ctx, const void* local_dev, size_t bytes, void* all_dev, const unsigned long long* counts, const unsigned long long* offs, int root, const char* who)
{
    nccl_api* api = static_cast<nccl_api*>(ctx->nccl_lib);
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl);
    ncclResult_t r = api->GroupStart();
    if (r == ncclSuccess && bytes) r = api->Send(local_dev, bytes, ncclChar, root, comm, ctx->s_comm);
    if (ctx->rank == root)
        for (int p = 0; p < ctx->world && r == ncclSuccess; ++p)
            if (counts[p]) r = api->Recv(static_cast<char*>(all_dev) + offs[p], counts[p], ncclChar, p, comm, ctx->s_comm);
    if (r == ncclSuccess) r = api->GroupEnd();
    if (r != ncclSuccess) return sx_fail(ctx, SX_ERR_NCCL, "%s: %s", who, api->GetErrorString(r));
    return SX_OK;
}
} // namespace

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
    if (ctx->nccl) return sx_fail(ctx, SX_ERR_ARG, "sx_comm_init: this context already has a communicator");
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
    SX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_comm, cudaStreamNonBlocking));
    SX_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    SX_CUDA(ctx, cudaMalloc(&ctx->d_comm_counts, sizeof(unsigned long long) * (size_t)(world_size + 2)));
    ctx->comm_counts.assign((size_t)world_size + 1, 0ull);
    return SX_OK;
}

void sx_comm_release(sx_ctx* ctx)
{
    if (ctx->nccl && ctx->nccl_lib) static_cast<nccl_api*>(ctx->nccl_lib)->CommDestroy(static_cast<ncclComm_t>(ctx->nccl));
    ctx->nccl = nullptr;
    if (ctx->d_comm_counts) cudaFree(ctx->d_comm_counts);
    ctx->d_comm_counts = nullptr;
    if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
    ctx->ev_comm = nullptr;
    if (ctx->s_comm) cudaStreamDestroy(ctx->s_comm);
    ctx->s_comm = nullptr;
}

static int gather_args(sx_ctx* ctx, const void* local_dev, size_t bytes, void* all_dev, int root, const char* who)
{
    if (!ctx) return SX_ERR_ARG;
    if (root < 0 || root >= ctx->world) return sx_fail(ctx, SX_ERR_ARG, "%s: root %d outside [0, %d)", who, root, ctx->world);
    if (bytes && !local_dev) return sx_fail(ctx, SX_ERR_ARG, "%s: local_dev is NULL", who);
    if (ctx->rank == root && !all_dev && ctx->world > 1) return sx_fail(ctx, SX_ERR_ARG, "%s: all_dev is NULL on the root rank", who);
    if (ctx->world > 1 && !ctx->nccl) return sx_fail(ctx, SX_ERR_NCCL, "%s: sx_comm_init has not been called", who);
    return SX_OK;
}

// fixed-size gather, asynchronous: returns once the exchange is enqueued behind the compute stream's work
extern "C" int sx_gather_records_async(sx_ctx* ctx, const void* local_dev, size_t bytes, void* all_dev, int root)
{
    int rc = gather_args(ctx, local_dev, bytes, all_dev, root, "sx_gather_records");
    if (rc) return rc;
    if (ctx->world == 1)
    {
        if (all_dev && all_dev != local_dev && bytes) SX_CUDA(ctx, cudaMemcpyAsync(all_dev, local_dev, bytes, cudaMemcpyDeviceToDevice, ctx->s_compute));
        return SX_OK;
    }
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = order_after_compute(ctx))) return rc;
    for (int p = 0; p < ctx->world; ++p) ctx->comm_counts[(size_t)p] = bytes;
    std::vector<unsigned long long> offs((size_t)ctx->world);
    for (int p = 0; p < ctx->world; ++p) offs[(size_t)p] = (unsigned long long)p * bytes;
    return grouped_gather(ctx, local_dev, bytes, all_dev, ctx->comm_counts.data(), offs.data(), root, "sx_gather_records");
}

extern "C" int sx_comm_wait(sx_ctx* ctx)
{
    if (!ctx) return SX_ERR_ARG;
    SX_CUDA(ctx, cudaStreamSynchronize(ctx->world > 1 && ctx->s_comm ? ctx->s_comm : ctx->s_compute));
    return SX_OK;
}

extern "C" int sx_gather_records(sx_ctx* ctx, const void* local_dev, size_t bytes, void* all_dev, int root)
{
    const int rc = sx_gather_records_async(ctx, local_dev, bytes, all_dev, root);
    return rc ? rc : sx_comm_wait(ctx);
}

// variable-size gather (region shards of unequal size, strelka_b200/shard.py): the per-rank byte counts travel first, rank p's
// block lands at the exclusive prefix sum of the counts.  Blocking (the counts have to reach the host to place the receives).
extern "C" int sx_gatherv_records(sx_ctx* ctx, const void* local_dev, size_t bytes, void* all_dev, size_t all_capacity, uint64_t* offsets_out, int root)
{
    int rc = gather_args(ctx, local_dev, bytes, all_dev, root, "sx_gatherv_records");
    if (rc) return rc;
    if (ctx->world == 1)
    {
        if (bytes > all_capacity) return sx_fail(ctx, SX_ERR_CAPACITY, "sx_gatherv_records: %zu bytes do not fit all_capacity %zu", bytes, all_capacity);
        if (offsets_out) offsets_out[0] = 0, offsets_out[1] = bytes;
        if (all_dev != local_dev && bytes) SX_CUDA(ctx, cudaMemcpyAsync(all_dev, local_dev, bytes, cudaMemcpyDeviceToDevice, ctx->s_compute));
        SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
        return SX_OK;
    }
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = order_after_compute(ctx))) return rc;
    const int W = ctx->world;
    unsigned long long* d = static_cast<unsigned long long*>(ctx->d_comm_counts);
    const unsigned long long mine = bytes;
    // slot W holds this rank's own count (the send buffer); slots 0..W-1 receive on the root
    SX_CUDA(ctx, cudaMemcpyAsync(d + W, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->s_comm));
    std::vector<unsigned long long> eight((size_t)W, sizeof(unsigned long long)), slot((size_t)W);
    for (int p = 0; p < W; ++p) slot[(size_t)p] = (unsigned long long)p * sizeof(unsigned long long);
    if ((rc = grouped_gather(ctx, d + W, sizeof(unsigned long long), d, eight.data(), slot.data(), root, "sx_gatherv_records (counts)"))) return rc;
    std::vector<unsigned long long> offs((size_t)W + 1, 0ull);
    unsigned long long verdict = 1; // 1 = go; 0 = the root's buffer is too small (every rank must learn it, or the senders would hang)
    if (ctx->rank == root)
    {
        SX_CUDA(ctx, cudaMemcpyAsync(ctx->comm_counts.data(), d, sizeof(unsigned long long) * (size_t)W, cudaMemcpyDeviceToHost, ctx->s_comm));
        SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_comm));
        for (int p = 0; p < W; ++p) offs[(size_t)p + 1] = offs[(size_t)p] + ctx->comm_counts[(size_t)p];
        if (offsets_out)
            for (int p = 0; p <= W; ++p) offsets_out[p] = offs[(size_t)p];
        verdict = offs[(size_t)W] <= all_capacity ? 1ull : 0ull;
        SX_CUDA(ctx, cudaMemcpyAsync(d + W + 1, &verdict, sizeof(verdict), cudaMemcpyHostToDevice, ctx->s_comm));
    }
    {   // the root's verdict to every rank (slot W + 1 -> everyone's slot W)
        nccl_api* api = static_cast<nccl_api*>(ctx->nccl_lib);
        ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl);
        ncclResult_t r = api->GroupStart();
        if (ctx->rank == root)
            for (int p = 0; p < W && r == ncclSuccess; ++p) r = api->Send(d + W + 1, sizeof(verdict), ncclChar, p, comm, ctx->s_comm);
        if (r == ncclSuccess) r = api->Recv(d + W, sizeof(verdict), ncclChar, root, comm, ctx->s_comm);
        if (r == ncclSuccess) r = api->GroupEnd();
        if (r != ncclSuccess) return sx_fail(ctx, SX_ERR_NCCL, "sx_gatherv_records (verdict): %s", api->GetErrorString(r));
        SX_CUDA(ctx, cudaMemcpyAsync(&verdict, d + W, sizeof(verdict), cudaMemcpyDeviceToHost, ctx->s_comm));
        SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_comm));
    }
    if (!verdict)
        return sx_fail(ctx, SX_ERR_CAPACITY, "sx_gatherv_records: the ranks' blocks do not fit the root's all_capacity (%zu here; size it from the shard sizes: shard.gathered_offsets)",
                       all_capacity);
    if ((rc = grouped_gather(ctx, local_dev, bytes, all_dev, ctx->comm_counts.data(), offs.data(), root, "sx_gatherv_records"))) return rc;
    SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_comm));
    return SX_OK;
}
