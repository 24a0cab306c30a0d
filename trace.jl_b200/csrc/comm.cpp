// comm.cpp — the library's own NCCL communicator (one rank per trace_ctx / GPU).
//
// the path has exactly two exchange steps - ONE sum of the private films at the end of a Whitted render
// (reference: merge_film_tile! into the shared film, src/film.jl:182-193) and, per SPPM iteration, the all-gather of the
// visible points and ONE all-reduce(sum) of the per-pixel (Phi, M) statistics (reference: the atomics of
// src/integrators/sppm.jl:385-401 on shared memory).  They run inside the library on the context's stream, so a caller
// (Julia, C, Python) needs no collective library of its own: rank 0 obtains an id with trace_comm_unique_id, hands it
// to the other ranks by whatever means it has (threads of one process: a shared variable; processes: a file, MPI, a
// socket), and every rank calls trace_comm_init.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded into the process if there is one - e.g.
// the one PyTorch ships - else the system's), so libtrace_cuda.so itself has no link-time dependency on it and single-GPU
// users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <unistd.h>

#include "context.hpp"

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    if (api.handle || !api.error.empty()) return &api;
    const char* override_path = getenv("TRACE_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);       // the copy the process already has, if any
        if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return &api; }
    bool ok = true;
    auto sym = [&](const char* name) -> void* { void* p = dlsym(api.handle, name); if (!p) { ok = false; api.error = std::string("libnccl lacks ") + name; } return p; };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    return &api;
}

int nccl_fail(trace_ctx* c, const char* what, ncclResult_t r) {
    NcclApi* a = nccl_api();
    return c->fail("%s failed: %s", what, a->GetErrorString ? a->GetErrorString(r) : "?");
}
}  // namespace

#define TR_NCCL(ctx, call)                                          \
    do {                                                            \
        ncclResult_t r__ = (call);                                  \
        if (r__ != ncclSuccess) return nccl_fail((ctx), #call, r__); \
    } while (0)

extern "C" int trace_comm_unique_id(void* id_out) {
    if (!id_out) return 1;
    NcclApi* a = nccl_api();
    if (!a->handle) return 2;
    static_assert(sizeof(ncclUniqueId) == TRACE_COMM_ID_BYTES, "TRACE_COMM_ID_BYTES must be NCCL's unique-id size");
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return 3;
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

extern "C" int trace_comm_init(trace_ctx* c, const void* id, int rank, int world) {
    if (!c) return 1;
    if (!id || world < 1 || rank < 0 || rank >= world) return c->fail("trace_comm_init: bad arguments (rank %d, world %d)", rank, world);
    NcclApi* a = nccl_api();
    if (!a->handle) return c->fail("trace_comm_init: %s", a->error.c_str());
    cudaSetDevice(c->device);
    if (c->comm) { a->CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    TR_NCCL(c, a->CommInitRank(&comm, world, uid, rank));
    c->comm = comm;
    c->rank = rank;
    c->world = world;
    int v = 0;
    a->GetVersion(&v);
    c->nccl_version = v;
    return 0;
}

extern "C" int trace_comm_destroy(trace_ctx* c) {
    if (!c) return 1;
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        comm_p2p_close(c);
        NcclApi* a = nccl_api();
        if (a->handle) a->CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
    }
    c->rank = 0;
    c->world = 1;
    return 0;
}

extern "C" int trace_comm_info(const trace_ctx* c, int* rank, int* world, int* nccl_version) {
    if (!c) return 1;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) *nccl_version = c->comm ? c->nccl_version : 0;
    return 0;
}

// ---- collectives on the context's stream (float32 sums / gathers); used by whitted.cu and sppm.cu
int comm_reduce_sum(trace_ctx* c, const float* send, float* recv, size_t count, int root) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    TR_NCCL(c, a->Reduce(send, recv, count, ncclFloat, ncclSum, root, (ncclComm_t)c->comm, c->stream));
    return 0;
}
int comm_reduce_scatter_sum(trace_ctx* c, const float* send, float* recv, size_t recv_count) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    TR_NCCL(c, a->ReduceScatter(send, recv, recv_count, ncclFloat, ncclSum, (ncclComm_t)c->comm, c->stream));
    return 0;
}
int comm_allreduce_sum(trace_ctx* c, float* buf, size_t count) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    TR_NCCL(c, a->AllReduce(buf, buf, count, ncclFloat, ncclSum, (ncclComm_t)c->comm, c->stream));
    return 0;
}
int comm_allreduce_sum_int(trace_ctx* c, int* buf, size_t count) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    TR_NCCL(c, a->AllReduce(buf, buf, count, ncclInt, ncclSum, (ncclComm_t)c->comm, c->stream));
    return 0;
}
int comm_allgather(trace_ctx* c, const float* send, float* recv, size_t send_count) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    TR_NCCL(c, a->AllGather(send, recv, send_count, ncclFloat, (ncclComm_t)c->comm, c->stream));
    return 0;
}
// several collectives of one exchange step fused into one NCCL launch (the five visible-point arrays of an SPPM iteration)
int comm_group_begin(trace_ctx* c) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    TR_NCCL(c, a->GroupStart());
    return 0;
}
int comm_group_end(trace_ctx* c) {
    NcclApi* a = nccl_api();
    TR_NCCL(c, a->GroupEnd());
    return 0;
}
// Sum of the ranks' buffers delivered to `root` as reduce-scatter (every link busy, each rank sums 1/world of the
// buffer) + a gather of the summed chunks onto the root (grouped send / recv): for a 33 MB film on 8 GPUs ~0.1 ms
// where ncclReduce's ring takes ~0.26 ms.  `buf` holds world * chunk floats; in place; on return the root's buf is the sum.
int comm_reduce_sum_via_scatter(trace_ctx* c, float* buf, size_t chunk, int root) {
    NcclApi* a = nccl_api();
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    ncclComm_t comm = (ncclComm_t)c->comm;
    TR_NCCL(c, a->ReduceScatter(buf, buf + (size_t)c->rank * chunk, chunk, ncclFloat, ncclSum, comm, c->stream));
    TR_NCCL(c, a->GroupStart());
    ncclResult_t r = ncclSuccess;
    if (c->rank == root) {
        for (int p = 0; p < c->world && r == ncclSuccess; ++p)
            if (p != root) r = a->Recv(buf + (size_t)p * chunk, chunk, ncclFloat, p, comm, c->stream);
    } else r = a->Send(buf + (size_t)c->rank * chunk, chunk, ncclFloat, root, comm, c->stream);
    const ncclResult_t e = a->GroupEnd();
    if (r != ncclSuccess) return nccl_fail(c, "ncclSend / ncclRecv", r);
    if (e != ncclSuccess) return nccl_fail(c, "ncclGroupEnd", e);
    return 0;
}

// ---- peer mappings of one device allocation per rank (the private films of a Whitted render)
// Every rank publishes {IPC handle, raw pointer, pid, device} of `base` through an all-gather; a peer in ANOTHER process
// is mapped with cudaIpcOpenMemHandle, a peer in THIS process (threads of one process, julia/TraceCUDA.jl MultiContext)
// is used through its raw pointer after cudaDeviceEnablePeerAccess.  Collective; *all_ok = 1 only when every rank
// mapped every peer (the ranks agree on it through a second tiny all-reduce), else the caller stays on NCCL.
namespace {
struct P2PInfo { cudaIpcMemHandle_t handle; unsigned long long ptr; long long pid; int device; int pad; };
}
void comm_p2p_close(trace_ctx* c) {
    for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    c->p2p_opened.clear();
    c->p2p_peer.clear();
    c->p2p_state = 0;
    c->p2p_npix = 0;
}
int comm_p2p_exchange(trace_ctx* c, void* base, std::vector<void*>& peers_out, std::vector<void*>& opened_out, int* all_ok) {
    NcclApi* a = nccl_api();
    *all_ok = 0;
    if (!c->comm) return c->fail("no communicator: call trace_comm_init first");
    ncclComm_t comm = (ncclComm_t)c->comm;
    const int world = c->world;
    P2PInfo mine;
    memset(&mine, 0, sizeof(mine));
    int ok = cudaIpcGetMemHandle(&mine.handle, base) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    mine.ptr = (unsigned long long)(size_t)base; mine.pid = (long long)getpid(); mine.device = c->device;
    char* d_buf = nullptr;
    TR_CUDA(c, cudaMalloc((void**)&d_buf, (size_t)world * sizeof(P2PInfo) + 2 * sizeof(int)));
    std::vector<P2PInfo> all((size_t)world);
    TR_CUDA(c, cudaMemcpyAsync(d_buf + (size_t)c->rank * sizeof(P2PInfo), &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = a->AllGather(d_buf + (size_t)c->rank * sizeof(P2PInfo), d_buf, sizeof(P2PInfo), ncclChar, comm, c->stream);
    if (r != ncclSuccess) { cudaFree(d_buf); return nccl_fail(c, "ncclAllGather (peer handles)", r); }
    TR_CUDA(c, cudaMemcpyAsync(all.data(), d_buf, (size_t)world * sizeof(P2PInfo), cudaMemcpyDeviceToHost, c->stream));
    TR_CUDA(c, cudaStreamSynchronize(c->stream));
    peers_out.assign((size_t)world, nullptr);
    for (int p = 0; p < world && ok; ++p) {
        if (p == c->rank) { peers_out[(size_t)p] = base; continue; }
        if (all[(size_t)p].pid == mine.pid) {               // same process: the pointer is valid here, the device needs peer access
            if (all[(size_t)p].device != c->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, all[(size_t)p].device);
                if (!can) { ok = 0; break; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(all[(size_t)p].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                cudaGetLastError();
            }
            peers_out[(size_t)p] = (void*)(size_t)all[(size_t)p].ptr;
        } else {
            void* mapped = nullptr;
            if (cudaIpcOpenMemHandle(&mapped, all[(size_t)p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
            opened_out.push_back(mapped);
            peers_out[(size_t)p] = mapped;
        }
    }
    // agreement: the number of ranks that failed, summed
    int* d_fail = reinterpret_cast<int*>(d_buf + (size_t)world * sizeof(P2PInfo));
    const int failed = ok ? 0 : 1;
    int total = 1;
    TR_CUDA(c, cudaMemcpyAsync(d_fail, &failed, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    r = a->AllReduce(d_fail, d_fail, 1, ncclInt, ncclSum, comm, c->stream);
    if (r != ncclSuccess) { cudaFree(d_buf); return nccl_fail(c, "ncclAllReduce (peer mapping agreement)", r); }
    TR_CUDA(c, cudaMemcpyAsync(&total, d_fail, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TR_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_buf);
    *all_ok = total == 0 ? 1 : 0;
    return 0;
}
