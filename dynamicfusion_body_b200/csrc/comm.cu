// comm.cu -- SURVEY 8b item 9 / 8e: the three collectives of the sharded hot path over raw NCCL, so that a host
// without torch.distributed can drive N GPUs through include/dfb.h:
//   * per frame: root -> all broadcast of the sensor data + node transforms + global rigid dq (dfb_comm_broadcast_frame),
//   * per Gauss-Newton iteration: sum of the flat [H | g | cost] buffer over ranks (dfb_comm_allreduce_f64),
//   * surface extraction: neighbour exchange of halo planes (dfb_comm_sendrecv).
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch already loaded),
// so libdfb_b200.so itself has no link-time dependency on it and loads on a box without NCCL (the calls then fail).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"

namespace {

// the handful of NCCL declarations used here (stable since NCCL 2.0; nccl.h:40-60,230-400)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };

struct Api {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

Api* api() {
    static Api a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("DFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (a.handle) {
#define DFB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
            DFB_SYM(GetVersion, "ncclGetVersion");
            DFB_SYM(GetUniqueId, "ncclGetUniqueId");
            DFB_SYM(CommInitRank, "ncclCommInitRank");
            DFB_SYM(CommDestroy, "ncclCommDestroy");
            DFB_SYM(GetErrorString, "ncclGetErrorString");
            DFB_SYM(Broadcast, "ncclBroadcast");
            DFB_SYM(AllReduce, "ncclAllReduce");
            DFB_SYM(Send, "ncclSend");
            DFB_SYM(Recv, "ncclRecv");
            DFB_SYM(GroupStart, "ncclGroupStart");
            DFB_SYM(GroupEnd, "ncclGroupEnd");
#undef DFB_SYM
            if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast || !a.AllReduce || !a.Send || !a.Recv ||
                !a.GroupStart || !a.GroupEnd) {
                dlclose(a.handle);
                a.handle = nullptr;
            }
        }
    }
    return a.handle ? &a : nullptr;
}

int nccl_fail(ncclResult_t r, const char* what) {
    Api* a = api();
    dfb::set_error("NCCL error in %s: %s", what, (a && a->GetErrorString) ? a->GetErrorString(r) : "?");
    return DFB_ERR_CUDA;
}

#define DFB_NCCL(call, what)                      \
    do {                                          \
        ncclResult_t _r = (call);                 \
        if (_r != 0) return nccl_fail(_r, what);  \
    } while (0)

}  // namespace

struct dfb_comm {
    ncclComm_t comm;
    int world, rank, device;
};

static_assert(DFB_COMM_ID_BYTES == sizeof(ncclUniqueId), "unique id size");

extern "C" int dfb_comm_available(void) {
    Api* a = api();
    if (!a) return 0;
    int v = 0;
    if (a->GetVersion) a->GetVersion(&v);
    return v > 0 ? v : 1;
}

extern "C" int dfb_comm_unique_id(void* id_out) {
    DFB_REQUIRE(id_out, "null pointer");
    Api* a = api();
    DFB_REQUIRE(a, "NCCL (libnccl.so.2) could not be loaded");
    ncclUniqueId id;
    DFB_NCCL(a->GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof(id));
    return DFB_OK;
}

extern "C" int dfb_comm_init(dfb_comm** out, const void* id, int world, int rank, int device) {
    DFB_REQUIRE(out && id, "null pointer");
    DFB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
    Api* a = api();
    DFB_REQUIRE(a, "NCCL (libnccl.so.2) could not be loaded");
    DFB_CUDA(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    DFB_NCCL(a->CommInitRank(&c, world, uid, rank), "ncclCommInitRank");
    dfb_comm* h = new dfb_comm;
    h->comm = c; h->world = world; h->rank = rank; h->device = device;
    *out = h;
    return DFB_OK;
}

extern "C" int dfb_comm_destroy(dfb_comm* c) {
    if (!c) return DFB_OK;
    Api* a = api();
    if (a && c->comm) a->CommDestroy(c->comm);
    delete c;
    return DFB_OK;
}

extern "C" int dfb_comm_rank(const dfb_comm* c) { return c ? c->rank : -1; }
extern "C" int dfb_comm_world(const dfb_comm* c) { return c ? c->world : -1; }

extern "C" int dfb_comm_broadcast(dfb_comm* c, void* buf, int64_t bytes, int root, dfb_stream_t stream) {
    DFB_REQUIRE(c && buf && bytes >= 0 && root >= 0 && root < c->world, "bad arguments");
    if (bytes == 0 || c->world == 1) return DFB_OK;
    DFB_NCCL(api()->Broadcast(buf, buf, (size_t)bytes, NCCL_UINT8, root, c->comm, (cudaStream_t)stream), "ncclBroadcast");
    return DFB_OK;
}

extern "C" int dfb_comm_broadcast_frame(dfb_comm* c, float* depths, int64_t n_depth, float* node_dq, int n_nodes, double* lw, int root,
                                        dfb_stream_t stream) {
    DFB_REQUIRE(c && root >= 0 && root < c->world && n_depth >= 0 && n_nodes >= 0, "bad arguments");
    if (c->world == 1) return DFB_OK;
    Api* a = api();
    cudaStream_t s = (cudaStream_t)stream;
    // one group: the three messages leave as one fused launch
    DFB_NCCL(a->GroupStart(), "ncclGroupStart");
    ncclResult_t r = 0;
    if (depths && n_depth) r = a->Broadcast(depths, depths, (size_t)n_depth, NCCL_FLOAT32, root, c->comm, s);
    if (!r && node_dq && n_nodes) r = a->Broadcast(node_dq, node_dq, (size_t)n_nodes * 8, NCCL_FLOAT32, root, c->comm, s);
    if (!r && lw) r = a->Broadcast(lw, lw, 8, NCCL_FLOAT64, root, c->comm, s);
    const ncclResult_t e = a->GroupEnd();
    if (r) return nccl_fail(r, "ncclBroadcast (frame)");
    if (e) return nccl_fail(e, "ncclGroupEnd");
    return DFB_OK;
}

extern "C" int dfb_comm_allreduce_f64(dfb_comm* c, double* buf, int64_t n, int op, dfb_stream_t stream) {
    DFB_REQUIRE(c && buf && n >= 0, "bad arguments");
    DFB_REQUIRE(op == DFB_COMM_SUM || op == DFB_COMM_MAX, "bad reduction");
    if (n == 0 || c->world == 1) return DFB_OK;
    DFB_NCCL(api()->AllReduce(buf, buf, (size_t)n, NCCL_FLOAT64, op == DFB_COMM_SUM ? NCCL_SUM : NCCL_MAX, c->comm, (cudaStream_t)stream),
             "ncclAllReduce");
    return DFB_OK;
}

extern "C" int dfb_comm_sendrecv(dfb_comm* c, const void* send_buf, int64_t send_bytes, int send_peer, void* recv_buf, int64_t recv_bytes,
                                 int recv_peer, dfb_stream_t stream) {
    DFB_REQUIRE(c, "null communicator");
    Api* a = api();
    cudaStream_t s = (cudaStream_t)stream;
    const bool do_send = send_buf && send_bytes > 0 && send_peer >= 0 && send_peer < c->world;
    const bool do_recv = recv_buf && recv_bytes > 0 && recv_peer >= 0 && recv_peer < c->world;
    if (!do_send && !do_recv) return DFB_OK;
    DFB_NCCL(a->GroupStart(), "ncclGroupStart");
    ncclResult_t r = 0;
    if (do_send) r = a->Send(send_buf, (size_t)send_bytes, NCCL_UINT8, send_peer, c->comm, s);
    if (!r && do_recv) r = a->Recv(recv_buf, (size_t)recv_bytes, NCCL_UINT8, recv_peer, c->comm, s);
    const ncclResult_t e = a->GroupEnd();
    if (r) return nccl_fail(r, "ncclSend/ncclRecv");
    if (e) return nccl_fail(e, "ncclGroupEnd");
    return DFB_OK;
}
