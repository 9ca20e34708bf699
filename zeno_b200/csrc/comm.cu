// libflipb200 -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
// NCCL is resolved at run time with dlopen (the PyTorch wheel ships libnccl.so.2), so the
// single-GPU library has no link-time dependency on it. The communicator is created from a
// 128-byte unique id that the host side distributes (torch.distributed / MPI / a file).
#include "world.cuh"
#include <dlfcn.h>
#include <cstring>

namespace fb {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_UINT8 = 1 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi& nccl() {
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) throw Error(FLIPB200_ERR_COMM, "libnccl.so.2 not found (import torch first, or add nvidia/nccl/lib to LD_LIBRARY_PATH)");
    auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) throw Error(FLIPB200_ERR_COMM, std::string("missing NCCL symbol ") + s); return p; };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclAllReduce");
    api.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclRecv");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    return api;
}
#define FB_NCCL(call)                                                                                  \
    do {                                                                                               \
        int r_ = (call);                                                                               \
        if (r_ != 0) throw fb::Error(FLIPB200_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r_)); \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
};

void comm_destroy(World* w) {
    if (w->comm) {
        if (w->comm->comm) nccl().CommDestroy(w->comm->comm);
        delete w->comm;
        w->comm = nullptr;
    }
}
void comm_allreduce_f32(World* w, float* buf, size_t n, bool isMax) {
    if (!w->comm || w->nRanks == 1) return;
    FB_NCCL(nccl().AllReduce(buf, buf, n, NCCL_FLOAT32, isMax ? NCCL_MAX : NCCL_SUM, w->comm->comm, w->stream));
    w->launches++;
}

}  // namespace fb

extern "C" {
int flipb200_comm_unique_id(uint8_t id[128]) {
    try {
        fb::NcclUniqueId u;
        int r = fb::nccl().GetUniqueId(&u);
        if (r != 0) return FLIPB200_ERR_COMM;
        std::memcpy(id, u.internal, 128);
        return FLIPB200_OK;
    } catch (...) { return FLIPB200_ERR_COMM; }
}
int flipb200_comm_init(flipb200_world* w, int rank, int nRanks, const uint8_t id[128]) {
    try {
        if (!w || rank < 0 || rank >= nRanks) return FLIPB200_ERR_ARG;
        cudaSetDevice(w->device);
        fb::comm_destroy(w);
        fb::NcclUniqueId u;
        std::memcpy(u.internal, id, 128);
        w->comm = new fb::Comm();
        int r = fb::nccl().CommInitRank(&w->comm->comm, nRanks, u, rank);
        if (r != 0) { delete w->comm; w->comm = nullptr; return FLIPB200_ERR_COMM; }
        w->rank = rank; w->nRanks = nRanks;
        return FLIPB200_OK;
    } catch (...) { return FLIPB200_ERR_COMM; }
}
}
