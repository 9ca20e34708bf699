// libflipb200 -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
// NCCL is resolved at run time with dlopen (the PyTorch wheel ships libnccl.so.2), so the
// single-GPU library has no link-time dependency on it. The communicator is created from a
// 128-byte unique id that the host side distributes (torch.distributed / MPI / a file).
// A second, in-process backend ("local") connects several worlds of ONE process (one host thread each) through
// mailboxes and device-to-device copies: it carries exactly the same message sequence as the NCCL backend and is
// what lets the slab-decomposition logic run on a single-GPU box (tests/test_dd_gpu.py).
#include "world.cuh"
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <atomic>

namespace fb {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { NCCL_FLOAT32 = 7, NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_UINT8 = 1 };

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

// ---------------------------------------------------------------- in-process backend
struct LocalMsg { const void* ptr; size_t bytes; };
struct LocalGroup {
    int n = 0;
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::deque<LocalMsg>> box;       // [src * n + dst], FIFO
    std::vector<uint64_t> posted, done;          // [src * n + dst]
    std::vector<void*> arPtr;                    // allreduce operands, by rank
    int arrived = 0;
    uint64_t generation = 0;
    std::atomic<int> refs{0};
    bool failed = false;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = generation;
        if (++arrived == n) { arrived = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != g || failed; });
        if (failed) throw Error(FLIPB200_ERR_COMM, "local communicator: a peer failed");
    }
    void abort() {
        { std::lock_guard<std::mutex> lk(m); failed = true; }
        cv.notify_all();
    }
};
struct PendingOp { int peer; const void* sbuf; void* rbuf; size_t bytes; bool isSend; };

struct Comm {
    ncclComm_t comm = nullptr;
    LocalGroup* grp = nullptr;
    std::vector<PendingOp> pending;   // local backend: ops of the open group
    DBuf<unsigned char> arTmp;
};

template <typename T, bool IS_MAX>
__global__ void local_reduce_kernel(const T* p0, const T* p1, const T* p2, const T* p3, const T* p4, const T* p5, const T* p6,
                                    const T* p7, int n, size_t count, T* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const T* ps[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
    T a = ps[0][i];
    for (int r = 1; r < n; r++) { T v = ps[r][i]; a = IS_MAX ? (v > a ? v : a) : (T)(a + v); }
    out[i] = a;
}

void comm_destroy(World* w) {
    if (w->comm) {
        if (w->comm->comm) nccl().CommDestroy(w->comm->comm);
        if (w->comm->grp && --w->comm->grp->refs == 0) delete w->comm->grp;
        delete w->comm;
        w->comm = nullptr;
    }
    w->rank = 0; w->nRanks = 1;
}
bool comm_active(World* w) { return w->comm && w->nRanks > 1; }
bool comm_is_local(World* w) { return w->comm && w->comm->grp != nullptr; }

void comm_group_begin(World* w) {
    FB_REQUIRE(comm_active(w), FLIPB200_ERR_COMM, "communicator not initialised");
    if (w->comm->comm) FB_NCCL(nccl().GroupStart());
    else w->comm->pending.clear();
}
void comm_send(World* w, int peer, const void* buf, size_t bytes) {
    if (bytes == 0) return;
    if (w->comm->comm) FB_NCCL(nccl().Send(buf, bytes, NCCL_UINT8, peer, w->comm->comm, w->stream));
    else w->comm->pending.push_back(PendingOp{peer, buf, nullptr, bytes, true});
}
void comm_recv(World* w, int peer, void* buf, size_t bytes) {
    if (bytes == 0) return;
    if (w->comm->comm) FB_NCCL(nccl().Recv(buf, bytes, NCCL_UINT8, peer, w->comm->comm, w->stream));
    else w->comm->pending.push_back(PendingOp{peer, nullptr, buf, bytes, false});
}
void comm_group_end(World* w) {
    w->launches++;
    if (w->comm->comm) { FB_NCCL(nccl().GroupEnd()); return; }
    LocalGroup& G = *w->comm->grp;
    const int me = w->rank, n = G.n;
    FB_CUDA(cudaStreamSynchronize(w->stream));   // everything this rank sends is complete
    {
        std::lock_guard<std::mutex> lk(G.m);
        for (auto& op : w->comm->pending)
            if (op.isSend) { G.box[me * n + op.peer].push_back(LocalMsg{op.sbuf, op.bytes}); G.posted[me * n + op.peer]++; }
    }
    G.cv.notify_all();
    std::vector<int> took(n, 0);
    for (auto& op : w->comm->pending) {
        if (op.isSend) continue;
        LocalMsg msg;
        {
            std::unique_lock<std::mutex> lk(G.m);
            auto& q = G.box[op.peer * n + me];
            G.cv.wait(lk, [&] { return !q.empty() || G.failed; });
            FB_REQUIRE(!G.failed, FLIPB200_ERR_COMM, "local communicator: a peer failed");
            msg = q.front();
            q.pop_front();
        }
        if (msg.bytes != op.bytes) {
            { std::lock_guard<std::mutex> lk(G.m); G.failed = true; }
            G.cv.notify_all();
            throw Error(FLIPB200_ERR_COMM, "local communicator: message size mismatch (" + std::to_string(msg.bytes) + " sent, " + std::to_string(op.bytes) + " expected)");
        }
        FB_CUDA(cudaMemcpyAsync(op.rbuf, msg.ptr, op.bytes, cudaMemcpyDefault, w->stream));
        took[op.peer]++;
    }
    FB_CUDA(cudaStreamSynchronize(w->stream));
    {
        std::unique_lock<std::mutex> lk(G.m);
        for (int r = 0; r < n; r++) G.done[r * n + me] += took[r];
        G.cv.notify_all();
        // my send buffers may be reused once every message I posted has been copied out
        G.cv.wait(lk, [&] {
            if (G.failed) return true;
            for (int r = 0; r < n; r++) if (G.done[me * n + r] < G.posted[me * n + r]) return false;
            return true;
        });
        FB_REQUIRE(!G.failed, FLIPB200_ERR_COMM, "local communicator: a peer failed");
    }
    w->comm->pending.clear();
}

void comm_allreduce(World* w, void* buf, size_t n, int type, bool isMax) {
    if (!comm_active(w) || n == 0) return;
    w->launches++;
    if (w->comm->comm) {
        const int dt = type == CT_F32 ? NCCL_FLOAT32 : (type == CT_U32 ? NCCL_UINT32 : NCCL_INT32);
        FB_NCCL(nccl().AllReduce(buf, buf, n, dt, isMax ? NCCL_MAX : NCCL_SUM, w->comm->comm, w->stream));
        return;
    }
    LocalGroup& G = *w->comm->grp;
    FB_REQUIRE(G.n <= 8, FLIPB200_ERR_COMM, "local communicator supports at most 8 ranks");
    FB_CUDA(cudaStreamSynchronize(w->stream));
    { std::lock_guard<std::mutex> lk(G.m); G.arPtr[w->rank] = buf; }
    G.barrier();
    if (w->comm->arTmp.n < n * 4) w->comm->arTmp.alloc(n * 4, w->stream);
    void* p[8];
    for (int r = 0; r < 8; r++) p[r] = G.arPtr[r < G.n ? r : 0];
    const unsigned blocks = (unsigned)((n + 255) / 256);
#define FB_LR(T, MX) local_reduce_kernel<T, MX><<<blocks, 256, 0, w->stream>>>((const T*)p[0], (const T*)p[1], (const T*)p[2], (const T*)p[3], (const T*)p[4], (const T*)p[5], (const T*)p[6], (const T*)p[7], G.n, n, (T*)w->comm->arTmp.p)
    if (type == CT_F32) { if (isMax) FB_LR(float, true); else FB_LR(float, false); }
    else if (type == CT_U32) { if (isMax) FB_LR(uint32_t, true); else FB_LR(uint32_t, false); }
    else { if (isMax) FB_LR(int32_t, true); else FB_LR(int32_t, false); }
#undef FB_LR
    check_launch("local_reduce");
    FB_CUDA(cudaStreamSynchronize(w->stream));
    G.barrier();   // every rank has read every operand
    FB_CUDA(cudaMemcpyAsync(buf, w->comm->arTmp.p, n * 4, cudaMemcpyDeviceToDevice, w->stream));
    FB_CUDA(cudaStreamSynchronize(w->stream));
    G.barrier();
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
int flipb200_comm_abort(flipb200_world* w) {
    // wakes every rank of an in-process group that is blocked in a collective (a peer hit an error)
    if (w && w->comm && w->comm->grp) w->comm->grp->abort();
    return FLIPB200_OK;
}
int flipb200_comm_init_local(flipb200_world** worlds, int n) {
    // ranks 0..n-1 = the given worlds, all living in this process (one host thread drives each)
    try {
        if (!worlds || n < 1 || n > 8) return FLIPB200_ERR_ARG;
        auto* G = new fb::LocalGroup();
        G->n = n;
        G->box.resize((size_t)n * n);
        G->posted.assign((size_t)n * n, 0);
        G->done.assign((size_t)n * n, 0);
        G->arPtr.assign(8, nullptr);
        G->refs = n;
        {
            // The worlds of one process share the device's default stream-ordered pool. By default the pool may hand a block that
            // one world's stream has freed-but-not-yet-reached to another world's stream and insert a dependency on the freeing
            // stream. Worlds that exchange ghost leaves through peer memory WAIT for each other on the device (dd_wait_kernel), so
            // such a dependency can close a cycle (A waits for B's push, B's push waits for A's free point behind A's wait).
            // Reuse across streams stays allowed once the free has completed (opportunistic reuse).
            cudaMemPool_t pool;
            int dev = worlds[0]->device, zero = 0;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &zero);
            cudaGetLastError();
        }
        for (int r = 0; r < n; r++) {
            flipb200_world* w = worlds[r];
            cudaSetDevice(w->device);
            fb::comm_destroy(w);
            w->comm = new fb::Comm();
            w->comm->grp = G;
            w->rank = r; w->nRanks = n;
        }
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
