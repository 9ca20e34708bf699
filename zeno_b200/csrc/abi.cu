// libflipb200 -- the extern "C" boundary (include/flipb200.h): handle management, VDB-layout
// marshalling, error translation, profiling hooks, and the device-resident substep.
#include "world.cuh"
#include <cstring>
#include <algorithm>

using namespace fb;

namespace {
thread_local std::string g_lastError;

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return FLIPB200_OK;
    } catch (const fb::Error& e) {
        g_lastError = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return FLIPB200_ERR_ARG;
    }
}
void use_device(flipb200_world* w) { FB_CUDA(cudaSetDevice(w->device)); }
}  // namespace
namespace fb { void set_last_error(const std::string& m) { g_lastError = m; } }
namespace {

__global__ void aos_to_soa_kernel(const float* __restrict__ in, float* __restrict__ c0, float* __restrict__ c1,
                                  float* __restrict__ c2, size_t nVox) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVox) return;
    c0[i] = in[3 * i]; c1[i] = in[3 * i + 1]; c2[i] = in[3 * i + 2];
}
__global__ void soa_to_aos_kernel(const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ c2,
                                  float* __restrict__ out, size_t nVox) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVox) return;
    out[3 * i] = c0[i]; out[3 * i + 1] = c1[i]; out[3 * i + 2] = c2[i];
}
// upload helpers: scatter caller-ordered leaves into topology slots
__global__ void scatter_leaves_kernel(TopoView t, const int3* __restrict__ origins, int nIn, const float* __restrict__ in,
                                      int inStrideLeaf, int inOffset, float* __restrict__ out,
                                      const uint64_t* __restrict__ inMask, uint64_t* __restrict__ outMask,
                                      uint8_t* __restrict__ alloc) {
    int l = blockIdx.x;
    int3 o = origins[l];
    int s = topo_find(t, o.x, o.y, o.z);
    if (s < 0) return;
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x) out[(size_t)s * LEAF + i] = in[(size_t)l * inStrideLeaf + inOffset + i];
    if (inMask && threadIdx.x < 8) outMask[(size_t)s * 8 + threadIdx.x] = inMask[(size_t)l * 8 + threadIdx.x];
    if (alloc && threadIdx.x == 0) alloc[s] = 1;
}
__global__ void select_leaves_kernel(int n, const uint64_t* __restrict__ mask, const uint8_t* __restrict__ alloc,
                                     uint32_t* __restrict__ flag) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    uint64_t a = alloc ? alloc[l] : 0;
    for (int k = 0; k < 8; k++) a |= mask[(size_t)l * 8 + k];
    flag[l] = a ? 1u : 0u;
}
__global__ void compact_leaves_kernel(TopoView t, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                      const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ c2,
                                      int nch, const uint64_t* __restrict__ mask, int3* __restrict__ oOut,
                                      uint64_t* __restrict__ mOut, float* __restrict__ vOut) {
    int l = blockIdx.x;
    if (!flag[l]) return;
    uint32_t d = pos[l];
    if (threadIdx.x == 0) oOut[d] = t.origin[l];
    if (threadIdx.x < 8) mOut[(size_t)d * 8 + threadIdx.x] = mask[(size_t)l * 8 + threadIdx.x];
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x) {
        vOut[((size_t)d * nch + 0) * LEAF + i] = c0[(size_t)l * LEAF + i];
        if (nch == 3) {
            vOut[((size_t)d * nch + 1) * LEAF + i] = c1[(size_t)l * LEAF + i];
            vOut[((size_t)d * nch + 2) * LEAF + i] = c2[(size_t)l * LEAF + i];
        }
    }
}
// particle store <-> reference layout
__global__ void pts_pack_kernel(const uint16_t* __restrict__ P, const uint16_t* __restrict__ v, uint64_t n,
                                uint32_t* __restrict__ w0, uint32_t* __restrict__ w1, uint32_t* __restrict__ w2) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    w0[i] = (uint32_t)P[3 * i] | ((uint32_t)P[3 * i + 1] << 16);
    w1[i] = (uint32_t)P[3 * i + 2] | ((uint32_t)v[3 * i] << 16);
    w2[i] = (uint32_t)v[3 * i + 1] | ((uint32_t)v[3 * i + 2] << 16);
}
__global__ void pts_unpack_kernel(const uint32_t* __restrict__ w0, const uint32_t* __restrict__ w1, const uint32_t* __restrict__ w2,
                                  uint64_t n, uint16_t* __restrict__ P, uint16_t* __restrict__ v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = w0[i], b = w1[i], c = w2[i];
    P[3 * i] = (uint16_t)(a & 0xffffu); P[3 * i + 1] = (uint16_t)(a >> 16); P[3 * i + 2] = (uint16_t)(b & 0xffffu);
    v[3 * i] = (uint16_t)(b >> 16); v[3 * i + 1] = (uint16_t)(c & 0xffffu); v[3 * i + 2] = (uint16_t)(c >> 16);
}
// caller's per-leaf cumulative ends -> per-voxel counts in slot order
__global__ void pts_counts_kernel(TopoView t, const int3* __restrict__ origins, const uint32_t* __restrict__ voxelEnd,
                                  uint32_t* __restrict__ counts, uint32_t* __restrict__ leafTotal) {
    int l = blockIdx.x;
    int3 o = origins[l];
    int s = topo_find(t, o.x, o.y, o.z);
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x) {
        uint32_t e = voxelEnd[(size_t)l * LEAF + i];
        uint32_t b = i == 0 ? 0u : voxelEnd[(size_t)l * LEAF + i - 1];
        counts[(size_t)s * LEAF + i] = e - b;
    }
    if (threadIdx.x == 0) leafTotal[l] = voxelEnd[(size_t)l * LEAF + LEAF - 1];
}
// move each caller leaf's particles to where the slot-ordered store expects them
__global__ void pts_place_kernel(TopoView t, const int3* __restrict__ origins, const uint32_t* __restrict__ inLeafStart,
                                 const uint32_t* __restrict__ voxelStart, const uint32_t* __restrict__ i0,
                                 const uint32_t* __restrict__ i1, const uint32_t* __restrict__ i2, uint32_t* __restrict__ o0,
                                 uint32_t* __restrict__ o1, uint32_t* __restrict__ o2) {
    int l = blockIdx.x;
    int3 o = origins[l];
    int s = topo_find(t, o.x, o.y, o.z);
    uint32_t src = inLeafStart[l], cnt = inLeafStart[l + 1] - src;
    uint32_t dst = voxelStart[(size_t)s * LEAF];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) { o0[dst + i] = i0[src + i]; o1[dst + i] = i1[src + i]; o2[dst + i] = i2[src + i]; }
}
__global__ void pts_export_kernel(TopoView t, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                  const uint32_t* __restrict__ voxelStart, int3* __restrict__ oOut, uint32_t* __restrict__ veOut) {
    int l = blockIdx.x;
    if (!flag[l]) return;
    uint32_t d = pos[l];
    if (threadIdx.x == 0) oOut[d] = t.origin[l];
    uint32_t base = voxelStart[(size_t)l * LEAF];
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x) veOut[(size_t)d * LEAF + i] = voxelStart[(size_t)l * LEAF + i + 1] - base;
}
__global__ void pts_leaf_flag_kernel(int n, const uint32_t* __restrict__ voxelStart, uint32_t* __restrict__ flag) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    flag[l] = voxelStart[(size_t)(l + 1) * LEAF] > voxelStart[(size_t)l * LEAF] ? 1u : 0u;
}
inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

// Where the device->host copies of a download go: the world's stream followed by a sync (blocking calls), or the
// copy stream behind an event recorded after the staging kernels (the *_begin calls).
struct Copier {
    flipb200_world* w; bool async; bool armed = false;
    template <typename T> T* hold(DBuf<T>&& b) {
        auto sp = std::make_shared<DBuf<T>>(std::move(b));
        T* p = sp->p;
        w->held.push_back(sp);
        return p;
    }
    void arm() {
        if (!async || armed) return;
        if (!w->copyStream) {
            FB_CUDA(cudaStreamCreateWithFlags(&w->copyStream, cudaStreamNonBlocking));
            FB_CUDA(cudaEventCreateWithFlags(&w->copyEvt, cudaEventDisableTiming));
        }
        FB_CUDA(cudaEventRecord(w->copyEvt, w->stream));
        FB_CUDA(cudaStreamWaitEvent(w->copyStream, w->copyEvt, 0));
        armed = true;
    }
    void d2h(void* dst, const void* src, size_t bytes) {
        if (!bytes) return;
        arm();
        FB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, async ? w->copyStream : w->stream));
    }
    void finish() { if (!async) sync(w); }
};

void resolve_profile(flipb200_world* w) {
    if (w->pending.empty()) return;
    cudaStreamSynchronize(w->stream);
    for (auto& e : w->pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        auto& p = w->prof[e.name];
        p.ms += ms; p.launches++; p.bytes += e.bytes;
        w->evtPool.push_back(e.a);
        w->evtPool.push_back(e.b);
    }
    w->pending.clear();
}
}  // namespace

extern "C" {

const char* flipb200_last_error(void) { return g_lastError.c_str(); }
const char* flipb200_build_info(void) {
    return "libflipb200 abi=1 arch=sm_100a cuda=" FB_STR(CUDART_VERSION) " fmad=off";
}
int flipb200_abi_version(void) { return 1; }
int flipb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int flipb200_world_create(int device, float dx, flipb200_world** out) {
    return guarded([&] {
        FB_REQUIRE(out != nullptr && dx > 0.f, FLIPB200_ERR_ARG, "world_create: bad argument");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) { cudaGetLastError(); throw fb::Error(FLIPB200_ERR_CUDA, "no CUDA device: libflipb200 has no CPU fallback"); }
        FB_REQUIRE(device >= 0 && device < n, FLIPB200_ERR_ARG, "world_create: bad device index");
        FB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        FB_CUDA(cudaGetDeviceProperties(&prop, device));
        FB_REQUIRE(prop.major == 10, FLIPB200_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) + ", libflipb200 is built for sm_100a only");
        auto* w = new flipb200_world();
        w->device = device;
        w->dx = dx;
        FB_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
        FB_CUDA(cudaHostAlloc((void**)&w->hostScratch, 1024, cudaHostAllocMapped));
        memset(w->hostScratch, 0, 1024);
        FB_CUDA(cudaHostGetDevicePointer((void**)&w->hostScratchDev, w->hostScratch, 0));
        // keep freed blocks in the stream-ordered pool: per-substep temporaries are recycled
        cudaMemPool_t pool;
        FB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = ~0ull;
        FB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        // SetFLIPWorld backgrounds (FF/nosys/FLIP_Creator.cpp:85,95)
        w->F(FLIPB200_LIQUID_SDF).bg = 1.0f * dx;
        w->F(FLIPB200_SOLID_SDF).bg = 3.0f * dx;
        *out = w;
    });
}
int flipb200_world_destroy(flipb200_world* w) {
    return guarded([&] {
        if (!w) return;
        cudaSetDevice(w->device);
        cudaStreamSynchronize(w->stream);
        if (!w->phase.empty()) {
            fprintf(stderr, "[flipb200 rank %d] host phases (wall ms total / calls):\n", w->rank);
            for (auto& kv : w->phase) fprintf(stderr, "  %-28s %10.3f ms  %6llu calls  %8.3f ms/call\n", kv.first.c_str(), kv.second.ms, (unsigned long long)kv.second.launches, kv.second.ms / kv.second.launches);
        }
        comm_destroy(w);
        dd_destroy(w);
        resolve_profile(w);
        for (auto e : w->evtPool) cudaEventDestroy(e);
        if (w->hostScratch) cudaFreeHost(w->hostScratch);
        if (w->copyStream) { cudaStreamSynchronize(w->copyStream); w->held.clear(); cudaStreamDestroy(w->copyStream); cudaEventDestroy(w->copyEvt); }
        cudaStream_t s = w->stream;
        delete w;
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
    });
}

int flipb200_host_alloc(size_t bytes, void** out) {
    return guarded([&] {
        FB_REQUIRE(out != nullptr, FLIPB200_ERR_ARG, "host_alloc: null out");
        *out = nullptr;
        if (bytes) FB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    });
}
int flipb200_host_free(void* p) {
    return guarded([&] { if (p) FB_CUDA(cudaFreeHost(p)); });
}

int flipb200_grid_upload(flipb200_world* w, int grid, int nLeaves, const int32_t* origins, const uint64_t* masks,
                         const float* values, int layout, const float* background) {
    return guarded([&] {
        FB_REQUIRE(w && (is_vec_grid(grid) || is_float_grid(grid)) && nLeaves >= 0, FLIPB200_ERR_ARG, "grid_upload: bad argument");
        use_device(w);
        const int nch = is_vec_grid(grid) ? 3 : 1;
        TopoPtr t = topo_from_origins_host(w, origins, nLeaves, false);
        FB_REQUIRE(t->n == nLeaves, FLIPB200_ERR_ARG, "grid_upload: duplicate leaf origins");
        DBuf<int3> o(nLeaves + 1, w->stream);
        DBuf<uint64_t> m((size_t)nLeaves * 8 + 1, w->stream);
        DBuf<float> v((size_t)nLeaves * nch * LEAF + 1, w->stream);
        if (nLeaves) {
            FB_CUDA(cudaMemcpyAsync(o.p, origins, sizeof(int3) * (size_t)nLeaves, cudaMemcpyHostToDevice, w->stream));
            FB_CUDA(cudaMemcpyAsync(m.p, masks, 64 * (size_t)nLeaves, cudaMemcpyHostToDevice, w->stream));
            FB_CUDA(cudaMemcpyAsync(v.p, values, sizeof(float) * (size_t)nLeaves * nch * LEAF, cudaMemcpyHostToDevice, w->stream));
        }
        if (nch == 1) {
            GridF g;
            grid_alloc(w, g, t, background ? background[0] : w->F(grid).bg);
            if (nLeaves) {
                FB_LAUNCH(w, "upload_scatter", (size_t)nLeaves * 4096) scatter_leaves_kernel<<<nLeaves, 256, 0, w->stream>>>(t->view(), o.p, nLeaves, v.p, LEAF, 0, g.val.p, m.p, g.mask.p, g.alloc.p);
                check_launch("upload_scatter");
            }
            w->F(grid) = std::move(g);
        } else {
            GridV g;
            float bg[3] = {background ? background[0] : w->V(grid).bg[0], background ? background[1] : w->V(grid).bg[1], background ? background[2] : w->V(grid).bg[2]};
            grid_alloc(w, g, t, bg);
            DBuf<float> soa;
            const float* src = v.p;
            if (layout == FLIPB200_AOS && nLeaves) {
                // [leaf][512][3] -> [leaf][3][512] is done per channel below through a strided view:
                // first de-interleave to three planar arrays
                size_t nVox = (size_t)nLeaves * LEAF;
                soa.alloc(3 * nVox, w->stream);
                FB_LAUNCH(w, "upload_aos_to_soa", nVox * 24) aos_to_soa_kernel<<<nblk(nVox, 256), 256, 0, w->stream>>>(v.p, soa.p, soa.p + nVox, soa.p + 2 * nVox, nVox);
                check_launch("aos_to_soa");
                for (int c = 0; c < 3; c++) {
                    FB_LAUNCH(w, "upload_scatter", (size_t)nLeaves * 4096) scatter_leaves_kernel<<<nLeaves, 256, 0, w->stream>>>(t->view(), o.p, nLeaves, soa.p + c * nVox, LEAF, 0, g.val[c].p, c == 0 ? m.p : nullptr, g.mask.p, nullptr);
                    check_launch("upload_scatter");
                }
            } else if (nLeaves) {
                for (int c = 0; c < 3; c++) {
                    FB_LAUNCH(w, "upload_scatter", (size_t)nLeaves * 4096) scatter_leaves_kernel<<<nLeaves, 256, 0, w->stream>>>(t->view(), o.p, nLeaves, src, 3 * LEAF, c * LEAF, g.val[c].p, c == 0 ? m.p : nullptr, g.mask.p, nullptr);
                    check_launch("upload_scatter");
                }
            }
            w->V(grid) = std::move(g);
        }
        if (grid == FLIPB200_SOLID_SDF) { w->hasSolidSDF = true; w->solidViewEpoch = ~0ull; }
        if (grid == FLIPB200_SOLID_VELOCITY) { w->hasSolidVel = true; w->solidViewEpoch = ~0ull; }
        sync(w);
    });
}

static void select_leaves(flipb200_world* w, int grid, DBuf<uint32_t>& flag, DBuf<uint32_t>& pos, int* count) {
    TopoPtr t = is_vec_grid(grid) ? w->V(grid).topo : w->F(grid).topo;
    int n = t ? t->n : 0;
    flag.alloc(n + 1, w->stream);
    pos.alloc(n + 1, w->stream);
    flag.zero();
    if (n) {
        const uint64_t* m = is_vec_grid(grid) ? w->V(grid).mask.p : w->F(grid).mask.p;
        const uint8_t* al = is_vec_grid(grid) ? nullptr : w->F(grid).alloc.p;
        // static grids keep every uploaded leaf
        FB_LAUNCH(w, "download_select", (size_t)n * 70) select_leaves_kernel<<<nblk(n, 128), 128, 0, w->stream>>>(n, m, al, flag.p);
        check_launch("select_leaves");
    }
    uint64_t total = 0;
    exclusive_scan_u32(w, flag.p, pos.p, n + 1, &total);
    *count = (int)total;
}

int flipb200_grid_leaf_count(flipb200_world* w, int grid, int* nLeaves) {
    return guarded([&] {
        FB_REQUIRE(w && nLeaves && (is_vec_grid(grid) || is_float_grid(grid)), FLIPB200_ERR_ARG, "grid_leaf_count: bad argument");
        use_device(w);
        DBuf<uint32_t> flag, pos;
        select_leaves(w, grid, flag, pos, nLeaves);
    });
}
static void grid_download_impl(flipb200_world* w, int grid, int32_t* origins, uint64_t* masks, float* values, int layout,
                               float* background, bool async, int* nLeavesOut, int capLeaves = -1) {
    FB_REQUIRE(w && (is_vec_grid(grid) || is_float_grid(grid)), FLIPB200_ERR_ARG, "grid_download: bad argument");
    use_device(w);
    Copier cp{w, async};
    const int nch = is_vec_grid(grid) ? 3 : 1;
    if (background) {
        if (nch == 1) background[0] = w->F(grid).bg;
        else for (int c = 0; c < 3; c++) background[c] = w->V(grid).bg[c];
    }
    DBuf<uint32_t> flag, pos;
    int cnt = 0;
    select_leaves(w, grid, flag, pos, &cnt);
    if (nLeavesOut) *nLeavesOut = cnt;
    FB_REQUIRE(capLeaves < 0 || cnt <= capLeaves, FLIPB200_ERR_ARG, "grid_download_begin: " + std::to_string(cnt) + " leaves do not fit the caller's buffers");
    if (cnt == 0) return;
    TopoPtr t = nch == 3 ? w->V(grid).topo : w->F(grid).topo;
    DBuf<int3> o(cnt, w->stream);
    DBuf<uint64_t> m((size_t)cnt * 8, w->stream);
    DBuf<float> v((size_t)cnt * nch * LEAF, w->stream);
    const float *c0, *c1 = nullptr, *c2 = nullptr;
    const uint64_t* mk;
    if (nch == 3) { c0 = w->V(grid).val[0].p; c1 = w->V(grid).val[1].p; c2 = w->V(grid).val[2].p; mk = w->V(grid).mask.p; }
    else { c0 = w->F(grid).val.p; mk = w->F(grid).mask.p; }
    FB_LAUNCH(w, "download_compact", (size_t)cnt * nch * 4096) compact_leaves_kernel<<<t->n, 256, 0, w->stream>>>(t->view(), flag.p, pos.p, c0, c1, c2, nch, mk, o.p, m.p, v.p);
    check_launch("compact_leaves");
    const float* vsrc = v.p;
    if (nch == 3 && layout == FLIPB200_AOS) {
        // [leaf][3][512] -> [leaf][512][3]: per leaf transposition on the device
        size_t nVox = (size_t)cnt * LEAF;
        DBuf<float> planar(3 * nVox, w->stream), aos(3 * nVox, w->stream);
        for (int c = 0; c < 3; c++)
            FB_CUDA(cudaMemcpy2DAsync(planar.p + c * nVox, LEAF * 4, v.p + c * LEAF, 3 * LEAF * 4, LEAF * 4, cnt, cudaMemcpyDeviceToDevice, w->stream));
        FB_LAUNCH(w, "download_soa_to_aos", nVox * 24) soa_to_aos_kernel<<<nblk(nVox, 256), 256, 0, w->stream>>>(planar.p, planar.p + nVox, planar.p + 2 * nVox, aos.p, nVox);
        check_launch("soa_to_aos");
        vsrc = async ? cp.hold(std::move(aos)) : aos.p;
        if (!async) {   // aos dies at the end of this scope: copy + sync here
            cp.d2h(values, vsrc, sizeof(float) * 3 * nVox);
            cp.d2h(origins, o.p, sizeof(int3) * (size_t)cnt);
            cp.d2h(masks, m.p, 64 * (size_t)cnt);
            cp.finish();
            return;
        }
    }
    const int3* op = async ? cp.hold(std::move(o)) : o.p;
    const uint64_t* mp = async ? cp.hold(std::move(m)) : m.p;
    if (async && vsrc == v.p) vsrc = cp.hold(std::move(v));
    cp.d2h(origins, op, sizeof(int3) * (size_t)cnt);
    cp.d2h(masks, mp, 64 * (size_t)cnt);
    cp.d2h(values, vsrc, sizeof(float) * (size_t)cnt * nch * LEAF);
    cp.finish();
}
int flipb200_grid_download(flipb200_world* w, int grid, int32_t* origins, uint64_t* masks, float* values, int layout,
                           float* background) {
    return guarded([&] { grid_download_impl(w, grid, origins, masks, values, layout, background, false, nullptr); });
}
int flipb200_grid_download_begin(flipb200_world* w, int grid, int capLeaves, int32_t* origins, uint64_t* masks, float* values,
                                 int layout, float* background, int* nLeaves) {
    return guarded([&] { grid_download_impl(w, grid, origins, masks, values, layout, background, true, nLeaves, capLeaves < 0 ? 0 : capLeaves); });
}
int flipb200_download_wait(flipb200_world* w) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "download_wait: null world");
        use_device(w);
        if (w->copyStream) FB_CUDA(cudaStreamSynchronize(w->copyStream));
        w->held.clear();   // staging buffers go back to the pool in stream order of the world's stream
    });
}

int flipb200_particles_upload(flipb200_world* w, int nLeaves, const int32_t* origins, const uint32_t* voxelEnd,
                              uint64_t nParticles, const uint16_t* P, const uint16_t* v) {
    return guarded([&] {
        FB_REQUIRE(w && nLeaves >= 0, FLIPB200_ERR_ARG, "particles_upload: bad argument");
        use_device(w);
        reserve_pool(w, ((uint64_t)1 << 30) + 512 * nParticles);
        TopoPtr pool = topo_from_origins_host(w, origins, nLeaves, true);
        const uint64_t n = nParticles;
        DBuf<int3> o(nLeaves + 1, w->stream);
        DBuf<uint32_t> ve((size_t)nLeaves * LEAF + 1, w->stream), leafTot(nLeaves + 1, w->stream);
        DBuf<uint16_t> dP(3 * n + 1, w->stream), dv(3 * n + 1, w->stream);
        DBuf<uint32_t> i0(n + 1, w->stream), i1(n + 1, w->stream), i2(n + 1, w->stream);
        Particles out;
        out.topo = pool; out.n = n;
        size_t nv = (size_t)pool->n * LEAF;
        out.voxelStart.alloc(nv + 1, w->stream);
        out.voxelStart.zero();
        out.w0.alloc(n + 1, w->stream); out.w1.alloc(n + 1, w->stream); out.w2.alloc(n + 1, w->stream);
        leafTot.zero();
        if (nLeaves) {
            FB_CUDA(cudaMemcpyAsync(o.p, origins, sizeof(int3) * (size_t)nLeaves, cudaMemcpyHostToDevice, w->stream));
            FB_CUDA(cudaMemcpyAsync(ve.p, voxelEnd, 4 * (size_t)nLeaves * LEAF, cudaMemcpyHostToDevice, w->stream));
        }
        if (n) {
            FB_CUDA(cudaMemcpyAsync(dP.p, P, 6 * n, cudaMemcpyHostToDevice, w->stream));
            FB_CUDA(cudaMemcpyAsync(dv.p, v, 6 * n, cudaMemcpyHostToDevice, w->stream));
            FB_LAUNCH(w, "pts_pack", n * 24) pts_pack_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(dP.p, dv.p, n, i0.p, i1.p, i2.p);
            check_launch("pts_pack");
        }
        if (nLeaves) {
            FB_LAUNCH(w, "pts_counts", (size_t)nLeaves * 4096) pts_counts_kernel<<<nLeaves, 256, 0, w->stream>>>(pool->view(), o.p, ve.p, out.voxelStart.p, leafTot.p);
            check_launch("pts_counts");
        }
        uint64_t total = 0;
        exclusive_scan_u32(w, out.voxelStart.p, out.voxelStart.p, nv + 1, nullptr);
        exclusive_scan_u32(w, leafTot.p, leafTot.p, nLeaves + 1, &total);
        FB_REQUIRE(total == n, FLIPB200_ERR_ARG, "particles_upload: voxelEnd totals do not match nParticles");
        if (nLeaves && n) {
            FB_LAUNCH(w, "pts_place", n * 24) pts_place_kernel<<<nLeaves, 256, 0, w->stream>>>(pool->view(), o.p, leafTot.p, out.voxelStart.p, i0.p, i1.p, i2.p, out.w0.p, out.w1.p, out.w2.p);
            check_launch("pts_place");
        }
        w->pts = std::move(out);
        w->pool = pool;
        sync(w);
    });
}
int flipb200_particles_info(flipb200_world* w, int* nLeaves, uint64_t* nParticles) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "particles_info: bad argument");
        use_device(w);
        if (nParticles) *nParticles = w->pts.n;
        if (nLeaves) {
            *nLeaves = 0;
            if (w->pts.topo && w->pts.topo->n) {
                int n = w->pts.topo->n;
                DBuf<uint32_t> flag(n + 1, w->stream), pos(n + 1, w->stream);
                flag.zero();
                pts_leaf_flag_kernel<<<nblk(n, 128), 128, 0, w->stream>>>(n, w->pts.voxelStart.p, flag.p);
                w->launches++;
                uint64_t total = 0;
                exclusive_scan_u32(w, flag.p, pos.p, n + 1, &total);
                *nLeaves = (int)total;
            }
        }
    });
}
static void particles_download_impl(flipb200_world* w, int32_t* origins, uint32_t* voxelEnd, uint16_t* P, uint16_t* v, bool async,
                                    int* nLeavesOut, uint64_t* nParticlesOut, int64_t capLeaves = -1, int64_t capParticles = -1) {
    FB_REQUIRE(w, FLIPB200_ERR_ARG, "particles_download: bad argument");
    use_device(w);
    Copier cp{w, async};
    if (nLeavesOut) *nLeavesOut = 0;
    if (nParticlesOut) *nParticlesOut = 0;
    if (!w->pts.topo || w->pts.topo->n == 0) return;
    int n = w->pts.topo->n;
    uint64_t np = w->pts.n;
    DBuf<uint32_t> flag(n + 1, w->stream), pos(n + 1, w->stream);
    flag.zero();
    pts_leaf_flag_kernel<<<nblk(n, 128), 128, 0, w->stream>>>(n, w->pts.voxelStart.p, flag.p);
    w->launches++;
    uint64_t total = 0;
    exclusive_scan_u32(w, flag.p, pos.p, n + 1, &total);
    if (nLeavesOut) *nLeavesOut = (int)total;
    if (nParticlesOut) *nParticlesOut = np;
    FB_REQUIRE(capLeaves < 0 || ((int64_t)total <= capLeaves && (int64_t)np <= capParticles), FLIPB200_ERR_ARG,
               "particles_download_begin: " + std::to_string(total) + " leaves / " + std::to_string(np) + " particles do not fit the caller's buffers");
    if (total) {
        DBuf<int3> o(total, w->stream);
        DBuf<uint32_t> ve((size_t)total * LEAF, w->stream);
        FB_LAUNCH(w, "pts_export", total * 4096) pts_export_kernel<<<n, 256, 0, w->stream>>>(w->pts.topo->view(), flag.p, pos.p, w->pts.voxelStart.p, o.p, ve.p);
        check_launch("pts_export");
        const int3* op = async ? cp.hold(std::move(o)) : o.p;
        const uint32_t* vp = async ? cp.hold(std::move(ve)) : ve.p;
        cp.d2h(origins, op, sizeof(int3) * total);
        cp.d2h(voxelEnd, vp, 4 * (size_t)total * LEAF);
        cp.finish();
    }
    if (np) {
        DBuf<uint16_t> dP(3 * np, w->stream), dv(3 * np, w->stream);
        FB_LAUNCH(w, "pts_unpack", np * 24) pts_unpack_kernel<<<nblk(np, 256), 256, 0, w->stream>>>(w->pts.w0.p, w->pts.w1.p, w->pts.w2.p, np, dP.p, dv.p);
        check_launch("pts_unpack");
        cp.armed = false;   // the copies below must follow the unpack kernel
        const uint16_t* pp = async ? cp.hold(std::move(dP)) : dP.p;
        const uint16_t* vp = async ? cp.hold(std::move(dv)) : dv.p;
        cp.d2h(P, pp, 6 * np);
        cp.d2h(v, vp, 6 * np);
        cp.finish();
    }
}
int flipb200_particles_download(flipb200_world* w, int32_t* origins, uint32_t* voxelEnd, uint16_t* P, uint16_t* v) {
    return guarded([&] { particles_download_impl(w, origins, voxelEnd, P, v, false, nullptr, nullptr); });
}
int flipb200_particles_download_begin(flipb200_world* w, int capLeaves, uint64_t capParticles, int32_t* origins, uint32_t* voxelEnd,
                                      uint16_t* P, uint16_t* v, int* nLeaves, uint64_t* nParticles) {
    return guarded([&] { particles_download_impl(w, origins, voxelEnd, P, v, true, nLeaves, nParticles, capLeaves < 0 ? 0 : capLeaves, (int64_t)capParticles); });
}

int flipb200_bin_from_points(flipb200_world* w, const float* pos, const float* vel, uint64_t n) {
    return guarded([&] {
        FB_REQUIRE(w && (pos || n == 0), FLIPB200_ERR_ARG, "bin_from_points: bad argument");
        use_device(w);
        bin_from_points(w, pos, vel, n);
        sync(w);
    });
}
int flipb200_p2g(flipb200_world* w, float dx, int velExtraLayer) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "p2g: null world"); use_device(w); p2g(w, dx, velExtraLayer); sync(w); check_p2g_overflow(w); });
}
int flipb200_g2p_advect_sheetty(flipb200_world* w, float dt, float dx, int surfaceSize, int rkOrder, float picMin,
                                float picMax, int flags) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "g2p_advect: null world");
        use_device(w);
        g2p_advect_sheetty(w, dt, dx, surfaceSize, rkOrder, picMin, picMax, flags);
        sync(w);
    });
}
int flipb200_kill_particles_in_sdf(flipb200_world* w, int sdfGrid, int keep) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "kill_particles_in_sdf: null world");
        use_device(w);
        kill_particles_in_sdf(w, sdfGrid, keep != 0);
        sync(w);
    });
}
int flipb200_fluid_reseed(flipb200_world* w, uint32_t seed) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "fluid_reseed: null world");
        use_device(w);
        fluid_reseed(w, seed);
        sync(w);
    });
}
int flipb200_emit_liquid(flipb200_world* w, int shapeGrid, float vx, float vy, float vz, uint32_t seed) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "emit_liquid: null world");
        use_device(w);
        emit_liquid(w, shapeGrid, vx, vy, vz, seed);
        sync(w);
    });
}
int flipb200_apply_boundary(flipb200_world* w, int movingGrid, int movingVertexCentred) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "apply_boundary: null world");
        use_device(w);
        apply_boundary(w, movingGrid, movingVertexCentred != 0);
        sync(w);
    });
}
int flipb200_set_surface_tension(flipb200_world* w, float density, float tensionCoef) {
    return guarded([&] {
        FB_REQUIRE(w && density > 0.f, FLIPB200_ERR_ARG, "set_surface_tension: bad argument");
        // (slab decomposition: the curvature grid is read by voxel coordinate; the caller gives every rank the grid over its owned + ghost layers)
        w->density = density; w->tensionCoef = tensionCoef;
    });
}
int flipb200_particles_to_points(flipb200_world* w, float* pos, float* vel) {
    return guarded([&] {
        FB_REQUIRE(w && pos, FLIPB200_ERR_ARG, "particles_to_points: bad argument");
        use_device(w);
        particles_to_points(w, pos, vel);
    });
}
int flipb200_particles_add_dv(flipb200_world* w, float dvx, float dvy, float dvz) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "particles_add_dv: null world");
        use_device(w);
        particles_add_dv(w, dvx, dvy, dvz);
        sync(w);
    });
}
int flipb200_g2p_advect(flipb200_world* w, float dt, float dx, int rkOrder, float picSmoothness) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "g2p_advect: null world");
        use_device(w);
        g2p_advect_sheetty(w, dt, dx, /*surfaceSize=*/0, rkOrder, picSmoothness, 0.05f, /*flags: same field | plain*/ 3);
        sync(w);
    });
}
int flipb200_renormalize_sdf(flipb200_world* w, int grid, int iterations, int dilateIters) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "renormalize_sdf: null world");
        FB_REQUIRE(dilateIters == 0, FLIPB200_ERR_ARG, "renormalize_sdf: the tracker's dilate / erode (dilateIters != 0) is not accelerated");
        use_device(w);
        renormalize_sdf(w, grid, iterations);
        sync(w);
    });
}
int flipb200_erode_sdf(flipb200_world* w, int grid, float depth) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "erode_sdf: null world"); use_device(w); erode_sdf(w, grid, depth); sync(w); });
}
int flipb200_smooth_sdf(flipb200_world* w, int grid, int width, int iterations) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "smooth_sdf: null world"); use_device(w); smooth_sdf(w, grid, width, iterations); sync(w); });
}
int flipb200_dropped(flipb200_world* w, uint64_t* n) {
    return guarded([&] { FB_REQUIRE(w && n, FLIPB200_ERR_ARG, "dropped: bad argument"); *n = w->dropped; });
}
int flipb200_capture_precodec(flipb200_world* w, int on) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "capture_precodec: null world"); w->capturePreCodec = on != 0; });
}
int flipb200_get_precodec(flipb200_world* w, float* pos, float* vel, uint8_t* alive) {
    return guarded([&] {
        FB_REQUIRE(w && pos && vel && alive, FLIPB200_ERR_ARG, "get_precodec: bad argument");
        use_device(w);
        uint64_t n = w->preCodecN;
        if (!n) return;
        FB_CUDA(cudaMemcpyAsync(pos, w->preCodecPos.p, 12 * n, cudaMemcpyDeviceToHost, w->stream));
        FB_CUDA(cudaMemcpyAsync(vel, w->preCodecVel.p, 12 * n, cudaMemcpyDeviceToHost, w->stream));
        FB_CUDA(cudaMemcpyAsync(alive, w->preCodecAlive.p, n, cudaMemcpyDeviceToHost, w->stream));
        sync(w);
    });
}
int flipb200_face_weights(flipb200_world* w) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "face_weights: null world"); use_device(w); face_weights(w); sync(w); });
}
int flipb200_pushout_sdf(flipb200_world* w, float dx) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "pushout_sdf: null world"); use_device(w); pushout_sdf(w, dx); sync(w); });
}
int flipb200_add_vector(flipb200_world* w, float x, float y, float z) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "add_vector: null world"); use_device(w); add_vector(w, x, y, z); sync(w); });
}
int flipb200_cfl(flipb200_world* w, float* dtOut) {
    return guarded([&] { FB_REQUIRE(w && dtOut, FLIPB200_ERR_ARG, "cfl: bad argument"); use_device(w); *dtOut = cfl(w); });
}
int flipb200_solve_ppe_ex(flipb200_world* w, float dt, float dx, float relTol, int maxIter, int* iterations,
                          float* relResidual, int* status) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "solve_ppe: null world");
        use_device(w);
        solve_ppe(w, dt, dx, relTol, maxIter);
        if (iterations) *iterations = w->solver.iterations;
        if (relResidual) *relResidual = w->solver.relResidual;
        if (status) *status = w->solver.status;
    });
}
int flipb200_solve_ppe(flipb200_world* w, float dt, float dx, int* iterations, float* relResidual, int* status) {
    // mRelativeTolerance = 5e-5, mMaxIteration = 100 (FF/FLIP_vdb.cpp:3052, uaamg.h:164)
    return flipb200_solve_ppe_ex(w, dt, dx, 5e-5f, 100, iterations, relResidual, status);
}
int flipb200_solver_info(flipb200_world* w, int* levels, int* numDof, int* nHistory) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "solver_info: null world");
        if (levels) *levels = w->solver.levels;
        if (numDof) *numDof = w->solver.numDof;
        if (nHistory) *nHistory = (int)w->solver.history.size();
    });
}
int flipb200_residual_history(flipb200_world* w, float* out) {
    return guarded([&] {
        FB_REQUIRE(w && out, FLIPB200_ERR_ARG, "residual_history: bad argument");
        std::memcpy(out, w->solver.history.data(), sizeof(float) * w->solver.history.size());
    });
}
int flipb200_subtract_grad(flipb200_world* w, float dt, float dx, int velExtraLayer) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "subtract_grad: null world"); use_device(w); subtract_grad(w, dt, dx, velExtraLayer); sync(w); });
}

int flipb200_substep(flipb200_world* w, float dt, float dx, int surfaceSize, int rkOrder, float picMin, float picMax,
                     float gx, float gy, float gz, int velExtraLayer, int flags, float* stageMs) {
    return guarded([&] {
        FB_REQUIRE(w, FLIPB200_ERR_ARG, "substep: null world");
        use_device(w);
        cudaEvent_t ev[6];
        if (stageMs) for (auto& e : ev) FB_CUDA(cudaEventCreate(&e));
        auto mark = [&](int i) { if (stageMs) FB_CUDA(cudaEventRecord(ev[i], w->stream)); };
        mark(0);
        { FB_PHASE(w, "S1 g2p_advect_rebin"); g2p_advect_sheetty(w, dt, dx, surfaceSize, rkOrder, picMin, picMax, flags); }
        mark(1);
        { FB_PHASE(w, "S2 p2g"); p2g(w, dx, velExtraLayer); }
        mark(2);
        { FB_PHASE(w, "S3 stencils");
          face_weights(w);
          pushout_sdf(w, dx);
          add_vector(w, gx * dt, gy * dt, gz * dt); }
        mark(3);
        { FB_PHASE(w, "S4 solve_ppe"); solve_ppe(w, dt, dx, 5e-5f, 100); }
        mark(4);
        { FB_PHASE(w, "S5 subtract_grad"); subtract_grad(w, dt, dx, velExtraLayer); }
        mark(5);
        sync(w);
        check_p2g_overflow(w);
        if (stageMs) {
            for (int i = 0; i < 5; i++) FB_CUDA(cudaEventElapsedTime(&stageMs[i], ev[i], ev[i + 1]));
            for (auto& e : ev) cudaEventDestroy(e);
        }
    });
}

int flipb200_launch_count(flipb200_world* w, uint64_t* n) {
    return guarded([&] { FB_REQUIRE(w && n, FLIPB200_ERR_ARG, "launch_count: bad argument"); *n = w->launches; });
}
int flipb200_sync_count(flipb200_world* w, uint64_t* n) {
    return guarded([&] { FB_REQUIRE(w && n, FLIPB200_ERR_ARG, "sync_count: bad argument"); *n = w->syncs; });
}
int flipb200_profile_enable(flipb200_world* w, int on) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "profile_enable: null world"); use_device(w); resolve_profile(w); w->profiling = on != 0; });
}
int flipb200_profile_reset(flipb200_world* w) {
    return guarded([&] { FB_REQUIRE(w, FLIPB200_ERR_ARG, "profile_reset: null world"); use_device(w); resolve_profile(w); w->prof.clear(); });
}
int flipb200_profile_get(flipb200_world* w, char* names, size_t namesCap, float* ms, uint64_t* launches, uint64_t* bytes,
                         int cap, int* nOut) {
    return guarded([&] {
        FB_REQUIRE(w && names && nOut, FLIPB200_ERR_ARG, "profile_get: bad argument");
        use_device(w);
        resolve_profile(w);
        std::string s;
        int i = 0;
        for (auto& kv : w->prof) {
            if (i >= cap) break;
            if (!s.empty()) s += ";";
            s += kv.first;
            if (ms) ms[i] = (float)kv.second.ms;
            if (launches) launches[i] = kv.second.launches;
            if (bytes) bytes[i] = kv.second.bytes;
            i++;
        }
        FB_REQUIRE(s.size() + 1 <= namesCap, FLIPB200_ERR_ARG, "profile_get: name buffer too small");
        std::memcpy(names, s.c_str(), s.size() + 1);
        *nOut = i;
    });
}
int flipb200_stream(flipb200_world* w, void** stream) {
    return guarded([&] { FB_REQUIRE(w && stream, FLIPB200_ERR_ARG, "stream: bad argument"); *stream = (void*)w->stream; });
}

}  // extern "C"
