// libflipb200 -- world state + internal API shared between the .cu files.
#pragma once
#include "common.cuh"
#include "../../include/flipb200.h"
#include <ctime>
#include <cstring>
#include <cstdlib>

namespace fb {

struct ProfEntry { double ms = 0; uint64_t launches = 0; uint64_t bytes = 0; };
struct PendingEvt { cudaEvent_t a, b; std::string name; uint64_t bytes; };

struct SolverStats {
    int iterations = 0, status = 0, levels = 0, numDof = 0;
    float relResidual = 0.f;
    std::vector<float> history;
};

struct Comm;     // NCCL / in-process communicator (comm.cu)
struct DDState;  // slab decomposition state (dd.cu)

}  // namespace fb

struct flipb200_world {
    int device = 0;
    float dx = 0.f;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    uint64_t syncs = 0;      // host waits on the stream (each one is a bubble on the device)
    bool profiling = false;
    std::map<std::string, fb::ProfEntry> prof;
    std::map<std::string, fb::ProfEntry> phase;   // FLIPB200_PHASE_TRACE
    std::vector<fb::PendingEvt> pending;
    std::vector<cudaEvent_t> evtPool;

    // the fluid-band pool: particle leaves + one ring of leaves; all band grids live on it
    fb::TopoPtr pool;
    uint64_t epochCounter = 0;

    fb::GridV vgrid[5];   // ids 0..4
    fb::GridF fgrid[FLIPB200_NUM_GRIDS];  // ids 5..10 used
    bool hasSolidSDF = false, hasSolidVel = false;
    fb::Particles pts;

    // views of the static solid grids resampled onto the pool (rebuilt when the pool changes)
    uint64_t solidViewEpoch = ~0ull;
    float density = 1000.f, tensionCoef = 0.f;   // the Density / SurfaceTension sockets (flipb200_set_surface_tension); coef > 0 enables the tension terms
    fb::DBuf<float> solidSdfView;        // [pool n][512]
    fb::DBuf<float> solidVelView[3];     // [pool n][512]
    fb::DBuf<uint8_t> solidLeafExists;   // [pool n]

    // FLIP_P2G's staging-overflow flag: written by the kernel, copied to page-locked host memory in stream order and
    // checked at the next host wait that happens anyway (flipb200_p2g / flipb200_substep), not with a wait of its own
    fb::DBuf<int> p2gOverflow;
    fb::DBuf<float> ddCoarse[4];   // slab decomposition: the global level-1 coefficient arrays, kept between solves (no pool traffic for 4 x nv floats per solve)
    fb::DBuf<float> ddStage;   // slab decomposition: persistent operand of the coarse-level all-reduce (NCCL sees the same buffer every solve)
    int* p2gOverflowHost = nullptr;

    // page-locked scratch for the few-byte read-backs (a pageable destination makes the driver stage the copy)
    unsigned char* hostScratch = nullptr;   // 1 KB, pinned + mapped, allocated by flipb200_world_create
    unsigned char* hostScratchDev = nullptr;   // its device alias (d2h_words)

    uint64_t reservedBytes = 0;   // device memory parked in the stream-ordered pool (reserve_pool)

    uint64_t dropped = 0;
    bool capturePreCodec = false;
    fb::DBuf<float> preCodecPos, preCodecVel;
    fb::DBuf<uint8_t> preCodecAlive;
    uint64_t preCodecN = 0;

    fb::SolverStats solver;
    // asynchronous downloads (flipb200_*_download_begin / flipb200_download_wait): device->host copies run on a
    // second stream behind an event, so they overlap the nodes that follow; staging buffers are held until the wait
    cudaStream_t copyStream = nullptr;
    cudaEvent_t copyEvt = nullptr;
    std::vector<std::shared_ptr<void>> held;

    fb::Comm* comm = nullptr;
    int rank = 0, nRanks = 1;
    fb::DDState* dd = nullptr;   // non-null once flipb200_dd_set_slab was called

    fb::GridV& V(int id) { return vgrid[id]; }
    fb::GridF& F(int id) { return fgrid[id]; }
};

namespace fb {

using World = flipb200_world;

inline bool is_vec_grid(int id) { return id >= 0 && id <= FLIPB200_FACE_WEIGHT; }
inline bool is_float_grid(int id) { return id >= FLIPB200_LIQUID_SDF && id < FLIPB200_NUM_GRIDS; }
inline bool is_static_grid(int id) { return id == FLIPB200_SOLID_SDF || id == FLIPB200_SOLID_VELOCITY; }

// ---- launch accounting -------------------------------------------------------------
struct LaunchScope {
    World* w; const char* name; uint64_t bytes; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(World* w_, const char* name_, uint64_t bytes_) : w(w_), name(name_), bytes(bytes_) {
        w->launches++;
        if (w->profiling) {
            auto get = [&]() {
                if (!w->evtPool.empty()) { cudaEvent_t e = w->evtPool.back(); w->evtPool.pop_back(); return e; }
                cudaEvent_t e; cudaEventCreate(&e); return e;
            };
            a = get(); b = get();
            cudaEventRecord(a, w->stream);
        }
    }
    ~LaunchScope() {
        if (a) {
            cudaEventRecord(b, w->stream);
            w->pending.push_back(PendingEvt{a, b, name, bytes});
        }
    }
};
// usage: FB_LAUNCH(w, "p2g_gather", bytes) kernel<<<grid, block, smem, w->stream>>>(...);
#define FB_LAUNCH(w, name, bytes) if (fb::LaunchScope ls_{(w), (name), (uint64_t)(bytes)}; true)

inline void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(2, std::string(what) + ": " + cudaGetErrorString(e));
}
// Read a few bytes back and wait for them (the data-dependent sizes of the path: leaf counts, totals, the PCG norm).
// part k of a multi-part read-back goes to scratch offset `at`; the last part waits and the caller copies out.
// The words travel by a one-warp kernel that stores through the device alias of the (mapped, pinned) scratch block, not by a
// DMA copy: a cudaMemcpyAsync of 4 bytes queues on the device-to-host copy engine BEHIND whatever bulk download another
// stream has in flight (measured: FLIP_P2G 2.2 -> 3.7 ms while a 200 MB particle download was running, its five read-backs
// each waiting for the engine), and costs ~10 us of DMA set-up even on an idle engine.
static __global__ void readback_words_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
// stream-ordered, no wait: `scratchDst` lies inside w->hostScratch; src and bytes are multiples of 4
inline void d2h_words(World* w, void* scratchDst, const void* src, size_t bytes) {
    const size_t at = (size_t)((unsigned char*)scratchDst - w->hostScratch);
    if (at + bytes > 1024 || (bytes & 3) || !w->hostScratchDev) {   // not ours: plain copy
        FB_CUDA(cudaMemcpyAsync(scratchDst, src, bytes, cudaMemcpyDeviceToHost, w->stream));
        return;
    }
    readback_words_kernel<<<1, 32, 0, w->stream>>>(reinterpret_cast<uint32_t*>(w->hostScratchDev + at), reinterpret_cast<const uint32_t*>(src), (int)(bytes >> 2));
}
inline void read_back(World* w, void* dst, const void* src, size_t bytes) {
    d2h_words(w, w->hostScratch, src, bytes);
    w->syncs++;
    FB_CUDA(cudaStreamSynchronize(w->stream));
    memcpy(dst, w->hostScratch, bytes);
}
// Park `bytes` of device memory in the stream-ordered pool (allocate once, free: the release threshold keeps it), so
// that the per-substep temporaries -- whose sizes creep up as the fluid spreads -- never make the pool map new physical
// memory in the middle of a step (measured: occasional 20-500 ms substeps at 16.8 M particles without it).
inline void reserve_pool(World* w, uint64_t bytes) {
    if (bytes <= w->reservedBytes) return;
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess && bytes > freeB / 2) bytes = freeB / 2;
    if (bytes <= w->reservedBytes) return;
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes, w->stream) == cudaSuccess) { cudaFreeAsync(p, w->stream); w->reservedBytes = bytes; }
    else cudaGetLastError();
}
void set_last_error(const std::string& m);   // abi.cu: what flipb200_last_error() returns on this thread
inline void sync(World* w) {
    w->syncs++;
    FB_CUDA(cudaStreamSynchronize(w->stream));
    // a ghost exchange over peer memory whose neighbour never pushed (dd.cu: dd_wait_kernel's time limit) reports here
    int* err = reinterpret_cast<int*>(w->hostScratch + 800);
    if (w->hostScratch && *err) { *err = 0; throw Error(FLIPB200_ERR_COMM, "ghost exchange: a neighbour's push did not arrive within the time limit"); }
}

// ---- host-phase trace (FLIPB200_PHASE_TRACE=1): wall clock of named host phases, stream-synchronised on both
// sides, accumulated per world and printed when the world is destroyed. A diagnosis aid, off by default.
struct PhaseTimer {
    World* w; const char* name; double t0 = 0; bool on;
    static bool enabled() { static const bool e = getenv("FLIPB200_PHASE_TRACE") != nullptr; return e; }
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    PhaseTimer(World* w_, const char* n) : w(w_), name(n), on(enabled()) { if (on) { cudaStreamSynchronize(w->stream); t0 = now(); } }
    ~PhaseTimer() { if (on) { cudaStreamSynchronize(w->stream); auto& e = w->phase[name]; e.ms += now() - t0; e.launches++; } }
};
#define FB_PHASE_CAT2(a, b) a##b
#define FB_PHASE_CAT(a, b) FB_PHASE_CAT2(a, b)
#define FB_PHASE(w, name) fb::PhaseTimer FB_PHASE_CAT(phase_, __LINE__)((w), (name))

// ---- topo.cu ------------------------------------------------------------------------
// Build a topology from candidate leaf-origin voxel coordinates on the device (duplicates
// allowed). ring: also add the 26 neighbour leaves of every candidate.
// bbox (optional): leaf-coordinate bounds {xmin,ymin,zmin,xmax,ymax,zmax} known to contain every candidate; saves the
// bounding-box kernel and its host read-back (the directory may then be larger than tight, slot order is unaffected).
TopoPtr topo_from_origins_dev(World* w, const int3* origins_dev, int count, bool ring, const int* bbox = nullptr);
TopoPtr topo_from_origins_host(World* w, const int32_t* origins, int count, bool ring);
void grid_alloc(World* w, GridF& g, const TopoPtr& t, float bg);              // values = bg, mask = 0
void grid_alloc(World* w, GridV& g, const TopoPtr& t, const float bg[3]);
void grid_rebase(World* w, GridF& g, const TopoPtr& t);                        // move onto t (lossless if t covers g)
void grid_rebase(World* w, GridV& g, const TopoPtr& t);
void grid_copy(World* w, GridF& dst, const GridF& src);
void grid_copy(World* w, GridV& dst, const GridV& src);
// makes sure w->pool exists, covers the particle leaves (+ring) and the listed band grids, and
// that those grids live on it.
void ensure_pool(World* w, std::initializer_list<int> gridIds, bool includeParticles);
void refresh_solid_views(World* w);
// mask morphology on the pool: out = dilate(in) (26- or 6-neighbourhood), in/out may not alias
void mask_dilate(World* w, const Topo& t, const uint64_t* in, uint64_t* out, bool nn26);
uint64_t mask_count(World* w, const uint64_t* mask, int nLeaves);
// asynchronous form: adds the popcount to *out (device, zeroed by the caller); no host wait
void mask_count_async(World* w, const uint64_t* mask, int nLeaves, unsigned long long* out);

// ---- scan.cu ------------------------------------------------------------------------
// exclusive prefix sum of n uint32 (out may alias in); total (if non-null) receives the sum on the host
void exclusive_scan_u32(World* w, const uint32_t* in, uint32_t* out, size_t n, uint64_t* total);

// ---- particles.cu -------------------------------------------------------------------
void bin_from_points(World* w, const float* pos_host, const float* vel_host, uint64_t n);
// re-bin after advect: keys = target (pool slot*512 + off) or 0xffffffff for dropped particles
// KillParticlesInSDF (FF/nosys/KillParticles.cpp): drop the particles the killer SDF (float grid sdfGrid) rejects
void kill_particles_in_sdf(World* w, int sdfGrid, bool keep);
// ParticleAddDV (FF/nosys/ParticleAddGravity.cpp): v += dv on every stored particle, through the half codec
void particles_add_dv(World* w, float dvx, float dvy, float dvz);
void rebin_particles(World* w, const TopoPtr& newPool, const uint32_t* keys_dev, uint64_t nOld,
                     DBuf<uint32_t>& w0, DBuf<uint32_t>& w1, DBuf<uint32_t>& w2);

// ---- p2g.cu / g2p.cu / stencils.cu / poisson.cu --------------------------------------
void p2g(World* w, float dx, int velExtraLayer);
void check_p2g_overflow(World* w);
void g2p_advect_sheetty(World* w, float dt, float dx, int surfaceSize, int rkOrder, float picMin,
                        float picMax, int flags);
void face_weights(World* w);
void pushout_sdf(World* w, float dx);
void add_vector(World* w, float x, float y, float z);
float cfl(World* w);
void subtract_grad(World* w, float dt, float dx, int velExtraLayer);
// VDBRenormalizeSDF (projects/zenvdb/VDBRenormalize.cpp): LevelSetTracker::normalize x iterations on a float grid
void renormalize_sdf(World* w, int grid, int iterations);
void erode_sdf(World* w, int grid, float depth);
void smooth_sdf(World* w, int grid, int width, int iterations);   // VDBSmoothSDF: openvdb Filter::gaussian   // VDBErodeSDF: active voxels += depth
// per-channel masks are [3][n][8]; target topology mask = liquid SDF mask
void union_extrapolate(World* w, int nLayer, GridV& vel, uint64_t* chMask, const uint64_t* targetMask);
void solve_ppe(World* w, float dt, float dx, float relTol, int maxIter);

// ---- comm.cu ------------------------------------------------------------------------
// Point-to-point messages are issued in groups (ncclGroupStart/End semantics): every send/recv between
// begin and end is matched with the peer's in issue order; buffers are device memory, sizes in bytes.
enum { CT_F32 = 0, CT_U32 = 1, CT_I32 = 2 };
void comm_destroy(World* w);
bool comm_active(World* w);
bool comm_is_local(World* w);   // the in-process backend (several worlds of one process, same address space)
void comm_group_begin(World* w);
void comm_send(World* w, int peer, const void* buf, size_t bytes);
void comm_recv(World* w, int peer, void* buf, size_t bytes);
void comm_group_end(World* w);
void comm_allreduce(World* w, void* buf, size_t n, int type, bool isMax);   // in place, sum or max

// ---- dd.cu --------------------------------------------------------------------------
// Slab decomposition along x: rank r owns the leaf layers [lo, hi) (leaf coordinate = voxel >> 3),
// keeps one ghost layer of particles on each side and a pool that reaches one further ring layer.
struct DDArray { void* base; int bytesPerLeaf; };
bool dd_on(World* w);
void sort_store_by_key(World* w, const TopoPtr& topo, const uint32_t* keys, uint64_t n, const uint32_t* i0, const uint32_t* i1, const uint32_t* i2, Particles& out);   // particles.cu
void particles_to_points(World* w, float* posHost, float* velHost);   // particles.cu
void fluid_reseed(World* w, uint32_t seed);   // reseed.cu
void apply_boundary(World* w, int movingGrid, bool movingVertexCentred);   // reseed.cu
void emit_liquid(World* w, int shapeGrid, float vx, float vy, float vz, uint32_t seed);   // reseed.cu
void dd_destroy(World* w);
void dd_set_slab(World* w, int lo, int hi);
void dd_owned_slots(World* w, int* ownLo, int* ownHi);          // slot range of the owned leaves in w->pool (collective)
void dd_owned_coords(World* w, int* lo, int* hi);               // owned leaf-layer range; +-2^29 at the open ends
void dd_refresh(World* w, const std::vector<DDArray>& arrays, int layers);   // ghost (+ring) leaves <- owner (collective)
void dd_refresh(World* w, GridF& g, int layers);
void dd_refresh(World* w, GridV& g, int layers);
// exchange after a move: particles [pLo,pHi) of the arrays (alive may be null = all alive) are kept, sent to the left /
// right neighbour (migrants + ghost copies), and the merged set [from left | kept | from right] is returned.
void dd_migrate(World* w, uint64_t pLo, uint64_t pHi, const uint32_t* w0, const uint32_t* w1, const uint32_t* w2, const int3* ijk,
                const uint8_t* alive, DBuf<uint32_t>& o0, DBuf<uint32_t>& o1, DBuf<uint32_t>& o2, DBuf<int3>& oijk, uint64_t* nOut);

}  // namespace fb
