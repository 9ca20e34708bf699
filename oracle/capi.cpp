// TEST INFRASTRUCTURE ONLY -- see flip_oracle.h.
// Flat C entry points so tests/ and bench.py can drive the oracle through ctypes.
// The argument layout mirrors include/flipb200.h so one Python harness drives both.
#include <random>
#include "flip_oracle.h"
#include <chrono>
#include <cstdio>

using namespace orc;

namespace {
enum GridId { G_VELOCITY = 0, G_POSTADV = 1, G_VISCOUS = 2, G_SOLIDVEL = 3, G_FACEWEIGHT = 4,
              G_LIQUIDSDF = 5, G_SOLIDSDF = 6, G_PRESSURE = 7, G_DIVERGENCE = 8, G_CURVATURE = 9, G_KILLERSDF = 10 };
Vec3Grid* vec3Of(World* w, int id) {
    switch (id) {
    case G_VELOCITY: return &w->velocity;
    case G_POSTADV: return &w->postAdvVelocity;
    case G_VISCOUS: return &w->viscousVelocity;
    case G_SOLIDVEL: return &w->solidVelocity;
    case G_FACEWEIGHT: return &w->faceWeight;
    }
    return nullptr;
}
FloatGrid* floatOf(World* w, int id) {
    switch (id) {
    case G_LIQUIDSDF: return &w->liquidSDF;
    case G_SOLIDSDF: return &w->solidSDF;
    case G_PRESSURE: return &w->pressure;
    case G_DIVERGENCE: return &w->divergence;
    case G_CURVATURE: return &w->curvature;
    case G_KILLERSDF: return &w->killerSDF;
    }
    return nullptr;
}
template <int NC>
void setGrid(Grid<NC>& g, int n, const int32_t* origins, const uint64_t* masks, const float* values, const float* bg) {
    g.clear();
    for (int c = 0; c < NC; c++) g.bg[c] = bg[c];
    for (int l = 0; l < n; l++) {
        int id = g.touchLeaf(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2]);
        for (int k = 0; k < 8; k++) g.masks[id][k] = masks[8 * size_t(l) + k];
        std::memcpy(g.leafVals(id, 0), values + size_t(l) * NC * 512, sizeof(float) * NC * 512);
    }
}
template <int NC>
void getGrid(const Grid<NC>& g, int32_t* origins, uint64_t* masks, float* values, float* bg) {
    for (int c = 0; c < NC; c++) bg[c] = g.bg[c];
    for (int l = 0; l < g.leafCount(); l++) {
        origins[3 * l] = g.origins[l].x; origins[3 * l + 1] = g.origins[l].y; origins[3 * l + 2] = g.origins[l].z;
        for (int k = 0; k < 8; k++) masks[8 * size_t(l) + k] = g.masks[l][k];
        std::memcpy(values + size_t(l) * NC * 512, g.leafVals(l, 0), sizeof(float) * NC * 512);
    }
}
}  // namespace

extern "C" {

void* orc_world_create(float dx) { return new World(dx); }
void orc_world_destroy(void* w) { delete static_cast<World*>(w); }

int orc_grid_channels(int id) { return id <= G_FACEWEIGHT ? 3 : 1; }

// values: [leaf][channel][512] fp32; masks: [leaf][8] u64 (word = x, bit = y*8+z); origins: [leaf][3] i32
int orc_grid_set(void* wp, int id, int nLeaves, const int32_t* origins, const uint64_t* masks,
                 const float* values, const float* bg) {
    World* w = static_cast<World*>(wp);
    if (Vec3Grid* g = vec3Of(w, id)) setGrid(*g, nLeaves, origins, masks, values, bg);
    else if (FloatGrid* f = floatOf(w, id)) setGrid(*f, nLeaves, origins, masks, values, bg);
    else return 1;
    if (id == G_SOLIDSDF) w->hasSolidSDF = true;
    if (id == G_SOLIDVEL) w->hasSolidVel = true;
    return 0;
}
int orc_grid_leaf_count(void* wp, int id) {
    World* w = static_cast<World*>(wp);
    if (Vec3Grid* g = vec3Of(w, id)) return g->leafCount();
    if (FloatGrid* f = floatOf(w, id)) return f->leafCount();
    return -1;
}
int orc_grid_get(void* wp, int id, int32_t* origins, uint64_t* masks, float* values, float* bg) {
    World* w = static_cast<World*>(wp);
    if (Vec3Grid* g = vec3Of(w, id)) getGrid(*g, origins, masks, values, bg);
    else if (FloatGrid* f = floatOf(w, id)) getGrid(*f, origins, masks, values, bg);
    else return 1;
    return 0;
}

// particle store in the reference layout
int orc_particles_set(void* wp, int nLeaves, const int32_t* origins, const uint32_t* voxelEnd, uint64_t n,
                      const uint16_t* P, const uint16_t* v) {
    World* w = static_cast<World*>(wp);
    Points& p = w->particles;
    p.clear();
    uint64_t begin = 0;
    for (int l = 0; l < nLeaves; l++) {
        p.dir.emplace(leafKeyOf(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2]), l);
        p.origins.push_back(Coord(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2]));
        std::array<uint32_t, 512> ve;
        std::memcpy(ve.data(), voxelEnd + size_t(l) * 512, sizeof(uint32_t) * 512);
        p.voxelEnd.push_back(ve);
        p.leafBegin.push_back(begin);
        begin += ve[511];
    }
    p.leafBegin.push_back(begin);
    if (begin != n) return 2;
    p.P.assign(P, P + 3 * n);
    p.v.assign(v, v + 3 * n);
    return 0;
}
int orc_particles_info(void* wp, int* nLeaves, uint64_t* n) {
    World* w = static_cast<World*>(wp);
    *nLeaves = w->particles.leafCount();
    *n = w->particles.size();
    return 0;
}
int orc_particles_get(void* wp, int32_t* origins, uint32_t* voxelEnd, uint16_t* P, uint16_t* v) {
    World* w = static_cast<World*>(wp);
    const Points& p = w->particles;
    for (int l = 0; l < p.leafCount(); l++) {
        origins[3 * l] = p.origins[l].x; origins[3 * l + 1] = p.origins[l].y; origins[3 * l + 2] = p.origins[l].z;
        std::memcpy(voxelEnd + size_t(l) * 512, p.voxelEnd[l].data(), sizeof(uint32_t) * 512);
    }
    std::memcpy(P, p.P.data(), sizeof(uint16_t) * p.P.size());
    std::memcpy(v, p.v.data(), sizeof(uint16_t) * p.v.size());
    return 0;
}

int orc_bin_from_points(void* wp, const float* pos, const float* vel, uint64_t n) {
    bin_from_points(*static_cast<World*>(wp), pos, vel, n);
    return 0;
}
int orc_p2g(void* wp, float dx, int velExtraLayer) {
    node_FLIP_P2G(*static_cast<World*>(wp), dx, velExtraLayer);
    return 0;
}
// flags bit0: ViscousVelocity socket is wired to the Velocity object (same grid)
int orc_g2p_advect_sheetty(void* wp, float dt, float dx, int surfaceSize, int rkOrder, float picMin,
                           float picMax, int flags) {
    World* w = static_cast<World*>(wp);
    if (flags & 1) w->viscousVelocity = w->velocity;
    node_G2PAdvectorSheetty(*w, dt, dx, surfaceSize, rkOrder, picMin, picMax);
    return 0;
}
int orc_kill_particles(void* wp, int sdfGrid, int keep) {
    World* w = static_cast<World*>(wp);
    FloatGrid* g = floatOf(w, sdfGrid);
    if (!g) return 1;
    node_KillParticlesInSDF(*w, *g, keep != 0);
    return 0;
}
// leafStart / leafEnd: nullable, one entry per particle leaf (see node_FluidReseed)
int orc_fluid_reseed(void* wp, uint32_t seed, const uint64_t* leafStart, uint64_t* leafEnd) {
    node_FluidReseed(*static_cast<World*>(wp), seed, leafStart, leafEnd);
    return 0;
}
int orc_emit_liquid(void* wp, int shapeGrid, float vx, float vy, float vz, uint32_t seed, const uint64_t* leafStart, uint64_t* leafEnd) {
    World* w = static_cast<World*>(wp);
    FloatGrid* g = floatOf(w, shapeGrid);
    if (!g) return 1;
    node_ParticleEmitter(*w, *g, vx, vy, vz, seed, leafStart, leafEnd);
    return 0;
}
int orc_apply_boundary(void* wp, int movingGrid, int movingVertexCentred) {
    World* w = static_cast<World*>(wp);
    FloatGrid* g = floatOf(w, movingGrid);
    if (!g) return 1;
    node_FLIPApplyBoundary(*w, *g, movingVertexCentred != 0);
    return 0;
}
int orc_set_surface_tension(void* wp, float density, float coef) {
    World* w = static_cast<World*>(wp);
    w->density = density; w->tensionCoef = coef;
    return 0;
}
int orc_particles_to_points(void* wp, float* pos, float* vel) { node_VDBPointsToPrimitive(*static_cast<World*>(wp), pos, vel); return 0; }
uint64_t orc_reseed_leaf_start(uint32_t seed, int ox, int oy, int oz) { return reseed_leaf_start(seed, ox, oy, oz); }
// the start the seeded build of the reference draws for a TBB chunk: uniform_int_distribution(0, 21474836)(mt19937(seed)), FF/FLIP_vdb.cpp:2081-2084
uint64_t orc_reseed_chunk_start(uint32_t seed) {
    std::uniform_int_distribution<> intdistrib(0, 21474836);
    std::mt19937 gen(seed);
    return (uint64_t)(size_t)intdistrib(gen);
}
int orc_particles_add_dv(void* wp, float x, float y, float z) { node_ParticleAddDV(*static_cast<World*>(wp), x, y, z); return 0; }
int orc_g2p_advect(void* wp, float dt, float dx, int rkOrder, float picSmoothness) { node_G2P_Advector(*static_cast<World*>(wp), dt, dx, rkOrder, picSmoothness); return 0; }
int orc_renormalize_sdf(void* wp, int grid, int iterations, int dilateIters) {
    World* w = static_cast<World*>(wp);
    FloatGrid* g = floatOf(w, grid);
    if (!g || dilateIters != 0) return 1;
    node_VDBRenormalizeSDF(*g, w->dx, iterations);
    return 0;
}
int orc_erode_sdf(void* wp, int grid, float depth) {
    World* w = static_cast<World*>(wp);
    FloatGrid* g = floatOf(w, grid);
    if (!g) return 1;
    node_VDBErodeSDF(*g, depth);
    return 0;
}
int orc_smooth_sdf(void* wp, int grid, int width, int iterations) {
    World* w = static_cast<World*>(wp);
    FloatGrid* g = floatOf(w, grid);
    if (!g) return 1;
    node_VDBSmoothSDF(*g, width, iterations);
    return 0;
}
int orc_face_weights(void* wp) { node_CutCellWeight(*static_cast<World*>(wp)); return 0; }
int orc_pushout_sdf(void* wp, float dx) { node_PushOutLiquidSDF(*static_cast<World*>(wp), dx); return 0; }
int orc_add_vector(void* wp, float x, float y, float z) { node_FieldAddVector(*static_cast<World*>(wp), x, y, z); return 0; }
float orc_cfl(void* wp) { return node_CFL_dt(*static_cast<World*>(wp)); }
int orc_solve_ppe(void* wp, float dt, float dx, int* iters, float* relResidual, int* status) {
    World* w = static_cast<World*>(wp);
    node_AssembleSolvePPE(*w, dt, dx);
    if (iters) *iters = w->pcgIterations;
    if (relResidual) *relResidual = w->pcgRelResidual;
    if (status) *status = w->pcgStatus;
    return 0;
}
// test hook: the same node with the solver's tolerance / iteration cap overridden (libflipb200's flipb200_solve_ppe_ex)
int orc_solve_ppe_ex(void* wp, float dt, float dx, float relTol, int maxIter, int* iters, float* relResidual, int* status) {
    World* w = static_cast<World*>(wp);
    const float t0 = w->solveRelTol; const int m0 = w->solveMaxIter;
    w->solveRelTol = relTol; w->solveMaxIter = maxIter;
    node_AssembleSolvePPE(*w, dt, dx);
    w->solveRelTol = t0; w->solveMaxIter = m0;
    if (iters) *iters = w->pcgIterations;
    if (relResidual) *relResidual = w->pcgRelResidual;
    if (status) *status = w->pcgStatus;
    return 0;
}
int orc_solver_info(void* wp, int* levels, int* numDof, int* nHistory) {
    World* w = static_cast<World*>(wp);
    *levels = w->mgLevels; *numDof = w->numDof; *nHistory = int(w->residualHistory.size());
    return 0;
}
int orc_residual_history(void* wp, float* out) {
    World* w = static_cast<World*>(wp);
    std::memcpy(out, w->residualHistory.data(), sizeof(float) * w->residualHistory.size());
    return 0;
}
int orc_subtract_grad(void* wp, float dt, float dx, int velExtraLayer) {
    node_SubtractPressureGradient(*static_cast<World*>(wp), dt, dx, velExtraLayer);
    return 0;
}
uint64_t orc_dropped(void* wp) { return static_cast<World*>(wp)->droppedParticles; }

// pre-codec capture for the G2P tolerance gate
int orc_capture_precodec(void* wp, int on) { static_cast<World*>(wp)->capturePreCodec = on != 0; return 0; }
int orc_get_precodec(void* wp, float* pos, float* vel, uint8_t* alive) {
    World* w = static_cast<World*>(wp);
    std::memcpy(pos, w->preCodecPos.data(), sizeof(float) * w->preCodecPos.size());
    std::memcpy(vel, w->preCodecVel.data(), sizeof(float) * w->preCodecVel.size());
    std::memcpy(alive, w->preCodecAlive.data(), w->preCodecAlive.size());
    return 0;
}

// codec tables for the golden tests
uint16_t orc_fxpt16_encode(float p) { return fxpt16_encode(p); }
float orc_fxpt16_decode(uint16_t u) { return fxpt16_decode(u); }
uint16_t orc_half_encode(float f) { return half_encode(f); }
float orc_half_decode(uint16_t h) { return half_decode(h); }
float orc_fraction_inside2(float a, float b) { return fraction_inside(a, b); }
float orc_fraction_inside4(float bl, float br, float tl, float tr) { return fraction_inside(bl, br, tl, tr); }

// one full substep of the test chain (last paragraph); returns per-stage seconds
int orc_substep(void* wp, float dt, float dx, int surfaceSize, int rkOrder, float picMin, float picMax,
                float gx, float gy, float gz, int velExtraLayer, int flags, double* stageSeconds) {
    World* w = static_cast<World*>(wp);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    auto t0 = now();
    if (flags & 1) w->viscousVelocity = w->velocity;
    node_G2PAdvectorSheetty(*w, dt, dx, surfaceSize, rkOrder, picMin, picMax);
    auto t1 = now();
    node_FLIP_P2G(*w, dx, velExtraLayer);
    auto t2 = now();
    node_CutCellWeight(*w);
    node_PushOutLiquidSDF(*w, dx);
    node_FieldAddVector(*w, gx * dt, gy * dt, gz * dt);
    auto t3 = now();
    node_AssembleSolvePPE(*w, dt, dx);
    auto t4 = now();
    node_SubtractPressureGradient(*w, dt, dx, velExtraLayer);
    auto t5 = now();
    if (stageSeconds) {
        stageSeconds[0] = secs(t0, t1); stageSeconds[1] = secs(t1, t2); stageSeconds[2] = secs(t2, t3);
        stageSeconds[3] = secs(t3, t4); stageSeconds[4] = secs(t4, t5);
    }
    return 0;
}

}  // extern "C"
