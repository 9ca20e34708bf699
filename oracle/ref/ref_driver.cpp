// TEST INFRASTRUCTURE ONLY -- flat C driver over the REAL reference FastFLIP CPU path.
//
// This file is ours; everything it calls is the reference's own, unmodified code compiled from
// /root/reference by oracle/ref/build_ref.sh: FLIP_vdb::* (projects/FastFLIP/FLIP_vdb.cpp),
// simd_uaamg::* (simd_vdb_poisson_uaamg.cpp), vdb_velocity_extrapolator::union_extrapolate,
// packed_FloatGrid3 (projects/zenvdb/include/zeno/packed3grids.cpp) and OpenVDB 9.0.1.
// Each ref_* entry point performs exactly the calls of the node shim it stands for
// (projects/FastFLIP/nosys/*.cpp, cited per function) on a world laid out like SetFLIPWorld's
// (nosys/FLIP_Creator.cpp:37-116), and marshals OpenVDB leaves to/from the flat leaf arrays of
// include/flipb200.h so that one Python harness (oracle/pyoracle.py) drives the oracle
// restatement, this library and the CUDA library alike.
//
// Used to (1) pin oracle/*.cpp against the real reference (tests/test_ref_pin_cpu.py and the
// committed fixtures under tests/golden/ref_*.npz) and (2) as bench.py's `--impl reference` arm.
#include "FLIP_vdb.h"
#include "levelset_util.h"
#include "simd_vdb_poisson_uaamg.h"
#include "vdb_velocity_extrapolator.h"

#include <openvdb/openvdb.h>
#include <openvdb/points/PointConversion.h>
#include <openvdb/points/PointCount.h>
#include <openvdb/tools/PointIndexGrid.h>

#include <tbb/task_scheduler_init.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

// openvdb/points/AttributeArray.h:353,756 befriends ::TestAttributeArray (OpenVDB's own unit-test
// hook); we use it to move the raw codec words (u16 fixed point / half bits) in and out without a
// decode->encode round trip (which is not the identity for FixedPointCodec, 5461 of 65536 codes move).
class TestAttributeArray {
public:
    static char* bytes(openvdb::points::AttributeArray& a) { return a.dataAsByteArray(); }
    static const char* bytes(const openvdb::points::AttributeArray& a) { return a.constDataAsByteArray(); }
};

namespace {
using openvdb::Coord;
using openvdb::FloatGrid;
using openvdb::Vec3fGrid;
using openvdb::points::PointDataGrid;

enum GridId { G_VELOCITY = 0, G_POSTADV = 1, G_VISCOUS = 2, G_SOLIDVEL = 3, G_FACEWEIGHT = 4,
              G_LIQUIDSDF = 5, G_SOLIDSDF = 6, G_PRESSURE = 7, G_DIVERGENCE = 8, G_CURVATURE = 9, G_KILLERSDF = 10, G_NUM = 11 };

struct RefWorld {
    float dx;
    openvdb::math::Transform::Ptr centre, vertex;
    PointDataGrid::Ptr particles;
    Vec3fGrid::Ptr vec[5];
    FloatGrid::Ptr flt[G_NUM];
    bool hasSolidSDF = false, hasSolidVel = false, hasCurvature = false;
    float density = 1000.f, tensionCoef = 0.f;   // the Density / SurfaceTension sockets
    int iterations = 0, status = 0, levels = 0, numDof = 0;
    float relResidual = 0.f;
    std::vector<float> history;
    uint64_t dropped = 0;

    explicit RefWorld(float dx_) : dx(dx_) {
        // SetFLIPWorld (FF/nosys/FLIP_Creator.cpp:37-116)
        centre = openvdb::math::Transform::createLinearTransform(dx);
        vertex = openvdb::math::Transform::createLinearTransform(dx);
        vertex->postTranslate(openvdb::Vec3d{-0.5, -0.5, -0.5} * double(dx));
        particles = PointDataGrid::create();
        particles->setTransform(centre);
        particles->setName("Particles");
        for (int i = 0; i < 5; i++) {
            vec[i] = Vec3fGrid::create(openvdb::Vec3f{0});
            vec[i]->setTransform(centre);
            vec[i]->setGridClass(openvdb::GridClass::GRID_STAGGERED);
        }
        for (int i = 0; i < G_NUM; i++) flt[i] = nullptr;
        flt[G_PRESSURE] = FloatGrid::create(0.f);
        flt[G_PRESSURE]->setTransform(centre);
        flt[G_PRESSURE]->setGridClass(openvdb::GridClass::GRID_FOG_VOLUME);
        flt[G_DIVERGENCE] = flt[G_PRESSURE]->deepCopy();
        flt[G_LIQUIDSDF] = FloatGrid::create(1.0f * dx);
        flt[G_LIQUIDSDF]->setGridClass(openvdb::GridClass::GRID_LEVEL_SET);
        flt[G_LIQUIDSDF]->setTransform(centre);
        flt[G_SOLIDSDF] = FloatGrid::create(3.0f * dx);
        flt[G_SOLIDSDF]->setTransform(vertex);
        flt[G_SOLIDSDF]->setGridClass(openvdb::GridClass::GRID_LEVEL_SET);
        flt[G_CURVATURE] = FloatGrid::create();
        flt[G_KILLERSDF] = FloatGrid::create(3.0f * dx);   // the "KillerSDF" socket object of KillParticlesInSDF
        flt[G_KILLERSDF]->setTransform(centre);
    }
};

using Mask = openvdb::util::NodeMask<3>;

template <typename GridT> int leafCountOf(const GridT& g) { return int(g.tree().leafCount()); }

void setFloatGrid(FloatGrid& g, int n, const int32_t* origins, const uint64_t* masks, const float* values, float bg) {
    auto tree = std::make_shared<openvdb::FloatTree>(bg);
    for (int l = 0; l < n; l++) {
        auto* leaf = tree->touchLeaf(Coord(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2]));
        Mask m;
        for (int k = 0; k < 8; k++) m.getWord<Mask::Word>(k) = masks[8 * size_t(l) + k];
        std::memcpy(leaf->buffer().data(), values + size_t(l) * 512, sizeof(float) * 512);
        leaf->setValueMask(m);
    }
    g.setTree(tree);
}
void setVecGrid(Vec3fGrid& g, int n, const int32_t* origins, const uint64_t* masks, const float* values, const float* bg) {
    auto tree = std::make_shared<openvdb::Vec3fTree>(openvdb::Vec3f(bg[0], bg[1], bg[2]));
    for (int l = 0; l < n; l++) {
        auto* leaf = tree->touchLeaf(Coord(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2]));
        Mask m;
        for (int k = 0; k < 8; k++) m.getWord<Mask::Word>(k) = masks[8 * size_t(l) + k];
        openvdb::Vec3f* d = leaf->buffer().data();
        const float* v = values + size_t(l) * 3 * 512;
        for (int i = 0; i < 512; i++) d[i] = openvdb::Vec3f(v[i], v[512 + i], v[1024 + i]);
        leaf->setValueMask(m);
    }
    g.setTree(tree);
}
void getFloatGrid(const FloatGrid& g, int32_t* origins, uint64_t* masks, float* values, float* bg) {
    bg[0] = g.background();
    int l = 0;
    for (auto it = g.tree().cbeginLeaf(); it; ++it, ++l) {
        const Coord& o = it->origin();
        origins[3 * l] = o.x(); origins[3 * l + 1] = o.y(); origins[3 * l + 2] = o.z();
        for (int k = 0; k < 8; k++) masks[8 * size_t(l) + k] = it->getValueMask().template getWord<Mask::Word>(k);
        std::memcpy(values + size_t(l) * 512, it->buffer().data(), sizeof(float) * 512);
    }
}
void getVecGrid(const Vec3fGrid& g, int32_t* origins, uint64_t* masks, float* values, float* bg) {
    for (int c = 0; c < 3; c++) bg[c] = g.background()[c];
    int l = 0;
    for (auto it = g.tree().cbeginLeaf(); it; ++it, ++l) {
        const Coord& o = it->origin();
        origins[3 * l] = o.x(); origins[3 * l + 1] = o.y(); origins[3 * l + 2] = o.z();
        for (int k = 0; k < 8; k++) masks[8 * size_t(l) + k] = it->getValueMask().template getWord<Mask::Word>(k);
        const openvdb::Vec3f* d = it->buffer().data();
        float* v = values + size_t(l) * 3 * 512;
        for (int i = 0; i < 512; i++) { v[i] = d[i][0]; v[512 + i] = d[i][1]; v[1024 + i] = d[i][2]; }
    }
}

// the attribute descriptor the reference builds for a new particle tree (FF/FLIP_vdb.cpp:3398-3404)
openvdb::points::AttributeSet::Descriptor::Ptr particleDescriptor() {
    auto pnamepair = FLIP_vdb::position_attribute::attributeType();
    auto descr = openvdb::points::AttributeSet::Descriptor::create(pnamepair);
    auto vnamepair = FLIP_vdb::velocity_attribute::attributeType();
    return descr->duplicateAppend("v", vnamepair);
}

// runs f with stdout captured; returns the text (the reference reports solver progress with printf)
template <typename F> std::string captureStdout(F&& f) {
    fflush(stdout);
    std::cout.flush();
    char path[] = "/tmp/flipref_stdout_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) { f(); return std::string(); }
    int saved = dup(1);
    dup2(fd, 1);
    try { f(); } catch (...) { fflush(stdout); std::cout.flush(); dup2(saved, 1); close(saved); close(fd); unlink(path); throw; }
    fflush(stdout);
    std::cout.flush();
    dup2(saved, 1);
    close(saved);
    std::string out;
    lseek(fd, 0, SEEK_SET);
    char buf[4096];
    ssize_t r;
    while ((r = read(fd, buf, sizeof(buf))) > 0) out.append(buf, size_t(r));
    close(fd);
    unlink(path);
    return out;
}

std::unique_ptr<tbb::task_scheduler_init> g_tbb;
}  // namespace

extern "C" {

const char* ref_build_info() { return "libflipref: reference FastFLIP (FLIP_vdb.cpp, simd_vdb_poisson_uaamg.cpp, vdb_velocity_extrapolator.cpp) + OpenVDB 9.0.1 + TBB 2020, Eigen/Boost stand-ins"; }

// number of TBB worker threads (0 = all hardware threads); returns the count in effect
int ref_set_threads(int n) {
    static bool vdbInit = false;
    if (!vdbInit) { openvdb::initialize(); vdbInit = true; }
    int hw = int(std::thread::hardware_concurrency());
    if (n <= 0) n = hw > 0 ? hw : 1;
    g_tbb.reset();
    g_tbb.reset(new tbb::task_scheduler_init(n));
    return n;
}

void* ref_world_create(float dx) {
    static bool init = false;
    if (!init) { openvdb::initialize(); init = true; }
    return new RefWorld(dx);
}
void ref_world_destroy(void* w) { delete static_cast<RefWorld*>(w); }

int ref_grid_set(void* wp, int id, int nLeaves, const int32_t* origins, const uint64_t* masks, const float* values, const float* bg) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    if (id >= 0 && id <= G_FACEWEIGHT) setVecGrid(*w->vec[id], nLeaves, origins, masks, values, bg);
    else if (id < G_NUM && w->flt[id]) setFloatGrid(*w->flt[id], nLeaves, origins, masks, values, bg[0]);
    else return 1;
    if (id == G_SOLIDSDF) w->hasSolidSDF = true;
    if (id == G_SOLIDVEL) w->hasSolidVel = true;
    if (id == G_CURVATURE) w->hasCurvature = true;
    return 0;
}
int ref_grid_leaf_count(void* wp, int id) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    if (id >= 0 && id <= G_FACEWEIGHT) return leafCountOf(*w->vec[id]);
    if (id < G_NUM && w->flt[id]) return leafCountOf(*w->flt[id]);
    return -1;
}
// the OpenVDB objects themselves, for oracle/ref/plugin_nodes_test.cpp (pointers to the world's grid Ptrs)
void* ref_internal_vec(void* wp, int id) { return (id >= 0 && id <= G_FACEWEIGHT) ? &static_cast<RefWorld*>(wp)->vec[id] : nullptr; }
void* ref_internal_flt(void* wp, int id) { return (id > G_FACEWEIGHT && id < G_NUM) ? &static_cast<RefWorld*>(wp)->flt[id] : nullptr; }
void* ref_internal_particles(void* wp) { return &static_cast<RefWorld*>(wp)->particles; }
int ref_grid_get(void* wp, int id, int32_t* origins, uint64_t* masks, float* values, float* bg) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    if (id >= 0 && id <= G_FACEWEIGHT) getVecGrid(*w->vec[id], origins, masks, values, bg);
    else if (id < G_NUM && w->flt[id]) getFloatGrid(*w->flt[id], origins, masks, values, bg);
    else return 1;
    return 0;
}

// particle store in the reference layout: raw codec words in, no re-encoding
int ref_particles_set(void* wp, int nLeaves, const int32_t* origins, const uint32_t* voxelEnd, uint64_t n,
                      const uint16_t* P, const uint16_t* v) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    auto descr = particleDescriptor();
    auto pdescr = openvdb::points::AttributeSet::Descriptor::create(FLIP_vdb::position_attribute::attributeType());
    auto tree = std::make_shared<openvdb::points::PointDataTree>();
    uint64_t begin = 0;
    for (int l = 0; l < nLeaves; l++) {
        auto* leaf = tree->touchLeaf(Coord(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2]));
        const uint32_t* ve = voxelEnd + size_t(l) * 512;
        const uint32_t cnt = ve[511];
        leaf->initializeAttributes(pdescr, cnt);   // only the position descriptor is accepted here; "v" is appended (FF/FLIP_vdb.cpp:3455-3460)
        leaf->appendAttribute(leaf->attributeSet().descriptor(), descr, 1);
        std::vector<openvdb::PointDataIndex32> offs(512);
        for (int i = 0; i < 512; i++) offs[i] = openvdb::PointDataIndex32(ve[i]);
        leaf->setOffsets(offs, /*updateValueMask=*/true);
        auto& pa = leaf->attributeArray("P");
        auto& va = leaf->attributeArray("v");
        pa.expand(); va.expand();
        std::memcpy(TestAttributeArray::bytes(pa), P + 3 * begin, size_t(cnt) * 6);
        std::memcpy(TestAttributeArray::bytes(va), v + 3 * begin, size_t(cnt) * 6);
        begin += cnt;
    }
    if (begin != n) return 2;
    w->particles->setTree(tree);
    return 0;
}
int ref_particles_info(void* wp, int* nLeaves, uint64_t* n) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    *nLeaves = int(w->particles->tree().leafCount());
    *n = openvdb::points::pointCount(w->particles->tree());
    return 0;
}
int ref_particles_get(void* wp, int32_t* origins, uint32_t* voxelEnd, uint16_t* P, uint16_t* v) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    int l = 0;
    uint64_t begin = 0;
    for (auto it = w->particles->tree().cbeginLeaf(); it; ++it, ++l) {
        const Coord& o = it->origin();
        origins[3 * l] = o.x(); origins[3 * l + 1] = o.y(); origins[3 * l + 2] = o.z();
        uint32_t cnt = 0;
        for (int i = 0; i < 512; i++) { cnt = uint32_t(it->getValue(openvdb::Index(i))); voxelEnd[size_t(l) * 512 + i] = cnt; }
        if (cnt) {
            const auto& pa = it->constAttributeArray("P");
            const auto& va = it->constAttributeArray("v");
            const uint16_t* ps = reinterpret_cast<const uint16_t*>(TestAttributeArray::bytes(pa));
            const uint16_t* vs = reinterpret_cast<const uint16_t*>(TestAttributeArray::bytes(va));
            for (uint32_t i = 0; i < cnt; i++)
                for (int c = 0; c < 3; c++) {
                    P[3 * (begin + i) + c] = ps[pa.isUniform() ? c : 3 * i + c];
                    v[3 * (begin + i) + c] = vs[va.isUniform() ? c : 3 * i + c];
                }
        }
        begin += cnt;
    }
    return 0;
}

// PrimToVDBPointDataGrid: the OpenVDB call sequence of particleArrayToGrid
// (projects/zenvdb/SetVDBPointDataGrid.cpp:17-72; that file is a Zeno node and needs libzeno, so
// its ~15 lines of OpenVDB calls are issued from here in the same order with the same arguments)
int ref_bin_from_points(void* wp, const float* pos, const float* vel, uint64_t n) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    std::vector<openvdb::Vec3f> positions(n), velocitys(n);
    for (uint64_t i = 0; i < n; i++) {
        positions[i] = openvdb::Vec3f(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        velocitys[i] = vel ? openvdb::Vec3f(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]) : openvdb::Vec3f(0.f);
    }
    openvdb::points::PointAttributeVector<openvdb::Vec3f> positionsWrapper(positions);
    openvdb::math::Transform::Ptr transform = openvdb::math::Transform::createLinearTransform(w->dx);
    openvdb::tools::PointIndexGrid::Ptr pointIndexGrid =
        openvdb::tools::createPointIndexGrid<openvdb::tools::PointIndexGrid>(positionsWrapper, *transform);
    auto vnamepair = FLIP_vdb::velocity_attribute::attributeType();
    PointDataGrid::Ptr grid = openvdb::points::createPointDataGrid<FLIP_vdb::PositionCodec, PointDataGrid>(
        *pointIndexGrid, positionsWrapper, *transform);
    openvdb::points::appendAttribute(grid->tree(), "v", vnamepair);
    openvdb::points::PointAttributeVector<openvdb::Vec3f> velocityWrapper(velocitys);
    openvdb::points::populateAttribute<openvdb::points::PointDataTree, openvdb::tools::PointIndexTree,
                                       openvdb::points::PointAttributeVector<openvdb::Vec3f>>(
        grid->tree(), pointIndexGrid->tree(), "v", velocityWrapper);
    grid->setName("Points");
    w->particles = grid;
    return 0;
}

// FLIP_P2G::apply (FF/nosys/P2G.cpp:11-42)
int ref_p2g(void* wp, float dx, int velExtraLayer) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    packed_FloatGrid3 packed_VelGrid, packed_PostP2GVelGrid;
    packed_VelGrid.from_vec3(w->vec[G_VELOCITY]);
    packed_PostP2GVelGrid.from_vec3(w->vec[G_POSTADV]);
    FLIP_vdb::particle_to_grid_collect_style(packed_VelGrid, packed_PostP2GVelGrid, w->flt[G_LIQUIDSDF], w->particles, dx);
    vdb_velocity_extrapolator::union_extrapolate(velExtraLayer, packed_VelGrid.v[0], packed_VelGrid.v[1], packed_VelGrid.v[2],
                                                 &(w->flt[G_LIQUIDSDF]->tree()));
    packed_VelGrid.to_vec3(w->vec[G_VELOCITY]);
    packed_PostP2GVelGrid.to_vec3(w->vec[G_POSTADV]);
    return 0;
}
// G2PAdvectorSheet::apply (FF/nosys/SheetG2PAdvector.cpp:15-54); flags bit0: ViscousVelocity is the Velocity object
int ref_g2p_advect_sheetty(void* wp, float dt, float dx, int surfaceSize, int rkOrder, float picMin, float picMax, int flags) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    picMin = picMin > picMax ? picMax : picMin;
    FloatGrid::Ptr solid_sdf = w->hasSolidSDF ? w->flt[G_SOLIDSDF] : nullptr;
    Vec3fGrid::Ptr solid_vel = w->hasSolidVel ? w->vec[G_SOLIDVEL] : nullptr;
    Vec3fGrid::Ptr& viscous = (flags & 1) ? w->vec[G_VELOCITY] : w->vec[G_VISCOUS];
    const uint64_t before = openvdb::points::pointCount(w->particles->tree());
    FLIP_vdb::AdvectSheetty(dt, dx, float(surfaceSize) * dx, w->particles, w->flt[G_LIQUIDSDF], w->vec[G_VELOCITY], viscous,
                            w->vec[G_POSTADV], solid_sdf, solid_vel, picMin, picMax, rkOrder);
    w->dropped = before - openvdb::points::pointCount(w->particles->tree());
    return 0;
}
uint64_t ref_dropped(void* wp) { return static_cast<RefWorld*>(wp)->dropped; }
// CutCellWeightEval::apply (FF/nosys/EvalFaceWeight.cpp:17-24)
int ref_face_weights(void* wp) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    FLIP_vdb::calculate_face_weights(w->vec[G_FACEWEIGHT], w->flt[G_LIQUIDSDF], w->flt[G_SOLIDSDF]);
    return 0;
}
// PushOutLiquidSDF::apply (FF/nosys/FixLiquidSDF.cpp:16-29)
int ref_pushout_sdf(void* wp, float dx) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    FLIP_vdb::immerse_liquid_phi_in_solids(w->flt[G_LIQUIDSDF], w->flt[G_SOLIDSDF], dx);
    return 0;
}
// FieldAddVector::apply (FF/nosys/FieldAddVector.cpp:16-31)
int ref_add_vector(void* wp, float x, float y, float z) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    packed_FloatGrid3 packed_velocity;
    packed_velocity.from_vec3(w->vec[G_VELOCITY]);
    FLIP_vdb::field_add_vector(packed_velocity, x, y, z, 1.0);
    packed_velocity.to_vec3(w->vec[G_VELOCITY]);
    return 0;
}
// CFL::apply (FF/nosys/CFL.cpp:13-27)
float ref_cfl(void* wp) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    float dt = 0.f;
    captureStdout([&] { dt = FLIP_vdb::cfl(w->vec[G_VELOCITY]); });
    float scaling = w->dx / float(w->vec[G_VELOCITY]->voxelSize()[0]);
    return scaling * dt;
}
// AssembleSolvePPE::apply (FF/nosys/SolvePoissonPressureEqn.cpp:23-64)
int ref_solve_ppe(void* wp, float dt, float dx, int* iters, float* relResidual, int* status) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    FloatGrid::Ptr curvatureGrid = w->hasCurvature ? w->flt[G_CURVATURE] : FloatGrid::create();
    packed_FloatGrid3 packed_velocity;
    packed_velocity.from_vec3(w->vec[G_VELOCITY]);
    std::string log = captureStdout([&] {
        FLIP_vdb::solve_pressure_simd_uaamg(w->flt[G_LIQUIDSDF], curvatureGrid, w->flt[G_DIVERGENCE], w->flt[G_PRESSURE],
                                            w->vec[G_FACEWEIGHT], packed_velocity, w->vec[G_SOLIDVEL], w->density, w->tensionCoef, w->tensionCoef > 0, dt, dx);
    });
    packed_velocity.to_vec3(w->vec[G_VELOCITY]);
    // progress lines printed by the reference: "levels: %zd Dof:%d" (uaamg.cpp:1990), "init error%e" (:2346),
    // "iter:%d err:%e" (:2349,2371), "MGPCG failed, begin pure MG solver" (FF/FLIP_vdb.cpp:3090)
    w->history.clear();
    w->iterations = 0; w->status = 0; w->levels = 0; w->numDof = 0;
    size_t pos = 0;
    bool pure = false;
    while (pos < log.size()) {
        size_t e = log.find('\n', pos);
        if (e == std::string::npos) e = log.size();
        std::string line = log.substr(pos, e - pos);
        pos = e + 1;
        int it; float err; long lv; int nd;
        if (line.find("MGPCG failed") != std::string::npos) { pure = true; w->status = 1; }
        else if (sscanf(line.c_str(), "levels: %ld Dof:%d", &lv, &nd) == 2) { if (w->levels == 0) { w->levels = int(lv); w->numDof = nd; } }
        else if (sscanf(line.c_str(), "init error%e", &err) == 1) { if (!pure) w->history.push_back(err); }
        else if (sscanf(line.c_str(), "iter:%d err:%e", &it, &err) == 2) { if (!pure) { w->history.push_back(err); w->iterations = it - 1; } }
    }
    w->relResidual = w->history.empty() ? 0.f : w->history.back();
    if (iters) *iters = w->iterations;
    if (relResidual) *relResidual = w->relResidual;
    if (status) *status = w->status;
    return 0;
}
// Test hook: the calls of FLIP_vdb::solve_pressure_simd_uaamg (FF/FLIP_vdb.cpp:3034-3100, tension off), made on the reference's own
// solver objects, with the PCG's tolerance and iteration cap chosen by the caller -- the node hard-codes 5e-5 / 100, so its
// pure-multigrid fallback (:3089-3097, uaamg.cpp:2405-2444) cannot be forced through the node. Everything that computes is the
// reference's library code; only the two numbers differ.
int ref_solve_ppe_ex(void* wp, float dt, float dx, float relTol, int maxIter, int* iters, float* relResidual, int* status) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    if (w->flt[G_LIQUIDSDF]->tree().leafCount() == 0) return 0;
    packed_FloatGrid3 velocity;
    velocity.from_vec3(w->vec[G_VELOCITY]);
    int st = 0, it = 0;
    std::string log = captureStdout([&] {
        auto lhs = simd_uaamg::LaplacianWithLevel::createPressurePoissonLaplacian(w->flt[G_LIQUIDSDF], w->vec[G_FACEWEIGHT], dt);
        auto solver = simd_uaamg::PoissonSolver(lhs);
        solver.mRelativeTolerance = relTol;
        solver.mMaxIteration = maxIter;
        solver.mSmoother = simd_uaamg::PoissonSolver::SmootherOption::RedBlackGaussSeidel;
        w->flt[G_DIVERGENCE] = lhs->createPressurePoissonRightHandSide(w->vec[G_FACEWEIGHT], velocity.v[0], velocity.v[1], velocity.v[2], w->vec[G_SOLIDVEL], dt);
        auto pressure = lhs->getZeroVectorGrid();
        pressure->setName("Pressure");
        auto state = solver.solveMultigridPCG(pressure, w->flt[G_DIVERGENCE]);
        it = solver.mIterationTaken;
        if (state != simd_uaamg::PoissonSolver::SUCCESS) {
            st = 1;
            FloatGrid::Ptr old = w->flt[G_PRESSURE];
            lhs->mDofLeafManager->foreach([&](openvdb::Int32Tree::LeafNodeType& leaf, openvdb::Index) {
                auto oldAxr{old->getConstUnsafeAccessor()};
                auto* np = pressure->tree().probeLeaf(leaf.origin());
                for (auto iter = np->beginValueOn(); iter; ++iter) {
                    float v = oldAxr.getValue(iter.getCoord());
                    if (std::isfinite(v)) iter.setValue(v);
                }
            });
            solver.mMaxIteration = 100;
            solver.solvePureMultigrid(pressure, w->flt[G_DIVERGENCE]);
        }
        w->flt[G_PRESSURE] = pressure;
        w->flt[G_DIVERGENCE]->setName("RHS");
    });
    velocity.to_vec3(w->vec[G_VELOCITY]);
    w->status = st; w->iterations = it;
    // residual history: the PCG's lines, then (after "pure" would start) the pure-multigrid iteration's "iter:%d err:%e" lines
    w->history.clear();
    {
        size_t pos = 0;
        while (pos < log.size()) {
            size_t e = log.find('\n', pos);
            if (e == std::string::npos) e = log.size();
            std::string line = log.substr(pos, e - pos);
            pos = e + 1;
            int k; float err;
            if (sscanf(line.c_str(), "iter:%d err:%e", &k, &err) == 2) w->history.push_back(err);
        }
    }
    if (iters) *iters = it;
    if (relResidual) *relResidual = w->history.empty() ? 0.f : w->history.back();
    if (status) *status = st;
    (void)dx;
    return 0;
}
int ref_solver_info(void* wp, int* levels, int* numDof, int* nHistory) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    *levels = w->levels; *numDof = w->numDof; *nHistory = int(w->history.size());
    return 0;
}
int ref_residual_history(void* wp, float* out) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    std::memcpy(out, w->history.data(), sizeof(float) * w->history.size());
    return 0;
}
int ref_set_surface_tension(void* wp, float density, float coef) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    w->density = density; w->tensionCoef = coef;
    return 0;
}
float ref_density(void* wp) { return static_cast<RefWorld*>(wp)->density; }
float ref_tension_coef(void* wp) { return static_cast<RefWorld*>(wp)->tensionCoef; }
// SubtractPressureGradient::apply (FF/nosys/SubtractPressureGradient.cpp:25-66)
int ref_subtract_grad(void* wp, float dt, float dx, int velExtraLayer) {
    RefWorld* w = static_cast<RefWorld*>(wp);
    FloatGrid::Ptr curvatureGrid = w->hasCurvature ? w->flt[G_CURVATURE] : FloatGrid::create();
    packed_FloatGrid3 packed_velocity;
    packed_velocity.from_vec3(w->vec[G_VELOCITY]);
    FLIP_vdb::apply_pressure_gradient(w->flt[G_LIQUIDSDF], w->flt[G_SOLIDSDF], w->flt[G_PRESSURE], w->vec[G_FACEWEIGHT], packed_velocity,
                                      w->vec[G_SOLIDVEL], curvatureGrid, w->density, w->tensionCoef, w->tensionCoef > 0, dt, dx);
    vdb_velocity_extrapolator::union_extrapolate(velExtraLayer, packed_velocity.v[0], packed_velocity.v[1], packed_velocity.v[2],
                                                 &(w->flt[G_LIQUIDSDF]->tree()));
    packed_velocity.to_vec3(w->vec[G_VELOCITY]);
    return 0;
}
int ref_capture_precodec(void*, int) { return 1; }  // not available: the reference encodes in place
int ref_get_precodec(void*, float*, float*, uint8_t*) { return 1; }

float ref_fraction_inside2(float a, float b) { return fraction_inside(a, b); }
float ref_fraction_inside4(float bl, float br, float tl, float tr) { return fraction_inside(bl, br, tl, tr); }

// one substep of the test chain; per-stage seconds (same stage split as flipb200_substep)
int ref_substep(void* wp, float dt, float dx, int surfaceSize, int rkOrder, float picMin, float picMax, float gx, float gy,
                float gz, int velExtraLayer, int flags, double* stageSeconds) {
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    auto t0 = now();
    ref_g2p_advect_sheetty(wp, dt, dx, surfaceSize, rkOrder, picMin, picMax, flags);
    auto t1 = now();
    ref_p2g(wp, dx, velExtraLayer);
    auto t2 = now();
    ref_face_weights(wp);
    ref_pushout_sdf(wp, dx);
    ref_add_vector(wp, gx * dt, gy * dt, gz * dt);
    auto t3 = now();
    ref_solve_ppe(wp, dt, dx, nullptr, nullptr, nullptr);
    auto t4 = now();
    ref_subtract_grad(wp, dt, dx, velExtraLayer);
    auto t5 = now();
    if (stageSeconds) {
        stageSeconds[0] = secs(t0, t1); stageSeconds[1] = secs(t1, t2); stageSeconds[2] = secs(t2, t3);
        stageSeconds[3] = secs(t3, t4); stageSeconds[4] = secs(t4, t5);
    }
    return 0;
}

}  // extern "C"

// =====================================================================================================
// Marshalling test of the Zeno-side drop-in (zeno_b200/plugin/flipb200_nodes.cpp): its upload()/download() /
// upload_particles()/download_particles() are compiled here against a LOOPBACK C ABI (the flipb200_* entry points it
// calls are renamed to lb_* and implemented below: they keep the flat leaf arrays and hand them back with the leaves
// in REVERSE order, as a device library is free to reorder them), and REAL reference objects -- the grids and the
// particle tree this world holds after reference nodes ran -- are round-tripped through them.
#define FLIPB200_PLUGIN_MARSHAL_ONLY
#define FLIPB200_HAVE_TEST_ATTRIBUTE_ARRAY
#define flipb200_last_error lb_last_error
#define flipb200_world_create lb_world_create
#define flipb200_world_destroy lb_world_destroy
#define flipb200_grid_upload lb_grid_upload
#define flipb200_grid_leaf_count lb_grid_leaf_count
#define flipb200_grid_download lb_grid_download
#define flipb200_particles_upload lb_particles_upload
#define flipb200_particles_info lb_particles_info
#define flipb200_particles_download lb_particles_download
#define flipb200_host_alloc lb_host_alloc
#define flipb200_host_free lb_host_free
#define flipb200_grid_download_begin lb_grid_download_begin
#define flipb200_particles_download_begin lb_particles_download_begin
#define flipb200_download_wait lb_download_wait
#include <map>
// (the plugin's namespace zeno::flipb200 is renamed in this translation unit: plugin_nodes_test.cpp compiles the same
// inline functions against a different ABI, and the two sets must not be merged by the linker)
#define flipb200 flipb200_loopback
#include "../../zeno_b200/plugin/flipb200_nodes.cpp"
#undef flipb200

struct flipb200_world {
    struct G { int n = 0, nch = 1, layout = 0; std::vector<int32_t> o; std::vector<uint64_t> m; std::vector<float> v; float bg[3] = {0, 0, 0}; };
    std::map<int, G> grids;
    int nl = 0;
    uint64_t np = 0;
    std::vector<int32_t> po;
    std::vector<uint32_t> ve;
    std::vector<uint16_t> P, V;
};
extern "C" {
const char* lb_last_error(void) { return "loopback"; }
int lb_world_create(int, float, flipb200_world** out) { *out = new flipb200_world(); return 0; }
int lb_world_destroy(flipb200_world* w) { delete w; return 0; }
int lb_grid_upload(flipb200_world* w, int grid, int n, const int32_t* o, const uint64_t* m, const float* v, int layout, const float* bg) {
    auto& g = w->grids[grid];
    g.n = n; g.nch = grid <= FLIPB200_FACE_WEIGHT ? 3 : 1; g.layout = layout;
    g.o.assign(o, o + 3 * size_t(n)); g.m.assign(m, m + 8 * size_t(n)); g.v.assign(v, v + size_t(512) * g.nch * n);
    for (int c = 0; c < g.nch; c++) g.bg[c] = bg[c];
    return 0;
}
int lb_grid_leaf_count(flipb200_world* w, int grid, int* n) { *n = w->grids[grid].n; return 0; }
int lb_grid_download(flipb200_world* w, int grid, int32_t* o, uint64_t* m, float* v, int layout, float* bg) {
    auto& g = w->grids[grid];
    if (layout != g.layout) return 1;
    const size_t per = size_t(512) * g.nch;
    for (int i = 0; i < g.n; i++) {   // reversed leaf order
        const int s = g.n - 1 - i;
        std::memcpy(o + 3 * size_t(i), &g.o[3 * size_t(s)], 12);
        std::memcpy(m + 8 * size_t(i), &g.m[8 * size_t(s)], 64);
        std::memcpy(v + per * i, &g.v[per * s], per * 4);
    }
    for (int c = 0; c < g.nch; c++) bg[c] = g.bg[c];
    return 0;
}
int lb_particles_upload(flipb200_world* w, int nl, const int32_t* o, const uint32_t* ve, uint64_t np, const uint16_t* P, const uint16_t* V) {
    w->nl = nl; w->np = np;
    w->po.assign(o, o + 3 * size_t(nl)); w->ve.assign(ve, ve + 512 * size_t(nl));
    w->P.assign(P, P + 3 * np); w->V.assign(V, V + 3 * np);
    return 0;
}
int lb_particles_info(flipb200_world* w, int* nl, uint64_t* np) { *nl = w->nl; *np = w->np; return 0; }
int lb_particles_download(flipb200_world* w, int32_t* o, uint32_t* ve, uint16_t* P, uint16_t* V) {
    // reversed leaf order; the attribute arrays follow the leaves
    std::vector<uint64_t> begin(size_t(w->nl) + 1, 0);
    for (int i = 0; i < w->nl; i++) begin[i + 1] = begin[i] + w->ve[512 * size_t(i) + 511];
    uint64_t at = 0;
    for (int i = 0; i < w->nl; i++) {
        const int s = w->nl - 1 - i;
        std::memcpy(o + 3 * size_t(i), &w->po[3 * size_t(s)], 12);
        std::memcpy(ve + 512 * size_t(i), &w->ve[512 * size_t(s)], 2048);
        const uint64_t cnt = begin[s + 1] - begin[s];
        std::memcpy(P + 3 * at, &w->P[3 * begin[s]], cnt * 6);
        std::memcpy(V + 3 * at, &w->V[3 * begin[s]], cnt * 6);
        at += cnt;
    }
    return 0;
}
}  // extern "C"

extern "C" {
int lb_host_alloc(size_t bytes, void** out) { *out = std::malloc(bytes); return *out ? 0 : 2; }
int lb_host_free(void* p) { std::free(p); return 0; }
int lb_download_wait(flipb200_world*) { return 0; }
int lb_grid_download_begin(flipb200_world* w, int grid, int cap, int32_t* o, uint64_t* m, float* v, int layout, float* bg, int* n) {
    *n = w->grids[grid].n;
    if (*n > cap) return 1;
    return lb_grid_download(w, grid, o, m, v, layout, bg);
}
int lb_particles_download_begin(flipb200_world* w, int capLeaves, uint64_t capParticles, int32_t* o, uint32_t* ve, uint16_t* P, uint16_t* V, int* nl, uint64_t* np) {
    *nl = w->nl; *np = w->np;
    if (*nl > capLeaves || *np > capParticles) return 1;
    return lb_particles_download(w, o, ve, P, V);
}
}  // extern "C"

namespace {
template <typename GridT>
bool sameGrid(const GridT& a, const GridT& b, std::string& why, const char* name) {
    using Leaf = typename GridT::TreeType::LeafNodeType;
    if (!(a.background() == b.background())) { why = std::string(name) + ": background differs"; return false; }
    if (a.tree().leafCount() != b.tree().leafCount()) { why = std::string(name) + ": leaf count differs"; return false; }
    if (a.activeVoxelCount() != b.activeVoxelCount()) { why = std::string(name) + ": active voxel count differs"; return false; }
    for (auto it = a.tree().cbeginLeaf(); it; ++it) {
        const Leaf* lb = b.tree().probeConstLeaf(it->origin());
        if (!lb) { why = std::string(name) + ": a leaf is missing"; return false; }
        if (!(it->getValueMask() == lb->getValueMask())) { why = std::string(name) + ": a value mask differs"; return false; }
        if (std::memcmp(it->buffer().data(), lb->buffer().data(), sizeof(typename GridT::ValueType) * 512) != 0) { why = std::string(name) + ": leaf values differ"; return false; }
    }
    return true;
}
bool sameParticles(const PointDataGrid& a, const PointDataGrid& b, std::string& why) {
    if (a.tree().leafCount() != b.tree().leafCount()) { why = "particles: leaf count differs"; return false; }
    if (openvdb::points::pointCount(a.tree()) != openvdb::points::pointCount(b.tree())) { why = "particles: point count differs"; return false; }
    for (auto it = a.tree().cbeginLeaf(); it; ++it) {
        const auto* lb = b.tree().probeConstLeaf(it->origin());
        if (!lb) { why = "particles: a leaf is missing"; return false; }
        for (openvdb::Index k = 0; k < 512; k++)
            if (it->getValue(k) != lb->getValue(k)) { why = "particles: voxel offsets differ"; return false; }
        if (!(it->getValueMask() == lb->getValueMask())) { why = "particles: value mask differs"; return false; }
        const openvdb::Index cnt = it->getLastValue();
        if (!cnt) continue;
        for (const char* attr : {"P", "v"}) {
            const auto& xa = it->constAttributeArray(attr);
            const auto& xb = lb->constAttributeArray(attr);
            if (xa.type() != xb.type()) { why = std::string("particles: attribute type of ") + attr + " differs"; return false; }
            const uint16_t* pa = reinterpret_cast<const uint16_t*>(TestAttributeArray::bytes(xa));
            const uint16_t* pb = reinterpret_cast<const uint16_t*>(TestAttributeArray::bytes(xb));
            for (openvdb::Index j = 0; j < cnt; j++)
                for (int c = 0; c < 3; c++)
                    if (pa[xa.isUniform() ? c : 3 * j + c] != pb[xb.isUniform() ? c : 3 * j + c]) { why = std::string("particles: codes of ") + attr + " differ"; return false; }
        }
    }
    return true;
}
}  // namespace

extern "C" int ref_plugin_roundtrip(void* h, char* msg, int cap) {
    // 0 = every grid and the particle tree of this world survive upload() -> loopback ABI -> download() unchanged
    RefWorld& w = *static_cast<RefWorld*>(h);
    std::string why;
    bool ok = true;
    try {
        zeno::flipb200_loopback::WorldHolder holder;
        holder.dx = w.dx;
        lb_world_create(0, w.dx, &holder.w);
        const char* vnames[5] = {"Velocity", "PostAdvVelocity", "ViscousVelocity", "SolidVelocity", "CellFWeight"};
        for (int i = 0; i < 5 && ok; i++) {
            if (!w.vec[i]) continue;
            zeno::flipb200_loopback::upload<Vec3fGrid>(holder, i, w.vec[i]);
            Vec3fGrid::Ptr back = Vec3fGrid::create(openvdb::Vec3f(-7.f));
            zeno::flipb200_loopback::download<Vec3fGrid>(holder, i, back);
            ok = sameGrid(*w.vec[i], *back, why, vnames[i]);
        }
        const char* fnames[10] = {"", "", "", "", "", "LiquidSDF", "SolidSDF", "Pressure", "Divergence", "Curvature"};
        for (int i = 5; i < 10 && ok; i++) {
            if (!w.flt[i]) continue;
            zeno::flipb200_loopback::upload<FloatGrid>(holder, i, w.flt[i]);
            FloatGrid::Ptr back = FloatGrid::create(-7.f);
            zeno::flipb200_loopback::download<FloatGrid>(holder, i, back);
            ok = sameGrid(*w.flt[i], *back, why, fnames[i]);
        }
        if (ok) {
            zeno::flipb200_loopback::upload_particles(holder, w.particles);
            PointDataGrid::Ptr back = PointDataGrid::create();
            zeno::flipb200_loopback::download_particles(holder, back);
            ok = sameParticles(*w.particles, *back, why);
        }
    } catch (const std::exception& e) { ok = false; why = std::string("exception: ") + e.what(); }
    if (msg && cap > 0) { std::strncpy(msg, why.c_str(), size_t(cap) - 1); msg[cap - 1] = 0; }
    return ok ? 0 : 1;
}
