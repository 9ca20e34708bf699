// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FastFLIP hot path (see sgrid.h).
// A plain, single-file-per-stage C++ restatement of the reference algorithms;
// every function cites the reference file:line it follows
// (paths relative to /root/reference; FF = projects/FastFLIP).
//
// PARITY STATUS: the reference ships no tests or golden vectors for this path
// (SURVEY.md section 4). The oracle is pinned against the real reference
// sources compiled by oracle/ref/build_ref.sh when that build is available
// (oracle/_ref/); see DESIGN.md "Oracle" for what is pinned and what is not.
#pragma once
#include "sgrid.h"

namespace orc {

// ---- codecs (openvdb/points/AttributeArray.h:50-65,940-976; math/Half.h:430-490)
uint16_t fxpt16_encode(float p);
float fxpt16_decode(uint16_t u);
uint16_t half_encode(float f);
float half_decode(uint16_t h);

struct Packed3 {  // packed_FloatGrid3 (projects/zenvdb/include/zeno/packed3grids.h:6-26)
    FloatGrid v[3];
};
void from_vec3(Packed3& out, const Vec3Grid& in, bool topologyOnly);
void to_vec3(Vec3Grid& out, const Packed3& in);

struct World {
    float dx;
    Points particles;
    Vec3Grid velocity, postAdvVelocity, viscousVelocity, solidVelocity, faceWeight;
    FloatGrid liquidSDF, solidSDF, pressure, divergence, curvature, killerSDF;
    bool hasSolidSDF = false, hasSolidVel = false;
    // diagnostics
    int pcgIterations = 0;
    int pcgStatus = 0;
    // AssembleSolvePPE / SubtractPressureGradient sockets Density and SurfaceTension (FF/nosys/SolvePoissonPressureEqn.cpp:43-45):
    // tension is enabled when the coefficient is > 0 and enters as 2 coef / density (FF/FLIP_vdb.cpp:3052, :2875)
    float density = 1000.f, tensionCoef = 0.f;
    float solveRelTol = 5e-5f;   // AssembleSolvePPE: the node's mRelativeTolerance / mMaxIteration (uaamg.cpp defaults), overridable by the tests
    int solveMaxIter = 100;
    int mgLevels = 0;
    int numDof = 0;
    float pcgRelResidual = 0.f;
    uint64_t droppedParticles = 0;
    std::vector<float> residualHistory;
    // optional pre-codec capture (SURVEY 8d "codec caveat"): fp32 particle
    // position (index space) and velocity before the codecs, in source order.
    bool capturePreCodec = false;
    std::vector<float> preCodecPos, preCodecVel;
    std::vector<uint8_t> preCodecAlive;
    explicit World(float dx_);
};

// K1  projects/zenvdb/SetVDBPointDataGrid.cpp:17-72
void bin_from_points(World& w, const float* pos, const float* vel, size_t n);
// K3-K6  FF/nosys/P2G.cpp:11-42
void node_FLIP_P2G(World& w, float dx, int velExtraLayer);
// K5  FF/vdb_velocity_extrapolator.cpp:584-661
void union_extrapolate(int nLayer, Packed3& v, const FloatGrid* targetTopo);
// K2,K7,K8  FF/nosys/SheetG2PAdvector.cpp:15-54 / FF/FLIP_vdb.cpp:3237-3490
void node_G2PAdvectorSheetty(World& w, float dt, float dx, int surfaceSize, int rkOrder,
                             float picMin, float picMax);
// the plain node: FF/nosys/G2P_Advector.cpp:16-47 -> FLIP_vdb::Advect (FF/FLIP_vdb.cpp:3209-3219)
void node_G2P_Advector(World& w, float dt, float dx, int rkOrder, float picSmoothness);
// K14 nodes
void node_CutCellWeight(World& w);                          // FF/FLIP_vdb.cpp:2644-2718
void node_PushOutLiquidSDF(World& w, float dx);             // FF/FLIP_vdb.cpp:2720-2803
void node_FieldAddVector(World& w, float x, float y, float z);  // FF/FLIP_vdb.cpp:3145-3158
float node_CFL_dt(World& w);                                // FF/FLIP_vdb.cpp:3160-3207
// K9-K13  FF/nosys/SolvePoissonPressureEqn.cpp:23-64
void node_AssembleSolvePPE(World& w, float dt, float dx);
// FF/nosys/SubtractPressureGradient.cpp:25-66
void node_SubtractPressureGradient(World& w, float dt, float dx, int velExtraLayer);

// FF/nosys/KillParticles.cpp:13-158: keep the particles whose KillerSDF sample is <= 0 (keep) / >= 0 (delete)
void node_KillParticlesInSDF(World& w, const FloatGrid& sdf, bool keep);

// FF/nosys/FLIP_Reseed.cpp:8-16 -> FLIP_vdb::reseed_fluid (FF/FLIP_vdb.cpp:2047-2220)
void node_FluidReseed(World& w, uint32_t seed, const uint64_t* leafStart, uint64_t* leafEnd);
uint64_t reseed_leaf_start(uint32_t seed, int ox, int oy, int oz);
// FF/nosys/ParticleEmitter.cpp:9-40 -> FLIP_vdb::emit_liquid (FF/FLIP_vdb.cpp:2222-2642), constant-velocity branch; leafStart / leafEnd
// are indexed by the leaves of the RESULT (store order); leafEnd = ~0 for leaves the shape does not touch
void node_ParticleEmitter(World& w, const FloatGrid& shape, float vx, float vy, float vz, uint32_t seed, const uint64_t* leafStart, uint64_t* leafEnd);

// FF/nosys/Update_Solid_SDF.cpp:9-31 -> FLIP_vdb::update_solid_sdf (FF/FLIP_vdb.cpp:1976-2046), one moving solid
void node_FLIPApplyBoundary(World& w, const FloatGrid& moving, bool movingVertexCentred);

// projects/zenvdb/GetVDBPoints.cpp:76-258: pos / vel = [N][3] floats, vel may be null
void node_VDBPointsToPrimitive(const World& w, float* pos, float* vel);

// FF/nosys/ParticleAddGravity.cpp:9-19 -> FLIP_vdb::point_integrate_vector (FF/FLIP_vdb.cpp:3492-3535)
void node_ParticleAddDV(World& w, float dvx, float dvy, float dvz);

// VDBRenormalizeSDF (projects/zenvdb/VDBRenormalize.cpp:18-37) = openvdb::tools::LevelSetTracker::normalize x iterations with
// {FIRST_BIAS, TVD_RK3, 1, 1}, trimming off, no dilation (tools/LevelSetTracker.h:510-655)
void node_VDBRenormalizeSDF(FloatGrid& g, float voxelSize, int iterations);

// VDBErodeSDF (projects/zenvdb/VDBRenormalize.cpp:155-172): every ACTIVE voxel += depth
void node_VDBErodeSDF(FloatGrid& g, float depth);

// VDBSmoothSDF (projects/zenvdb/VDBRenormalize.cpp:108-120) = openvdb::tools::Filter::gaussian (tools/Filter.h:503-514,574-630,744-768)
void node_VDBSmoothSDF(FloatGrid& g, int width, int iterations);

float fraction_inside(float phi_left, float phi_right);  // FF/levelset_util.cpp:5-15
float fraction_inside(float bl, float br, float tl, float tr);  // FF/levelset_util.cpp:26-99

}  // namespace orc
