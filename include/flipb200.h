/* flipb200.h -- C ABI of libflipb200: the B200 (sm_100a) implementation of the hot path of
 * Zeno's FastFLIP solver.  This is the drop-in boundary: the Zeno node shims in
 * zeno_b200/plugin/ (same ZENDEFNODE names and sockets as projects/FastFLIP/nosys/ in the
 * reference) unpack their sockets, flatten the OpenVDB leaves into the arrays below and
 * call these entry points instead of the FLIP_vdb statics. Each entry point cites the
 * reference interface it replaces (paths relative to the reference root, FF = projects/FastFLIP).
 *
 * Conventions
 *  - plain C: opaque handles, pointers + sizes, int error codes (0 = ok); the message of
 *    the last error on the calling thread is returned by flipb200_last_error().
 *  - all calls block the host thread until their device work is complete, unless stated.
 *  - pointers are HOST pointers unless the name ends in _dev.
 *  - there is no CPU fallback: every entry point fails with FLIPB200_ERR_CUDA when no
 *    sm_100 device is usable.
 *
 * Data layout shared by every grid call (exactly OpenVDB's leaf layout, so a shim can
 * memcpy leaf buffers):
 *  - a grid is a set of 8^3 leaves; leaf l has origin[l] = (x,y,z) int32, multiples of 8
 *    (openvdb/tree/LeafNode.h:1051-1057: voxel offset = x<<6 | y<<3 | z)
 *  - active mask: 8 x uint64 per leaf, bit n of the 512-bit mask is word n>>6, bit n&63
 *    (openvdb/util/NodeMasks.h NodeMask<3>)
 *  - values: float32; scalar grids [leaf][512]; vector grids either
 *    FLIPB200_SOA  [leaf][3][512]  or  FLIPB200_AOS [leaf][512][3] (= Vec3f leaf buffer)
 *  - a voxel in no leaf reads as (background, inactive), like a VDB accessor.
 *  - particles (openvdb::points::PointDataGrid with the FastFLIP codecs, FF/FLIP_vdb.h:28-37):
 *    per leaf 512 uint32 cumulative end offsets in voxel order, attribute "P" = 3 x uint16
 *    (FixedPointCodec<false>, voxel-local [-0.5,0.5)) and "v" = 3 x IEEE half bits
 *    (TruncateCodec), concatenated over leaves in the order of origin[].
 */
#ifndef FLIPB200_H
#define FLIPB200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct flipb200_world flipb200_world;

enum {
    FLIPB200_OK = 0,
    FLIPB200_ERR_ARG = 1,      /* bad argument */
    FLIPB200_ERR_CUDA = 2,     /* CUDA runtime / no device */
    FLIPB200_ERR_DOMAIN = 3,   /* leaf bounding box too large for the dense directory */
    FLIPB200_ERR_STATE = 4,    /* call needs state that has not been provided */
    FLIPB200_ERR_COMM = 5      /* NCCL */
};

/* Grid ids = the world objects created by SetFLIPWorld (FF/nosys/FLIP_Creator.cpp:121-143). */
enum {
    FLIPB200_VELOCITY = 0,        /* "Velocity"          Vec3f staggered */
    FLIPB200_POSTADV_VELOCITY = 1,/* "PostAdvVelocity"   Vec3f staggered (velocity right after P2G) */
    FLIPB200_VISCOUS_VELOCITY = 2,/* "ViscousVelocity"   Vec3f staggered */
    FLIPB200_SOLID_VELOCITY = 3,  /* "SolidVelocity"     Vec3f staggered */
    FLIPB200_FACE_WEIGHT = 4,     /* "CellFWeight"       Vec3f staggered */
    FLIPB200_LIQUID_SDF = 5,      /* "LiquidSDF"         float, cell centred, background dx */
    FLIPB200_SOLID_SDF = 6,       /* "SolidSDF"          float, VERTEX centred, background 3dx */
    FLIPB200_PRESSURE = 7,        /* "Pressure"          float */
    FLIPB200_DIVERGENCE = 8,      /* "Divergence"        float (the PPE right-hand side) */
    FLIPB200_CURVATURE = 9,       /* "Curvature"         float, read when flipb200_set_surface_tension enabled the tension terms */
    FLIPB200_KILLER_SDF = 10,     /* "KillerSDF" socket of KillParticlesInSDF: any float grid, sampled in ITS index space */
    FLIPB200_NUM_GRIDS = 11
};
enum { FLIPB200_SOA = 0, FLIPB200_AOS = 1 };

const char* flipb200_last_error(void);
/* compile-time facts for the loader test: "sm_100a", CUDA runtime version, ABI version */
const char* flipb200_build_info(void);
int flipb200_abi_version(void);
int flipb200_device_count(void);

/* SetFLIPWorld (FF/nosys/FLIP_Creator.cpp:9-118): creates the empty world objects with the
 * reference backgrounds (liquid SDF = dx, solid SDF = 3dx, everything else 0). */
int flipb200_world_create(int device, float dx, flipb200_world** out);
int flipb200_world_destroy(flipb200_world* w);

/* VDB <-> device marshalling (replaces the shared openvdb grid objects behind
 * VDBGridWrapper::m_grid, projects/zenvdb/include/zeno/VDBGrid.h:80-86). */
int flipb200_grid_upload(flipb200_world* w, int grid, int nLeaves, const int32_t* origins,
                         const uint64_t* masks, const float* values, int layout, const float* background);
int flipb200_grid_leaf_count(flipb200_world* w, int grid, int* nLeaves);
/* downloads only leaves with at least one active voxel unless keepEmpty != 0 */
int flipb200_grid_download(flipb200_world* w, int grid, int32_t* origins, uint64_t* masks, float* values,
                           int layout, float* background);
int flipb200_particles_upload(flipb200_world* w, int nLeaves, const int32_t* origins,
                              const uint32_t* voxelEnd, uint64_t nParticles, const uint16_t* P,
                              const uint16_t* v);
int flipb200_particles_info(flipb200_world* w, int* nLeaves, uint64_t* nParticles);
/* Asynchronous downloads: *_begin stages the data as of this point of the node sequence and returns the counts; the
 * device->host copies run on a second stream and overlap the node calls that follow (e.g. particles are final after
 * G2PAdvectorSheetty and cross PCIe while P2G and the pressure solve run). The host buffers (page-locked, e.g. from
 * flipb200_host_alloc; their capacities are passed in and FLIPB200_ERR_ARG is returned, before anything is written, if the
 * data does not fit) are valid after flipb200_download_wait.
 * This is the lazy write-back a node shim uses when a downstream un-accelerated node needs the OpenVDB object. */
int flipb200_particles_download_begin(flipb200_world* w, int capLeaves, uint64_t capParticles, int32_t* origins,
                                      uint32_t* voxelEnd, uint16_t* P, uint16_t* v, int* nLeaves, uint64_t* nParticles);
int flipb200_grid_download_begin(flipb200_world* w, int grid, int capLeaves, int32_t* origins, uint64_t* masks,
                                 float* values, int layout, float* background, int* nLeaves);
int flipb200_download_wait(flipb200_world* w);
int flipb200_particles_download(flipb200_world* w, int32_t* origins, uint32_t* voxelEnd, uint16_t* P,
                                uint16_t* v);

/* Page-locked host staging buffers for the marshalling calls above (what the node shims use to hand VDB leaf
 * buffers to the device at PCIe speed; pageable memory works too, at a third of the bandwidth). */
int flipb200_host_alloc(size_t bytes, void** out);
int flipb200_host_free(void* p);

/* K1: PrimToVDBPointDataGrid / particleArrayToGrid (projects/zenvdb/SetVDBPointDataGrid.cpp:17-72):
 * world positions + velocities -> voxel-sorted quantised particle store. vel may be NULL (zeros). */
int flipb200_bin_from_points(flipb200_world* w, const float* pos, const float* vel, uint64_t n);

/* FLIP_P2G::apply (FF/nosys/P2G.cpp:11-42): FLIP_vdb::particle_to_grid_collect_style
 * (FF/FLIP_vdb.cpp:1282-1388) + union_extrapolate (FF/vdb_velocity_extrapolator.cpp:584-661).
 * in: particles; out: Velocity, PostAdvVelocity, LiquidSDF. */
int flipb200_p2g(flipb200_world* w, float dx, int velExtraLayer);

/* G2PAdvectorSheet::apply (FF/nosys/SheetG2PAdvector.cpp:15-54) -> FLIP_vdb::AdvectSheetty ->
 * custom_move_points_and_set_flip_vel (FF/FLIP_vdb.cpp:3221-3490).
 * flags bit0: the ViscousVelocity socket carries the Velocity object itself. */
int flipb200_g2p_advect_sheetty(flipb200_world* w, float dt, float dx, int surfaceSize, int rkOrder,
                                float picMin, float picMax, int flags);
/* particles dropped by the last advect (deep in solid, FF/FLIP_vdb.cpp:682-685, or voxel cap :711-714) */
/* G2P_Advector (FF/nosys/G2P_Advector.cpp:16-47 -> FLIP_vdb::Advect, FF/FLIP_vdb.cpp:3209-3219): the plain node. It passes no liquid
 * SDF and surfacedist 0, so every particle takes one Euler step (RK_ORDER is accepted and has no effect, as in the reference) with
 * the FLIP factor 1 - pic_smoothness, and it is only usable without solids: with a SolidSDF connected the reference dereferences
 * the null liquid SDF. Particles + Velocity + PostAdvVelocity in, re-binned particles out. */
int flipb200_g2p_advect(flipb200_world* w, float dt, float dx, int rkOrder, float picSmoothness);
/* VDBRenormalizeSDF (projects/zenvdb/VDBRenormalize.cpp:18-52): openvdb::tools::LevelSetTracker::normalize()
 * `iterations` times with {FIRST_BIAS, TVD_RK3}, trimming off -- three Euler stages of the first-order upwind (Godunov)
 * re-distancing per call -- on the ACTIVE voxels of float grid `grid` (its topology and inactive values do not change).
 * dilateIters must be 0 (the tracker's dilate / erode is not accelerated: FLIPB200_ERR_ARG). */
int flipb200_renormalize_sdf(flipb200_world* w, int grid, int iterations, int dilateIters);
/* VDBErodeSDF (projects/zenvdb/VDBRenormalize.cpp:155-185, used once by the packaged FLIP template): adds `depth` to every
 * active voxel of float grid `grid`. */
int flipb200_erode_sdf(flipb200_world* w, int grid, float depth);
/* VDBSmoothSDF (projects/zenvdb/VDBRenormalize.cpp:108-133, used once by the packaged FLIP template) =
 * openvdb::tools::Filter::gaussian(width, iterations) without mask or tiles: per iteration four separable box filters
 * (passes X, Z, Y of 2*width+1 taps) on the active voxels of float grid `grid`. */
int flipb200_smooth_sdf(flipb200_world* w, int grid, int width, int iterations);
int flipb200_dropped(flipb200_world* w, uint64_t* n);
/* ParticleAddDV (FF/nosys/ParticleAddGravity.cpp:9-19 -> FLIP_vdb::point_integrate_vector, FF/FLIP_vdb.cpp:3492-3535; SURVEY 8b
 * last row / 8f-1): adds dv to the stored velocity of every particle -- read as double from the half codec, summed in double,
 * written back float -> half. Positions are untouched. */
int flipb200_particles_add_dv(flipb200_world* w, float dvx, float dvy, float dvz);
/* KillParticlesInSDF (FF/nosys/KillParticles.cpp:13-165; SURVEY 8f-1, the first node beyond the substep chain): every particle
 * samples float grid `sdfGrid` (normally FLIPB200_KILLER_SDF) with OpenVDB's BoxSampler at voxel + position IN THAT GRID'S INDEX
 * SPACE and survives when the sample is <= 0 (keep != 0, OpType KEEP) or >= 0 (OpType DEL); survivors' positions go through the
 * codec once more, as the reference's write-back does. */
int flipb200_kill_particles_in_sdf(flipb200_world* w, int sdfGrid, int keep);
/* FluidReseed (FF/nosys/FLIP_Reseed.cpp:8-16 -> FLIP_vdb::reseed_fluid, FF/FLIP_vdb.cpp:2047-2220): every voxel of
 * a particle leaf whose LiquidSDF value at the centre is < dx and that holds <= 4 particles is topped up -- up to 16 jittered
 * candidates from the reference's hash table (FF/FLIP_vdb.h:10-20), taken while the voxel holds < 8 when the SDF at the
 * candidate is <= -dx and its octant is empty, with the StaggeredBoxSampler velocity of FLIPB200_VELOCITY there. Reads
 * FLIPB200_LIQUID_SDF and FLIPB200_VELOCITY, replaces the particle store (same leaves). The reference starts the table at a
 * std::random_device draw per TBB chunk (:2081-2084); here leaf (ox, oy, oz) starts at
 *   h = seed ^ ox*73856093 ^ oy*19349663 ^ oz*83492791;  h ^= h>>16; h *= 0x7feb352d; h ^= h>>15; h *= 0x846ca68b; h ^= h>>16;
 *   start = h % 21474836
 * (uint32 arithmetic) and runs through its voxels in offset order exactly as the reference does inside a chunk. */
int flipb200_fluid_reseed(flipb200_world* w, uint32_t seed);
/* VDBPointsToPrimitive (projects/zenvdb/GetVDBPoints.cpp:76-258; SURVEY 8f-3, the viewport / export path): world position
 * float((double(P) + double(voxel)) * dx) and decoded velocity of every particle, [N][3] floats each (N from flipb200_particles_info),
 * store order (leaf, voxel, index; the reference walks leaves in tree order -- a primitive is an unordered point set). vel may be
 * NULL. The caller's arrays should be page-locked (flipb200_host_alloc). No OpenVDB tree is built on the way. */
int flipb200_particles_to_points(flipb200_world* w, float* pos, float* vel);
/* ParticleEmitter (FF/nosys/ParticleEmitter.cpp:9-62 -> FLIP_vdb::emit_liquid, FF/FLIP_vdb.cpp:2222-2642), the branch WITHOUT a
 * VelocityVolume (:2488-2624): every particle leaf box one of whose 9^3 lattice corners samples the shape SDF (float grid slot
 * `shapeGrid`, e.g. FLIPB200_KILLER_SDF) < 0 is filled -- per voxel whose centre samples < dx, up to 16 jittered candidates while
 * the voxel holds < 8, skipped when their octant is taken, taken when the shape at the candidate is < -0.1 dx; new particles get
 * the velocity (vx, vy, vz). Missing leaves are created; leaves the shape does not touch are left alone. The shape grid must share
 * the world's cell-centred transform (voxel size dx). Jitter and seeding as in flipb200_fluid_reseed. */
int flipb200_emit_liquid(flipb200_world* w, int shapeGrid, float vx, float vy, float vz, uint32_t seed);
/* FLIPApplyBoundary (FF/nosys/Update_Solid_SDF.cpp:9-49 -> FLIP_vdb::update_solid_sdf, FF/FLIP_vdb.cpp:1976-2046) with one moving
 * solid (float grid slot `movingGrid`, on the world's cell-centred transform or, movingVertexCentred != 0, on the static SDF's
 * vertex-centred one): FLIPB200_SOLID_SDF gains the leaf under every moving-solid leaf origin and every particle leaf with a box
 * corner inside the moving solid; then every voxel of every leaf = min(own value, moving solid sampled at the voxel's world
 * position) and is active. (Particle leaves that hold no particles are not considered: the device store does not keep them.) */
int flipb200_apply_boundary(flipb200_world* w, int movingGrid, int movingVertexCentred);
/* debug: keep / fetch the fp32 position (index space) and velocity before the codecs, in the
 * order of the particle store the advect call started from (codec caveat). */
int flipb200_capture_precodec(flipb200_world* w, int on);
int flipb200_get_precodec(flipb200_world* w, float* pos, float* vel, uint8_t* alive);

/* CutCellWeightEval::apply (FF/nosys/EvalFaceWeight.cpp:17-24) -> calculate_face_weights (FF/FLIP_vdb.cpp:2644-2718) */
int flipb200_face_weights(flipb200_world* w);
/* PushOutLiquidSDF::apply (FF/nosys/FixLiquidSDF.cpp:16-29) -> immerse_liquid_phi_in_solids (FF/FLIP_vdb.cpp:2720-2803) */
int flipb200_pushout_sdf(flipb200_world* w, float dx);
/* FieldAddVector::apply (FF/nosys/FieldAddVector.cpp:16-31) -> field_add_vector (FF/FLIP_vdb.cpp:3145-3158), dt = 1 */
int flipb200_add_vector(flipb200_world* w, float x, float y, float z);
/* CFL_dt (FF/nosys/CFL.cpp) -> FLIP_vdb::cfl (FF/FLIP_vdb.cpp:3160-3207) */
int flipb200_cfl(flipb200_world* w, float* dtOut);

/* AssembleSolvePPE::apply (FF/nosys/SolvePoissonPressureEqn.cpp:23-64) -> solve_pressure_simd_uaamg
 * (FF/FLIP_vdb.cpp:3034-3100): builds the variational Laplacian + multigrid hierarchy
 * (FF/simd_vdb_poisson_uaamg.cpp:357-747,1962-1991), the RHS (:18-93) and runs solveMultigridPCG
 * (:2332-2403; relative L-inf tolerance 5e-5, <=100 iterations, RBGS smoother), falling back to
 * solvePureMultigrid on failure. out: Pressure (new grid on the DOF mask), Divergence (= RHS).
 * status: 0 = PCG converged, 1 = fell back to pure multigrid. */
int flipb200_solve_ppe(flipb200_world* w, float dt, float dx, int* iterations, float* relResidual, int* status);
/* same, with the tolerance / iteration cap exposed (BASELINE config C4 uses 1e-6) */
int flipb200_solve_ppe_ex(flipb200_world* w, float dt, float dx, float relTol, int maxIter,
                          int* iterations, float* relResidual, int* status);
int flipb200_solver_info(flipb200_world* w, int* levels, int* numDof, int* nHistory);
int flipb200_residual_history(flipb200_world* w, float* out);

/* SubtractPressureGradient::apply (FF/nosys/SubtractPressureGradient.cpp:25-66) ->
 * apply_pressure_gradient (FF/FLIP_vdb.cpp:2863-2967) + union_extrapolate */
int flipb200_subtract_grad(flipb200_world* w, float dt, float dx, int velExtraLayer);
/* The Density / SurfaceTension sockets of AssembleSolvePPE and SubtractPressureGradient (FF/nosys/SolvePoissonPressureEqn.cpp:43-45,
 * SubtractPressureGradient.cpp:21-23): state of the world, read by the two calls above. tensionCoef > 0 enables the reference's tension
 * terms with tension = 2 tensionCoef / density -- BuildPoissonRhs_withTension (FF/simd_vdb_poisson_uaamg.cpp:95-209: a face towards an
 * air cell adds dt/dx^2 * weight * tension * (theta curvOther + (1 - theta) curvThis) / theta to the right-hand side) and the ghost
 * pressure of the gradient (FF/FLIP_vdb.cpp:2932-2939). The curvature is grid slot FLIPB200_CURVATURE (read by voxel coordinate, any
 * leaf set; background where absent, as the reference's accessor). Default: density 1000, tensionCoef 0 (off). */
int flipb200_set_surface_tension(flipb200_world* w, float density, float tensionCoef);

/* One substep of the packaged chain, device resident (no host copies in between):
 * G2PAdvectorSheetty -> FLIP_P2G -> CutCellWeight -> PushOutLiquidSDF -> FieldAddVector(g*dt)
 * -> AssembleSolvePPE -> SubtractPressureGradient (projects/tools/FLIPtools/stub.cpp:5-17).
 * stageMs (optional, 5 floats): device time of advect, p2g, small stencils, ppe, gradient. */
int flipb200_substep(flipb200_world* w, float dt, float dx, int surfaceSize, int rkOrder, float picMin,
                     float picMax, float gx, float gy, float gz, int velExtraLayer, int flags,
                     float* stageMs);

/* ---- measurement hooks used by bench.py (device timing on the world's stream) ---- */
/* number of kernels this library launched since the world was created */
int flipb200_launch_count(flipb200_world* w, uint64_t* n);
/* host waits on the world's stream so far (every one is an idle gap on the device) */
int flipb200_sync_count(flipb200_world* w, uint64_t* n);
/* per-kernel-family accumulated device time and launch counts since the last reset;
 * names is a ';'-separated list written into buf */
int flipb200_profile_enable(flipb200_world* w, int on);
int flipb200_profile_reset(flipb200_world* w);
int flipb200_profile_get(flipb200_world* w, char* names, size_t namesCap, float* ms, uint64_t* launches,
                         uint64_t* bytes, int cap, int* nOut);
/* the CUDA stream all work of this world is issued on (cudaStream_t as void*) */
int flipb200_stream(flipb200_world* w, void** stream);

/* ---- multi-GPU (one process per GPU; slab decomposition along x) ---- */
/* 128-byte NCCL unique id created on rank 0 and distributed by the host (torch.distributed / MPI) */
int flipb200_comm_unique_id(uint8_t id[128]);
int flipb200_comm_init(flipb200_world* w, int rank, int nRanks, const uint8_t id[128]);
/* in-process communicator: rank r = worlds[r], all in this process, each driven by its own host thread. Carries the
 * same message sequence as the NCCL one; lets the decomposition run (and be tested) on a single GPU. */
int flipb200_comm_init_local(flipb200_world** worlds, int n);
/* wake the ranks of an in-process communicator that are blocked in a collective after a peer failed */
int flipb200_comm_abort(flipb200_world* w);
/* Slab decomposition (no reference counterpart: the reference runs one TBB process).
 * This rank owns the leaf layers [leafLo, leafHi) along x (leaf coordinate = voxel >> 3); the first rank's lower and
 * the last rank's upper bound are open. Slabs must be at least two layers thick and cover the axis without gaps.
 * Every node of the packaged substep runs in this mode: VDBRenormalizeSDF / VDBSmoothSDF refresh the ghost layers between the
 * passes that read across a slab face; KillParticlesInSDF / ParticleAddDV / VDBErodeSDF need no exchange; FluidReseed and
 * ParticleEmitter decide per leaf from (seed, leaf origin) and data within two voxels of the leaf, so owner and ghost holder
 * agree without an exchange (pass the same seed and the same shape grid on every rank); FLIPApplyBoundary and the Curvature
 * socket read caller-supplied grids by voxel coordinate (give every rank the grid over its owned + ghost layers). Not
 * available: the pure-multigrid fallback after a failed PCG (the status is returned).
 * Ghost leaves travel over peer memory (cudaIpc-mapped receive boxes, device-side flags) when the GPUs can map each other,
 * else over NCCL send/recv; FLIPB200_DD_P2P=0 forces the latter.
 * From here on every node call is COLLECTIVE (all ranks issue the same sequence): flipb200_bin_from_points routes
 * points to their owners and fills the ghost layers, flipb200_g2p_advect_sheetty migrates particles,
 * P2G / solve / gradient exchange ghost leaves, CFL and the PCG scalars are all-reduced. Grids and particles
 * downloaded from a rank include its ghost layers; flipb200_dd_owned tells which leaves are authoritative. */
int flipb200_dd_set_slab(flipb200_world* w, int leafLo, int leafHi);
int flipb200_dd_owned(flipb200_world* w, int* leafLo, int* leafHi);
/* particles in the owned leaves (flipb200_particles_info counts the ghost copies as well); collective */
int flipb200_dd_owned_particles(flipb200_world* w, uint64_t* n);

#ifdef __cplusplus
}
#endif
#endif
