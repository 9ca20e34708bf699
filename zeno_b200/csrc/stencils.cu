// libflipb200 -- the small single-pass stencil stages that keep a substep device resident (K14):
// face weights, liquid-SDF push-out, body force, CFL, pressure-gradient subtraction.
#include "world.cuh"
#include "next_kernels.cuh"
#include "levelset.cuh"

namespace fb {
void finish_vec3(World* w, GridV& vel, const uint64_t* chMask);

namespace {

struct SolidAccess {
    TopoView pool; const float* view;
    TopoView st; const float* sval; float bg;
};
__device__ __forceinline__ float solid_get(const SolidAccess& s, int x, int y, int z) {
    int l = topo_find(s.pool, x, y, z);
    if (l >= 0) return __ldg(&s.view[(size_t)l * LEAF + voxel_off(x, y, z)]);
    if (s.st.n > 0) return grid_get(s.st, s.sval, s.bg, x, y, z);
    return s.bg;
}
__device__ __forceinline__ float ip64(float a, float b, double w) {
    return __fadd_rn(a, __double2float_rn(__dmul_rn((double)__fsub_rn(b, a), w)));
}
// BoxSampler(solid, ijk + 0.5): base = ijk, uvw = 0.5 (openvdb/tools/Interpolation.h:712-737)
__device__ __forceinline__ float solid_center_sample(const SolidAccess& s, int x, int y, int z) {
    float d[8];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) d[i * 4 + j * 2 + k] = solid_get(s, x + i, y + j, z + k);
    return ip64(ip64(ip64(d[0], d[1], 0.5), ip64(d[2], d[3], 0.5), 0.5), ip64(ip64(d[4], d[5], 0.5), ip64(d[6], d[7], 0.5), 0.5), 0.5);
}

// calculate_face_weights (FF/FLIP_vdb.cpp:2644-2718)
__global__ void __launch_bounds__(512) face_weight_kernel(SolidAccess s, const uint64_t* __restrict__ mask,
                                                          float* __restrict__ w0, float* __restrict__ w1, float* __restrict__ w2) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!mask_get(mask, leaf, off)) return;
    int3 o = s.pool.origin[leaf];
    int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
    float q[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) q[i][j][k] = solid_get(s, x + i, y + j, z + k);
    float u = __fsub_rn(1.0f, fraction_inside4(q[0][0][0], q[0][1][0], q[0][0][1], q[0][1][1]));
    u = fmaxf(0.f, fminf(u, 1.f));
    float v = __fsub_rn(1.0f, fraction_inside4(q[0][0][0], q[0][0][1], q[1][0][0], q[1][0][1]));
    v = fmaxf(0.f, fminf(v, 1.f));
    float ww = __fsub_rn(1.0f, fraction_inside4(q[0][0][0], q[1][0][0], q[0][1][0], q[1][1][0]));
    ww = fmaxf(0.f, fminf(ww, 1.f));
    size_t i = (size_t)leaf * LEAF + off;
    w0[i] = u; w1[i] = v; w2[i] = ww;
}

// immerse_liquid_phi_in_solids, first pass (FF/FLIP_vdb.cpp:2722-2736)
__global__ void __launch_bounds__(512) pushout_pass1_kernel(SolidAccess s, const uint8_t* __restrict__ solidLeaf,
                                                            const uint8_t* __restrict__ alloc,
                                                            const uint64_t* __restrict__ mask, float* __restrict__ phi, float dx) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!solidLeaf[leaf] || !alloc[leaf] || !mask_get(mask, leaf, off)) return;
    int3 o = s.pool.origin[leaf];
    float vs = solid_center_sample(s, o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
    size_t i = (size_t)leaf * LEAF + off;
    if (vs < 0.f) phi[i] = __fsub_rn(phi[i], __fmul_rn(0.5f, dx));
}
// second pass (:2741-2796): only leaves that existed before the dilation are visited
__global__ void __launch_bounds__(512) pushout_pass2_kernel(SolidAccess s, const uint8_t* __restrict__ solidLeaf,
                                                            const uint8_t* __restrict__ alloc,
                                                            const uint64_t* __restrict__ dilMask, const float* __restrict__ ref,
                                                            float* __restrict__ phi, float dx, float bg) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!solidLeaf[leaf] || !alloc[leaf] || !mask_get(dilMask, leaf, off)) return;
    int3 o = s.pool.origin[leaf];
    int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
    float vs = solid_center_sample(s, x, y, z);
    if (!(vs < 0.f)) return;
    bool found = false;
    float minFluid = __fmul_rn(dx, 3.0f);
    for (int i = 0; i < 6 && !found; i++) {
        int comp = i >> 1;
        int d = (i & 1) == 0 ? 1 : -1;
        float rv = grid_get(s.pool, ref, bg, x + (comp == 0 ? d : 0), y + (comp == 1 ? d : 0), z + (comp == 2 ? d : 0));
        minFluid = fminf(minFluid, rv);
        found |= (rv < 0.f);
    }
    size_t k = (size_t)leaf * LEAF + off;
    if (found) phi[k] = minFluid;
    else if (phi[k] < 0.f) phi[k] = fmaxf(bg, -vs);
}
__global__ void leaf_any_kernel(const uint64_t* __restrict__ mask, int n, uint8_t* __restrict__ alloc) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    uint64_t a = 0;
    for (int k = 0; k < 8; k++) a |= mask[(size_t)l * 8 + k];
    if (a) alloc[l] = 1;
}

// field_add_vector (FF/FLIP_vdb.cpp:3145-3158) on every voxel of the union mask
__global__ void __launch_bounds__(512) add_vector_kernel(const uint64_t* __restrict__ mask, float* __restrict__ v0,
                                                         float* __restrict__ v1, float* __restrict__ v2, float fx, float fy, float fz) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!mask_get(mask, leaf, off)) return;
    size_t i = (size_t)leaf * LEAF + off;
    v0[i] = __fadd_rn(v0[i], __fmul_rn(fx, 1.0f));
    v1[i] = __fadd_rn(v1[i], __fmul_rn(fy, 1.0f));
    v2[i] = __fadd_rn(v2[i], __fmul_rn(fz, 1.0f));
}
__global__ void __launch_bounds__(512) absmax3_kernel(const uint64_t* __restrict__ mask, const float* __restrict__ v0,
                                                      const float* __restrict__ v1, const float* __restrict__ v2,
                                                      unsigned* __restrict__ out) {
    int leaf = blockIdx.x, off = threadIdx.x;
    float m = 0.f;
    if (mask_get(mask, leaf, off)) {
        size_t i = (size_t)leaf * LEAF + off;
        m = fmaxf(fabsf(v0[i]), fmaxf(fabsf(v1[i]), fabsf(v2[i])));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// apply_pressure_gradient (FF/FLIP_vdb.cpp:2863-2967), one channel; tension disabled
struct GradTension { int on; float tension; TopoView ct; const float* curv; float curvBg; };
__global__ void __launch_bounds__(512) pressure_gradient_kernel(TopoView t, int ch, const uint64_t* __restrict__ velMask,
                                                                float* __restrict__ vel, const float* __restrict__ fw,
                                                                const float* __restrict__ phi, float phiBg,
                                                                const float* __restrict__ prs, const uint64_t* __restrict__ prsMask,
                                                                const float* __restrict__ svel, uint64_t* __restrict__ chMaskOut,
                                                                float dt, float dx, GradTension T) {
    int leaf = blockIdx.x, off = threadIdx.x;
    bool on = mask_get(velMask, leaf, off);
    bool keep = false;
    if (on) {
        int3 o = t.origin[leaf];
        int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
        size_t i = (size_t)leaf * LEAF + off;
        float faceW = fw[i];
        int lx = x - (ch == 0), ly = y - (ch == 1), lz = z - (ch == 2);
        if (faceW > 0.0f) {
            int ll = topo_find(t, lx, ly, lz);
            int lo = voxel_off(lx, ly, lz);
            bool hp = mask_get(prsMask, leaf, off);
            bool hpb = ll >= 0 && mask_get(prsMask, ll, lo);
            if (hp || hpb) {
                keep = true;
                float phiThis = phi[i];
                float phiBelow = ll >= 0 ? phi[(size_t)ll * LEAF + lo] : phiBg;
                float pThis = prs[i];
                float pBelow = ll >= 0 ? prs[(size_t)ll * LEAF + lo] : 0.f;
                float theta = 1.0f;
                if (phiThis >= 0.f || phiBelow >= 0.f) {
                    theta = fraction_inside2(phiBelow, phiThis);
                    if (theta < 0.02f) theta = 0.02f;
                    if (T.on) {   // FF/FLIP_vdb.cpp:2932-2939: the air cell's pressure is the tension jump
                        const float curvThis = grid_get(T.ct, T.curv, T.curvBg, x, y, z), curvBelow = grid_get(T.ct, T.curv, T.curvBg, lx, ly, lz);
                        if (phiThis >= 0.f) pThis = __fmul_rn(T.tension, __fadd_rn(__fmul_rn(theta, curvThis), __fmul_rn(__fsub_rn(1.f, theta), curvBelow)));
                        else if (phiBelow >= 0.f) pBelow = __fmul_rn(T.tension, __fadd_rn(__fmul_rn(theta, curvBelow), __fmul_rn(__fsub_rn(1.f, theta), curvThis)));
                    }
                }
                float velUpdate = __fdiv_rn(__fdiv_rn(__fmul_rn(-dt, __fsub_rn(pThis, pBelow)), dx), theta);
                float updated = __fadd_rn(vel[i], velUpdate);
                if (faceW < 1.0f) {
                    const float friction = 0.f;
                    float solidFraction = __fmul_rn(__fsub_rn(1.0f, faceW), friction);
                    updated = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, solidFraction), updated), __fmul_rn(solidFraction, svel[i]));
                }
                vel[i] = updated;
            }
        }
    }
    unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<uint32_t*>(chMaskOut)[(size_t)leaf * 16 + (threadIdx.x >> 5)] = b;
}
SolidAccess solid_access(World* w) {
    SolidAccess s;
    s.pool = w->pool->view();
    s.view = w->solidSdfView.p;
    GridF& g = w->F(FLIPB200_SOLID_SDF);
    s.st = (w->hasSolidSDF && g.topo) ? g.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
    s.sval = g.val.p;
    s.bg = g.bg;
    return s;
}
}  // namespace

void mark_alloc_from_mask(World* w, const uint64_t* mask, int n, uint8_t* alloc) {
    if (!n) return;
    FB_LAUNCH(w, "leaf_any", (size_t)n * 65) leaf_any_kernel<<<(n + 127) / 128, 128, 0, w->stream>>>(mask, n, alloc);
    check_launch("leaf_any");
}

void face_weights(World* w) {
    ensure_pool(w, {FLIPB200_LIQUID_SDF}, false);
    refresh_solid_views(w);
    TopoPtr pool = w->pool;
    GridF& phi = w->F(FLIPB200_LIQUID_SDF);
    GridV nf;
    const float one[3] = {1.f, 1.f, 1.f};
    grid_alloc(w, nf, pool, one);
    if (pool->n) {
        mask_dilate(w, *pool, phi.mask.p, nf.mask.p, true);
        FB_LAUNCH(w, "face_weights", (size_t)pool->n * LEAF * 16) face_weight_kernel<<<pool->n, 512, 0, w->stream>>>(solid_access(w), nf.mask.p, nf.val[0].p, nf.val[1].p, nf.val[2].p);
        check_launch("face_weights");
    }
    w->V(FLIPB200_FACE_WEIGHT) = std::move(nf);
}

void pushout_sdf(World* w, float dx) {
    ensure_pool(w, {FLIPB200_LIQUID_SDF}, false);
    refresh_solid_views(w);
    TopoPtr pool = w->pool;
    GridF& phi = w->F(FLIPB200_LIQUID_SDF);
    const int n = pool->n;
    if (!n) return;
    // phi.alloc: which pool leaves are leaves of the reference's phi tree (DESIGN.md "leaf existence")
    mark_alloc_from_mask(w, phi.mask.p, n, phi.alloc.p);
    uint8_t* allocp = phi.alloc.p;
    SolidAccess s = solid_access(w);
    FB_LAUNCH(w, "pushout_pass1", (size_t)n * LEAF * 8) pushout_pass1_kernel<<<n, 512, 0, w->stream>>>(s, w->solidLeafExists.p, allocp, phi.mask.p, phi.val.p, dx);
    check_launch("pushout1");
    DBuf<float> ref((size_t)n * LEAF, w->stream);
    FB_CUDA(cudaMemcpyAsync(ref.p, phi.val.p, (size_t)n * LEAF * 4, cudaMemcpyDeviceToDevice, w->stream));
    DBuf<uint64_t> dil((size_t)n * 8, w->stream);
    mask_dilate(w, *pool, phi.mask.p, dil.p, true);
    FB_LAUNCH(w, "pushout_pass2", (size_t)n * LEAF * 12) pushout_pass2_kernel<<<n, 512, 0, w->stream>>>(s, w->solidLeafExists.p, allocp, dil.p, ref.p, phi.val.p, dx, phi.bg);
    check_launch("pushout2");
    mask_dilate(w, *pool, dil.p, phi.mask.p, false);
    // the tree now also owns every leaf touched by the two dilations
    mark_alloc_from_mask(w, phi.mask.p, n, allocp);
}

void add_vector(World* w, float x, float y, float z) {
    GridV& v = w->V(FLIPB200_VELOCITY);
    if (!v.topo || v.topo->n == 0) return;
    int n = v.topo->n;
    FB_LAUNCH(w, "add_vector", (size_t)n * LEAF * 24) add_vector_kernel<<<n, 512, 0, w->stream>>>(v.mask.p, v.val[0].p, v.val[1].p, v.val[2].p, x, y, z);
    check_launch("add_vector");
}

float cfl(World* w) {
    GridV& v = w->V(FLIPB200_VELOCITY);
    const bool empty = !v.topo || v.topo->n == 0;
    if (empty && !dd_on(w)) return 3.402823466e+38f / 2;
    int n = empty ? 0 : v.topo->n;
    DBuf<unsigned> m(1, w->stream);
    m.zero();
    if (n) {
        FB_LAUNCH(w, "cfl_absmax", (size_t)n * LEAF * 12) absmax3_kernel<<<n, 512, 0, w->stream>>>(v.mask.p, v.val[0].p, v.val[1].p, v.val[2].p, m.p);
        check_launch("absmax3");
    }
    if (dd_on(w)) comm_allreduce(w, m.p, 1, CT_U32, true);   // bit patterns of non-negative floats order like the floats
    unsigned h = 0;
    read_back(w, &h, m.p, 4);
    float mv;
    memcpy(&mv, &h, 4);
    // see oracle/stencils.cpp node_CFL_dt: the value the reference reads is the global maximum
    return w->dx / (fabsf(mv) + 1e-6f);
}

// ---------------------------------------------------------------- VDBRenormalizeSDF / VDBErodeSDF (bodies: next_kernels.cuh)
namespace {
__global__ void __launch_bounds__(512) renorm_stage_kernel(TopoView t, const uint64_t* __restrict__ mask, const float* __restrict__ cur,
                                                           const float* __restrict__ phi0, float* __restrict__ out, float bg, float dt,
                                                           float invDx, float alpha, float beta, int useAlpha) {
    nextk::renorm_stage_one(t, mask, cur, phi0, out, bg, dt, invDx, alpha, beta, useAlpha, blockIdx.x, threadIdx.x);
}
__global__ void __launch_bounds__(512) box_avg_kernel(TopoView t, const uint64_t* __restrict__ mask, const float* __restrict__ cur,
                                                      float* __restrict__ out, float bg, int axis, int w, float frac) {
    nextk::box_avg_one(t, mask, cur, out, bg, axis, w, frac, blockIdx.x, threadIdx.x);
}
__global__ void __launch_bounds__(512) add_active_kernel(const uint64_t* __restrict__ mask, float* __restrict__ val, float d) {
    nextk::add_active_one(mask, val, blockIdx.x, threadIdx.x, d);
}
}  // namespace
// VDBSmoothSDF::apply (projects/zenvdb/VDBRenormalize.cpp:108-120) = openvdb::tools::Filter::gaussian(width, iterations), tiles off
// (tools/Filter.h:574-630): per iteration four box filters, each as three one-dimensional passes in the order X, Z, Y
void smooth_sdf(World* w, int grid, int width, int iterations) {
    FB_REQUIRE(is_float_grid(grid) && w->F(grid).topo != nullptr, FLIPB200_ERR_STATE, "VDBSmoothSDF: the grid does not exist");
    GridF& g = w->F(grid);
    // slab decomposition: an x pass reads `width` voxels across a slab face; the two ghost leaf layers are refreshed after every x pass
    const bool dd = dd_on(w);
    FB_REQUIRE(!dd || (g.topo == w->pool && width <= 8), FLIPB200_ERR_STATE, "VDBSmoothSDF under slab decomposition: the grid must live on the pool (e.g. LiquidSDF) and width <= 8");
    const int n = g.topo->n;
    if (!n || iterations <= 0) return;
    const int wd = width < 1 ? 1 : width;
    const float frac = 1.f / (float)(2 * wd + 1);
    DBuf<float> a((size_t)n * LEAF, w->stream);
    const TopoView t = g.topo->view();
    static const int order[3] = {0, 2, 1};
    for (int it = 0; it < iterations; it++)
        for (int rep = 0; rep < 4; rep++)
            for (int k = 0; k < 3; k++) {
                FB_LAUNCH(w, "box_avg", (size_t)n * LEAF * 8) box_avg_kernel<<<n, 512, 0, w->stream>>>(t, g.mask.p, g.val.p, a.p, g.bg, order[k], wd, frac);
                check_launch("box_avg");
                std::swap(g.val, a);
                if (dd && order[k] == 0) dd_refresh(w, {DDArray{g.val.p, LEAF * 4}}, 2);
            }
}

// VDBErodeSDF::apply (projects/zenvdb/VDBRenormalize.cpp:155-172): every active voxel += depth
void erode_sdf(World* w, int grid, float depth) {
    FB_REQUIRE(is_float_grid(grid) && w->F(grid).topo != nullptr, FLIPB200_ERR_STATE, "VDBErodeSDF: the grid does not exist");
    GridF& g = w->F(grid);
    const int n = g.topo->n;
    if (!n) return;
    FB_LAUNCH(w, "erode_sdf", (size_t)n * LEAF * 8) add_active_kernel<<<n, 512, 0, w->stream>>>(g.mask.p, g.val.p, depth);
    check_launch("erode_sdf");
}

// VDBRenormalizeSDF::apply (projects/zenvdb/VDBRenormalize.cpp:18-37): LevelSetTracker {FIRST_BIAS, TVD_RK3, 1, 1}, no trimming,
// `iterations` x normalize(); each normalize = three Euler stages (Normalizer::normalize, LevelSetTracker.h:535-604)
void renormalize_sdf(World* w, int grid, int iterations) {
    FB_REQUIRE(is_float_grid(grid) && w->F(grid).topo != nullptr, FLIPB200_ERR_STATE, "VDBRenormalizeSDF: the grid does not exist");
    GridF& g = w->F(grid);
    // slab decomposition: the three stages of an iteration reach three voxels into the ghost layer; it is refreshed once per iteration
    const bool dd = dd_on(w);
    FB_REQUIRE(!dd || g.topo == w->pool, FLIPB200_ERR_STATE, "VDBRenormalizeSDF under slab decomposition: the grid must live on the pool (e.g. LiquidSDF)");
    const int n = g.topo->n;
    if (!n || iterations <= 0) return;
    const size_t nv = (size_t)n * LEAF;
    const float h = w->dx, dt = h * 1.0f, invDx = 1.0f / h;
    DBuf<float> phi0(nv, w->stream), a(nv, w->stream);
    const TopoView t = g.topo->view();
    auto stage = [&](const float* cur, float* out, int N, int D) {
        const float alpha = D ? (float)N / (float)D : 0.f;
        FB_LAUNCH(w, "renorm_stage", nv * 12) renorm_stage_kernel<<<n, 512, 0, w->stream>>>(t, g.mask.p, cur, phi0.p, out, g.bg, dt, invDx, alpha, 1.0f - alpha, N ? 1 : 0);
        check_launch("renorm_stage");
    };
    for (int it = 0; it < iterations; it++) {
        FB_CUDA(cudaMemcpyAsync(phi0.p, g.val.p, nv * 4, cudaMemcpyDeviceToDevice, w->stream));
        stage(g.val.p, a.p, 0, 1);      // Phi_t1
        stage(a.p, g.val.p, 3, 4);      // Phi_t2
        stage(g.val.p, a.p, 1, 3);      // Phi_t3
        std::swap(g.val, a);
        if (dd) dd_refresh(w, {DDArray{g.val.p, LEAF * 4}}, 2);
    }
}

void subtract_grad(World* w, float dt, float dx, int velExtraLayer) {
    ensure_pool(w, {FLIPB200_VELOCITY, FLIPB200_LIQUID_SDF, FLIPB200_PRESSURE, FLIPB200_FACE_WEIGHT}, false);
    refresh_solid_views(w);
    TopoPtr pool = w->pool;
    const int n = pool->n;
    if (!n) { if (dd_on(w)) dd_refresh(w, w->V(FLIPB200_VELOCITY), 2); return; }
    GridV& vel = w->V(FLIPB200_VELOCITY);
    GridV& fw = w->V(FLIPB200_FACE_WEIGHT);
    GridF& phi = w->F(FLIPB200_LIQUID_SDF);
    GridF& prs = w->F(FLIPB200_PRESSURE);
    DBuf<uint64_t> chMask((size_t)3 * n * 8, w->stream);
    GradTension T;
    memset(&T, 0, sizeof(T));
    if (w->tensionCoef > 0.f) {   // enable_tension (FF/nosys/SubtractPressureGradient.cpp:23); tension = 2 coef / density (FF/FLIP_vdb.cpp:2875)
        GridF& cv = w->F(FLIPB200_CURVATURE);
        T.on = 1; T.tension = 2 * w->tensionCoef / w->density;
        if (cv.topo) { T.ct = cv.topo->view(); T.curv = cv.val.p; }
        T.curvBg = cv.bg;
    }
    for (int ch = 0; ch < 3; ch++) {
        FB_LAUNCH(w, "pressure_gradient", (size_t)n * LEAF * 20)
            pressure_gradient_kernel<<<n, 512, 0, w->stream>>>(pool->view(), ch, vel.mask.p, vel.val[ch].p, fw.val[ch].p, phi.val.p, phi.bg,
                                                               prs.val.p, prs.mask.p, w->solidVelView[ch].p, chMask.p + (size_t)ch * n * 8, dt, dx, T);
        check_launch("pressure_gradient");
    }
    union_extrapolate(w, velExtraLayer, vel, chMask.p, phi.mask.p);
    finish_vec3(w, vel, chMask.p);
    if (dd_on(w)) dd_refresh(w, vel, 2);
}

}  // namespace fb
