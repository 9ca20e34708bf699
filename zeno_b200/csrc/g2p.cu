// libflipb200 -- grid-to-particle transfer + advection + re-binning (K2, K7, K8).
// Follows point_to_counter_reducer2::operator() (FF/FLIP_vdb.cpp:513-754) per particle:
// fp32 staggered samples of the new/old velocity, FLIP/PIC blend from the liquid-SDF depth and
// the solid proximity, RK1..4 advection through OpenVDB's double-weight sampler, solid push-out
// with the least-squares corner normal (:3270-3367, evaluated on the fly for the few particles
// that end inside the solid), codec write-back, then the stable counting sort of particles.cu.
// Compiled with -fmad=false: the op sequence is the oracle's (no contraction).
#include "world.cuh"

namespace fb {
void origins_from_ijk(World* w, const int3* ijk, uint64_t n, int3* origins);
void keys_from_ijk(World* w, const TopoPtr& t, const int3* ijk, const uint8_t* alive, uint64_t n, uint32_t* keys);

namespace {
constexpr int G2P_THREADS = 256;

struct G2PParams {
    TopoView t;                       // pool
    const uint32_t* voxelStart;
    uint32_t *w0, *w1, *w2;           // in/out in place
    const float *vel[3], *oldv[3], *carr[3];
    const float* lsdf; float lsdfBg; int hasLiquid;
    const float* solidView; float solidBg; int hasSolid;
    TopoView st; const float* solidStatic;     // static solid grid, for reads outside the pool
    const float* svelView[3]; int hasSolidVel;
    const uint64_t* nmask;            // dilate5(liquid sdf topology)
    float dx, dt, picMin, picMax, surfacedist;
    int rkOrder, sameField;
    int3* ijkOut; uint8_t* alive;
    int leaf0;                        // first leaf of the launch (slab decomposition: owned leaves only)
    float *prePos, *preVel; uint8_t* preAlive;
};

// The particle's own leaf: most samples fall into it, and then the slot is known without the (dependent) directory load.
#ifndef FB_G2P_HOME
#define FB_G2P_HOME 0   // measured on B200: the shortcut branch costs more (4.1 ms) than the directory load it saves (2.4 ms)
#endif
struct Home { int leaf; int ox, oy, oz; };
__device__ __forceinline__ int find_leaf(const TopoView& t, const Home& h, int x, int y, int z) {
#if FB_G2P_HOME
    if ((((x ^ h.ox) | (y ^ h.oy) | (z ^ h.oz)) & ~7) == 0) return h.leaf;
#endif
    return topo_find(t, x, y, z);
}
__device__ __forceinline__ float grid_get_h(const TopoView& t, const Home& h, const float* val, float bg, int vx, int vy, int vz) {
    int l = find_leaf(t, h, vx, vy, vz);
    if (l < 0) return bg;
    return __ldg(&val[(size_t)l * LEAF + voxel_off(vx, vy, vz)]);
}
// eight corner values of the cell with base (bx,by,bz); index i*4+j*2+k
__device__ __forceinline__ void fetch8(const TopoView& t, const Home& h, const float* __restrict__ val, float bg, int bx, int by, int bz, float d[8]) {
    if (((bx & 7) != 7) && ((by & 7) != 7) && ((bz & 7) != 7)) {
        int l = find_leaf(t, h, bx, by, bz);
        if (l < 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) d[q] = bg;
            return;
        }
        const float* p = val + (size_t)l * LEAF + voxel_off(bx, by, bz);
        d[0] = __ldg(p); d[1] = __ldg(p + 1); d[2] = __ldg(p + 8); d[3] = __ldg(p + 9);
        d[4] = __ldg(p + 64); d[5] = __ldg(p + 65); d[6] = __ldg(p + 72); d[7] = __ldg(p + 73);
        return;
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) d[i * 4 + j * 2 + k] = grid_get_h(t, h, val, bg, bx + i, by + j, bz + k);
}
__device__ __forceinline__ float solid_get(const G2PParams& p, const Home& h, int x, int y, int z) {
    int l = find_leaf(p.t, h, x, y, z);
    if (l >= 0) return __ldg(&p.solidView[(size_t)l * LEAF + voxel_off(x, y, z)]);
    if (p.st.n > 0) return grid_get(p.st, p.solidStatic, p.solidBg, x, y, z);
    return p.solidBg;
}
__device__ __forceinline__ void fetch8_solid(const G2PParams& p, const Home& h, int bx, int by, int bz, float d[8]) {
    if (((bx & 7) != 7) && ((by & 7) != 7) && ((bz & 7) != 7)) {
        int l = find_leaf(p.t, h, bx, by, bz);
        if (l >= 0) {
            const float* q = p.solidView + (size_t)l * LEAF + voxel_off(bx, by, bz);
            d[0] = __ldg(q); d[1] = __ldg(q + 1); d[2] = __ldg(q + 8); d[3] = __ldg(q + 9);
            d[4] = __ldg(q + 64); d[5] = __ldg(q + 65); d[6] = __ldg(q + 72); d[7] = __ldg(q + 73);
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) d[i * 4 + j * 2 + k] = solid_get(p, h, bx + i, by + j, bz + k);
}
__device__ __forceinline__ float mixf(float a, float b, float w) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), w)); }
// local fp32 sampler (FF/FLIP_vdb.cpp:25-110)
__device__ __forceinline__ float samplec_f32(const TopoView& t, const Home& h, const float* val, float bg, float x, float y, float z) {
    int bx = (int)floor((double)x), by = (int)floor((double)y), bz = (int)floor((double)z);
    float d[8];
    fetch8(t, h, val, bg, bx, by, bz, d);
    float wx = __fsub_rn(x, (float)bx), wy = __fsub_rn(y, (float)by), wz = __fsub_rn(z, (float)bz);
    return mixf(mixf(mixf(d[0], d[1], wz), mixf(d[2], d[3], wz), wy), mixf(mixf(d[4], d[5], wz), mixf(d[6], d[7], wz), wy), wx);
}
__device__ __forceinline__ void staggered_f32(const TopoView& t, const Home& h, const float* const v[3], const float q[3], float out[3]) {
    out[0] = samplec_f32(t, h, v[0], 0.f, __fadd_rn(q[0], 0.5f), q[1], q[2]);
    out[1] = samplec_f32(t, h, v[1], 0.f, q[0], __fadd_rn(q[1], 0.5f), q[2]);
    out[2] = samplec_f32(t, h, v[2], 0.f, q[0], q[1], __fadd_rn(q[2], 0.5f));
}
// openvdb BoxSampler, double weights (openvdb/tools/Interpolation.h:712-737)
__device__ __forceinline__ float ip64(float a, float b, double w) {
    return __fadd_rn(a, __double2float_rn(__dmul_rn((double)__fsub_rn(b, a), w)));
}
__device__ __forceinline__ float tri64(const float d[8], double u, double v, double w) {
    return ip64(ip64(ip64(d[0], d[1], w), ip64(d[2], d[3], w), v), ip64(ip64(d[4], d[5], w), ip64(d[6], d[7], w), v), u);
}
__device__ __forceinline__ float box_f64(const TopoView& t, const Home& h, const float* val, float bg, double x, double y, double z) {
    int bx = (int)floor(x), by = (int)floor(y), bz = (int)floor(z);
    float d[8];
    fetch8(t, h, val, bg, bx, by, bz, d);
    return tri64(d, __dsub_rn(x, (double)bx), __dsub_rn(y, (double)by), __dsub_rn(z, (double)bz));
}
__device__ __forceinline__ float box_f64_solid(const G2PParams& p, const Home& h, double x, double y, double z) {
    int bx = (int)floor(x), by = (int)floor(y), bz = (int)floor(z);
    float d[8];
    fetch8_solid(p, h, bx, by, bz, d);
    return tri64(d, __dsub_rn(x, (double)bx), __dsub_rn(y, (double)by), __dsub_rn(z, (double)bz));
}
__device__ __forceinline__ void staggered_f64(const TopoView& t, const Home& h, const float* const v[3], const float q[3], float out[3]) {
    out[0] = box_f64(t, h, v[0], 0.f, __dadd_rn((double)q[0], 0.5), (double)q[1], (double)q[2]);
    out[1] = box_f64(t, h, v[1], 0.f, (double)q[0], __dadd_rn((double)q[1], 0.5), (double)q[2]);
    out[2] = box_f64(t, h, v[2], 0.f, (double)q[0], (double)q[1], __dadd_rn((double)q[2], 0.5));
}
// custom_integrator (FF/FLIP_vdb.cpp:169-214)
__device__ __forceinline__ void integrate(int order, const G2PParams& p, const Home& h, float dtinvx, float ipos[3], const float V0[3]) {
    float q[3], V1[3], V2[3], V3[3];
    if (order == 2) {
#pragma unroll
        for (int a = 0; a < 3; a++) q[a] = __fadd_rn(ipos[a], __fmul_rn(__fmul_rn(0.5f, V0[a]), dtinvx));
        staggered_f64(p.t, h, p.vel, q, V1);
#pragma unroll
        for (int a = 0; a < 3; a++) ipos[a] = __fadd_rn(ipos[a], __fmul_rn(V1[a], dtinvx));
    } else if (order == 3) {
#pragma unroll
        for (int a = 0; a < 3; a++) q[a] = __fadd_rn(ipos[a], __fmul_rn(__fmul_rn(0.5f, V0[a]), dtinvx));
        staggered_f64(p.t, h, p.vel, q, V1);
#pragma unroll
        for (int a = 0; a < 3; a++) q[a] = __fadd_rn(ipos[a], __fmul_rn(dtinvx, __fsub_rn(__fmul_rn(2.0f, V1[a]), V0[a])));
        staggered_f64(p.t, h, p.vel, q, V2);
#pragma unroll
        for (int a = 0; a < 3; a++)
            ipos[a] = __fadd_rn(ipos[a], __fmul_rn(__fmul_rn(dtinvx, __fadd_rn(__fadd_rn(V0[a], __fmul_rn(4.0f, V1[a])), V2[a])), (1.0f / 6.0f)));
    } else if (order == 4) {
#pragma unroll
        for (int a = 0; a < 3; a++) q[a] = __fadd_rn(ipos[a], __fmul_rn(__fmul_rn(0.5f, V0[a]), dtinvx));
        staggered_f64(p.t, h, p.vel, q, V1);
#pragma unroll
        for (int a = 0; a < 3; a++) q[a] = __fadd_rn(ipos[a], __fmul_rn(__fmul_rn(0.5f, V1[a]), dtinvx));
        staggered_f64(p.t, h, p.vel, q, V2);
#pragma unroll
        for (int a = 0; a < 3; a++) q[a] = __fadd_rn(ipos[a], __fmul_rn(V2[a], dtinvx));
        staggered_f64(p.t, h, p.vel, q, V3);
#pragma unroll
        for (int a = 0; a < 3; a++)
            ipos[a] = __fadd_rn(ipos[a], __fmul_rn(__fmul_rn(dtinvx, __fadd_rn(__fadd_rn(V0[a], __fmul_rn(2.0f, __fadd_rn(V1[a], V2[a]))), V3[a])), (1.0f / 6.0f)));
    } else {
#pragma unroll
        for (int a = 0; a < 3; a++) ipos[a] = __fadd_rn(ipos[a], __fmul_rn(V0[a], dtinvx));
    }
}
// K8 on the fly: normal / on-state of voxel q of the "solidnormal" grid (FF/FLIP_vdb.cpp:3270-3367)
__device__ bool solid_normal_at(const G2PParams& p, const Home& h, int qx, int qy, int qz, float n[3]) {
    n[0] = n[1] = n[2] = 0.f;
    int l = find_leaf(p.t, h, qx, qy, qz);
    if (l < 0 || !mask_get(p.nmask, l, voxel_off(qx, qy, qz))) return false;
    float data[8];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 2; c++) data[a * 4 + b * 2 + c] = solid_get(p, h, qx + a, qy + b, qz + c);
    const float dx = p.dx;
    const float s = __fdiv_rn(1.0f, __fmul_rn(dx, dx));
    const float invATA[4] = {__fmul_rn(0.5f, s), __fmul_rn(0.5f, s), __fmul_rn(0.5f, s), __fmul_rn(0.125f, s)};
    float abcd[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int a = q >> 2, b = (q >> 1) & 1, c = q & 1;
            float coef = r == 0 ? __fmul_rn((float)a - 0.5f, dx) : (r == 1 ? __fmul_rn((float)b - 0.5f, dx) : (r == 2 ? __fmul_rn((float)c - 0.5f, dx) : __fmul_rn(1.0f, dx)));
            acc = __fadd_rn(acc, __fmul_rn(coef, data[q]));
        }
        abcd[r] = __fmul_rn(invATA[r], acc);
    }
    if (!(abcd[3] < 0.5f)) return false;
    float len2 = __fadd_rn(__fadd_rn(__fmul_rn(abcd[0], abcd[0]), __fmul_rn(abcd[1], abcd[1])), __fmul_rn(abcd[2], abcd[2]));
    float d = __double2float_rn(sqrt((double)len2));
    n[0] = abcd[0]; n[1] = abcd[1]; n[2] = abcd[2];
    if (!(fabsf(d) <= 1.0e-7f)) {
        float inv = __fdiv_rn(1.0f, d);
        n[0] = __fmul_rn(n[0], inv); n[1] = __fmul_rn(n[1], inv); n[2] = __fmul_rn(n[2], inv);
    }
    return true;
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(G2P_THREADS, MIN_BLOCKS) g2p_advect_kernel(G2PParams p) {
    __shared__ uint32_t sStart[LEAF + 1];
    const int leaf = blockIdx.x + p.leaf0;
    const size_t vbase = (size_t)leaf * LEAF;
    const uint32_t leafBeg = __ldg(&p.voxelStart[vbase]), leafEnd = __ldg(&p.voxelStart[vbase + LEAF]);
    if (leafEnd == leafBeg) return;
    for (int i = threadIdx.x; i <= LEAF; i += G2P_THREADS) sStart[i] = __ldg(&p.voxelStart[vbase + i]);
    __syncthreads();
    const int3 o = p.t.origin[leaf];
    const Home h{leaf, o.x, o.y, o.z};
    const float dx = p.dx;
    const float deep_threshold = (float)(-4.0 * (double)dx);
    const float invdx = __fdiv_rn(1.0f, dx);
    const float dtinvx = __fdiv_rn(p.dt, dx);
    for (uint32_t gi = leafBeg + threadIdx.x; gi < leafEnd; gi += G2P_THREADS) {
        // voxel of this particle: largest off with sStart[off] <= gi
        int lo = 0, hi = LEAF;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sStart[mid] <= gi) lo = mid; else hi = mid; }
        const int off = lo;
        const int vx = o.x + (off >> 6), vy = o.y + ((off >> 3) & 7), vz = o.z + (off & 7);
        uint32_t a0 = p.w0[gi], a1 = p.w1[gi], a2 = p.w2[gi];
        float pIs[3] = {__fadd_rn((float)vx, fx_decode(a0 & 0xffffu)), __fadd_rn((float)vy, fx_decode(a0 >> 16)),
                        __fadd_rn((float)vz, fx_decode(a1 & 0xffffu))};
        float pvel[3] = {h_decode(a1 >> 16), h_decode(a2 & 0xffffu), h_decode(a2 >> 16)};
        float adv[3], old[3], carried[3];
        staggered_f32(p.t, h, p.vel, pIs, adv);
        staggered_f32(p.t, h, p.oldv, pIs, old);
        float flip = __fsub_rn(1.0f, p.picMin);
        float pls = p.hasLiquid ? box_f64(p.t, h, p.lsdf, p.lsdfBg, (double)pIs[0], (double)pIs[1], (double)pIs[2]) : p.lsdfBg;
        float t_coef = 1.f;
        if (pls < 0.f && pls >= -p.surfacedist) {
            t_coef = __fdiv_rn(pls, -p.surfacedist);
            t_coef = fminf(fmaxf(t_coef, 0.0f), 1.0f);
        }
        if (pls >= 0.f) t_coef = 0.f;
        if (p.surfacedist > 0.f)
            flip = __fadd_rn(__fmul_rn(t_coef, flip), __fmul_rn(__fsub_rn(1.0f, t_coef), fminf(__fsub_rn(1.0f, p.picMax), flip)));
        float pss = box_f64_solid(p, h, (double)__fadd_rn(pIs[0], 0.5f), (double)__fadd_rn(pIs[1], 0.5f), (double)__fadd_rn(pIs[2], 0.5f));
        if (pss >= 0.f && (double)pss <= 2.0 * (double)dx) {
            float scoef = __fdiv_rn(pss, __fmul_rn(2.0f, dx));
            flip = __fadd_rn(__fmul_rn(scoef, flip), __fmul_rn(__fsub_rn(1.0f, scoef), 1.0f));
        }
        if (p.sameField) { carried[0] = adv[0]; carried[1] = adv[1]; carried[2] = adv[2]; }
        else staggered_f32(p.t, h, p.carr, pIs, carried);
#pragma unroll
        for (int a = 0; a < 3; a++) pvel[a] = __fadd_rn(carried[a], __fmul_rn(flip, __fadd_rn(-old[a], pvel[a])));
        float pIt[3] = {pIs[0], pIs[1], pIs[2]};
        if (pls >= -p.surfacedist) integrate(1, p, h, dtinvx, pIt, adv);
        else integrate(p.rkOrder, p, h, dtinvx, pIt, adv);
        int pt[3];
#pragma unroll
        for (int a = 0; a < 3; a++) pt[a] = (int)floor((double)__fadd_rn(pIt[a], 0.5f));
        float nps = box_f64_solid(p, h, (double)__fadd_rn(pIt[0], 0.5f), (double)__fadd_rn(pIt[1], 0.5f), (double)__fadd_rn(pIt[2], 0.5f));
        bool dropped = false;
        if (nps < 0.f) {
            if (nps < deep_threshold) dropped = true;
            else {
                float sn[3];
                solid_normal_at(p, h, pt[0], pt[1], pt[2], sn);
#pragma unroll
                for (int a = 0; a < 3; a++) pIt[a] = __fsub_rn(pIt[a], __fmul_rn(__fmul_rn(__fmul_rn(nps, sn[a]), invdx), 1.0f));
#pragma unroll
                for (int a = 0; a < 3; a++) pt[a] = (int)floor((double)__fadd_rn(pIt[a], 0.5f));
                float vnv = 0.f;
                if (p.hasSolidVel) {
                    float n2[3];
                    if (solid_normal_at(p, h, pt[0], pt[1], pt[2], n2)) {
                        int l = find_leaf(p.t, h, pt[0], pt[1], pt[2]);  // on => inside the pool
                        size_t k = (size_t)l * LEAF + voxel_off(pt[0], pt[1], pt[2]);
                        vnv = __fadd_rn(__fadd_rn(__fmul_rn(p.svelView[0][k], n2[0]), __fmul_rn(p.svelView[1][k], n2[1])), __fmul_rn(p.svelView[2][k], n2[2]));
                    }
                }
                float dot = __fadd_rn(__fadd_rn(__fmul_rn(sn[0], pvel[0]), __fmul_rn(sn[1], pvel[1])), __fmul_rn(sn[2], pvel[2]));
                float coef = __fsub_rn(vnv, dot);
#pragma unroll
                for (int a = 0; a < 3; a++) pvel[a] = __fadd_rn(pvel[a], __fmul_rn(coef, sn[a]));
            }
        }
        if (dropped) {
            p.alive[gi] = 0;
            p.ijkOut[gi] = make_int3(vx, vy, vz);
            if (p.preAlive) p.preAlive[gi] = 0;
            continue;
        }
        uint32_t P[3], V[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float local = __double2float_rn(__dsub_rn((double)pIt[a], (double)pt[a]));
            P[a] = fx_encode(local);
            V[a] = h_encode(pvel[a]);
        }
        p.w0[gi] = P[0] | (P[1] << 16);
        p.w1[gi] = P[2] | (V[0] << 16);
        p.w2[gi] = V[1] | (V[2] << 16);
        p.alive[gi] = 1;
        p.ijkOut[gi] = make_int3(pt[0], pt[1], pt[2]);
        if (p.preAlive) {
            p.preAlive[gi] = 1;
#pragma unroll
            for (int a = 0; a < 3; a++) { p.prePos[3 * (size_t)gi + a] = pIt[a]; p.preVel[3 * (size_t)gi + a] = pvel[a]; }
        }
    }
}

// ---------------------------------------------------------------- g2p_tile_kernel (round 2)
// Same per-particle op sequence as g2p_advect_kernel, but every trilinear sample reads a SHARED-MEMORY TILE of the band grids
// around the particle's leaf instead of going through the leaf directory. ncu on the round-1 kernel: 10 800 SASS instructions,
// 38 % of all warp samples stalled on "no instruction" and a third of the issued instructions spent in the directory /
// leaf-boundary paths of fetch8 (a 2x2x2 cell straddles a leaf boundary for 1 - (7/8)^3 = 33 % of the samples, so in practice
// every warp executed both paths). With tiles a fetch is eight LDS off one base address and there is one code path.
//   14^3 tiles (leaf + 3 voxels each way; the CFL step is <= 3 voxels): the three new-velocity channels (first sample AND the
//         RK stage samples at displaced positions) and the solid SDF view (sampled at the old and the new position);
//   10^3 tiles (leaf + 1): old velocity, liquid SDF (+ the carried velocity when ViscousVelocity is a separate field).
// 63 KB per CTA (75 KB with the separate field): three CTAs per SM. A sample that leaves its tile (a particle faster than the
// CFL percentile) takes the round-1 global path, which lives in one non-inlined function per sampler.
constexpr int GT_H = 3, GT_T = 8 + 2 * GT_H, GT_N = GT_T * GT_T * GT_T;     // 14, 2744
constexpr int GS_H = 1, GS_T = 8 + 2 * GS_H, GS_N = GS_T * GS_T * GS_T;     // 10, 1000
__device__ __noinline__ float samplec_f32_slow(const G2PParams& p, const float* val, float x, float y, float z) {
    const Home h{0, 0, 0, 0};
    return samplec_f32(p.t, h, val, 0.f, x, y, z);
}
__device__ __noinline__ float box_f64_slow(const G2PParams& p, const float* val, float bg, double x, double y, double z) {
    const Home h{0, 0, 0, 0};
    return box_f64(p.t, h, val, bg, x, y, z);
}
__device__ __noinline__ float box_f64_solid_slow(const G2PParams& p, double x, double y, double z) {
    const Home h{0, 0, 0, 0};
    return box_f64_solid(p, h, x, y, z);
}
__device__ __noinline__ bool solid_normal_at_slow(const G2PParams& p, int qx, int qy, int qz, float n[3]) {
    const Home h{0, 0, 0, 0};
    return solid_normal_at(p, h, qx, qy, qz, n);
}
template <int T>
__device__ __forceinline__ bool tile_fetch8(const float* tile, int rx, int ry, int rz, float d[8]) {   // r = cell base - tile origin
    if ((unsigned)rx > (unsigned)(T - 2) || (unsigned)ry > (unsigned)(T - 2) || (unsigned)rz > (unsigned)(T - 2)) return false;
    const float* q = tile + (rx * T + ry) * T + rz;
    d[0] = q[0]; d[1] = q[1]; d[2] = q[T]; d[3] = q[T + 1];
    d[4] = q[T * T]; d[5] = q[T * T + 1]; d[6] = q[T * T + T]; d[7] = q[T * T + T + 1];
    return true;
}
struct TileCtx { int ox, oy, oz; };   // voxel coordinates of the leaf origin
template <int T, int H>
__device__ __forceinline__ float tile_samplec_f32(const G2PParams& p, const TileCtx& c, const float* tile, const float* val, float x, float y, float z) {
    const int bx = (int)floor((double)x), by = (int)floor((double)y), bz = (int)floor((double)z);
    float d[8];
    if (!tile_fetch8<T>(tile, bx - c.ox + H, by - c.oy + H, bz - c.oz + H, d)) return samplec_f32_slow(p, val, x, y, z);
    const float wx = __fsub_rn(x, (float)bx), wy = __fsub_rn(y, (float)by), wz = __fsub_rn(z, (float)bz);
    return mixf(mixf(mixf(d[0], d[1], wz), mixf(d[2], d[3], wz), wy), mixf(mixf(d[4], d[5], wz), mixf(d[6], d[7], wz), wy), wx);
}
template <int T, int H>
__device__ __forceinline__ float tile_box_f64(const G2PParams& p, const TileCtx& c, const float* tile, const float* val, float bg, double x, double y, double z) {
    const int bx = (int)floor(x), by = (int)floor(y), bz = (int)floor(z);
    float d[8];
    if (!tile_fetch8<T>(tile, bx - c.ox + H, by - c.oy + H, bz - c.oz + H, d)) return box_f64_slow(p, val, bg, x, y, z);
    return tri64(d, __dsub_rn(x, (double)bx), __dsub_rn(y, (double)by), __dsub_rn(z, (double)bz));
}
__device__ __forceinline__ float tile_box_f64_solid(const G2PParams& p, const TileCtx& c, const float* tile, double x, double y, double z) {
    const int bx = (int)floor(x), by = (int)floor(y), bz = (int)floor(z);
    float d[8];
    if (!tile_fetch8<GT_T>(tile, bx - c.ox + GT_H, by - c.oy + GT_H, bz - c.oz + GT_H, d)) return box_f64_solid_slow(p, x, y, z);
    return tri64(d, __dsub_rn(x, (double)bx), __dsub_rn(y, (double)by), __dsub_rn(z, (double)bz));
}
// leaf + H voxels of one channel -> shared memory; a voxel in no pool leaf reads the background (the solid view: the static grid)
template <int T, int H, bool SOLID>
__device__ __forceinline__ void tile_fill(const G2PParams& p, const TileCtx& c, float* tile, const float* __restrict__ val, float bg, const int* nb) {
    for (int i = threadIdx.x; i < T * T * T; i += G2P_THREADS) {
        const int rz = i % T, ry = (i / T) % T, rx = i / (T * T);
        const int vx = rx - H, vy = ry - H, vz = rz - H;          // relative to the leaf origin, in [-H, 8 + H)
        const int slot = nb[((vx >> 3) + 1) * 9 + ((vy >> 3) + 1) * 3 + ((vz >> 3) + 1)];
        float v = bg;
        if (slot >= 0) v = __ldg(&val[(size_t)slot * LEAF + (((vx & 7) << 6) | ((vy & 7) << 3) | (vz & 7))]);
        else if (SOLID && p.st.n > 0) v = grid_get(p.st, p.solidStatic, p.solidBg, c.ox + vx, c.oy + vy, c.oz + vz);
        tile[i] = v;
    }
}
template <bool SAME>
__global__ void __launch_bounds__(G2P_THREADS, 3) g2p_tile_kernel(const __grid_constant__ G2PParams p) {
    extern __shared__ __align__(16) float gsm[];
    __shared__ uint32_t sStart[LEAF + 1];
    __shared__ int sNb[27];
    float* tVel[3] = {gsm, gsm + GT_N, gsm + 2 * GT_N};
    float* tSolid = gsm + 3 * GT_N;
    float* tOld[3] = {tSolid + GT_N, tSolid + GT_N + GS_N, tSolid + GT_N + 2 * GS_N};
    float* tLsdf = tSolid + GT_N + 3 * GS_N;
    float* tCarr[3] = {tLsdf + GS_N, tLsdf + 2 * GS_N, tLsdf + 3 * GS_N};   // only allocated when !SAME
    const int leaf = blockIdx.x + p.leaf0;
    const size_t vbase = (size_t)leaf * LEAF;
    const uint32_t leafBeg = __ldg(&p.voxelStart[vbase]), leafEnd = __ldg(&p.voxelStart[vbase + LEAF]);
    if (leafEnd == leafBeg) return;
    for (int i = threadIdx.x; i <= LEAF; i += G2P_THREADS) sStart[i] = __ldg(&p.voxelStart[vbase + i]);
    if (threadIdx.x < 27) sNb[threadIdx.x] = __ldg(&p.t.nbr27[(size_t)leaf * 27 + threadIdx.x]);
    const int3 o = p.t.origin[leaf];
    const TileCtx c{o.x, o.y, o.z};
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; a++) {
        tile_fill<GT_T, GT_H, false>(p, c, tVel[a], p.vel[a], 0.f, sNb);
        tile_fill<GS_T, GS_H, false>(p, c, tOld[a], p.oldv[a], 0.f, sNb);
        if (!SAME) tile_fill<GS_T, GS_H, false>(p, c, tCarr[a], p.carr[a], 0.f, sNb);
    }
    tile_fill<GT_T, GT_H, true>(p, c, tSolid, p.solidView, p.solidBg, sNb);
    if (p.hasLiquid) tile_fill<GS_T, GS_H, false>(p, c, tLsdf, p.lsdf, p.lsdfBg, sNb);
    __syncthreads();
    const float dx = p.dx;
    const float deep_threshold = (float)(-4.0 * (double)dx);
    const float invdx = __fdiv_rn(1.0f, dx);
    const float dtinvx = __fdiv_rn(p.dt, dx);
    // staggered sample of the new velocity with OpenVDB's double-weight sampler (the RK stages)
    auto vel_f64 = [&](const float q[3], float out[3]) {
        out[0] = tile_box_f64<GT_T, GT_H>(p, c, tVel[0], p.vel[0], 0.f, __dadd_rn((double)q[0], 0.5), (double)q[1], (double)q[2]);
        out[1] = tile_box_f64<GT_T, GT_H>(p, c, tVel[1], p.vel[1], 0.f, (double)q[0], __dadd_rn((double)q[1], 0.5), (double)q[2]);
        out[2] = tile_box_f64<GT_T, GT_H>(p, c, tVel[2], p.vel[2], 0.f, (double)q[0], (double)q[1], __dadd_rn((double)q[2], 0.5));
    };
    for (uint32_t gi = leafBeg + threadIdx.x; gi < leafEnd; gi += G2P_THREADS) {
        int lo = 0, hi = LEAF;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sStart[mid] <= gi) lo = mid; else hi = mid; }
        const int off = lo;
        const int vx = o.x + (off >> 6), vy = o.y + ((off >> 3) & 7), vz = o.z + (off & 7);
        uint32_t a0 = p.w0[gi], a1 = p.w1[gi], a2 = p.w2[gi];
        float pIs[3] = {__fadd_rn((float)vx, fx_decode(a0 & 0xffffu)), __fadd_rn((float)vy, fx_decode(a0 >> 16)),
                        __fadd_rn((float)vz, fx_decode(a1 & 0xffffu))};
        float pvel[3] = {h_decode(a1 >> 16), h_decode(a2 & 0xffffu), h_decode(a2 >> 16)};
        float adv[3], old[3], carried[3];
        adv[0] = tile_samplec_f32<GT_T, GT_H>(p, c, tVel[0], p.vel[0], __fadd_rn(pIs[0], 0.5f), pIs[1], pIs[2]);
        adv[1] = tile_samplec_f32<GT_T, GT_H>(p, c, tVel[1], p.vel[1], pIs[0], __fadd_rn(pIs[1], 0.5f), pIs[2]);
        adv[2] = tile_samplec_f32<GT_T, GT_H>(p, c, tVel[2], p.vel[2], pIs[0], pIs[1], __fadd_rn(pIs[2], 0.5f));
        old[0] = tile_samplec_f32<GS_T, GS_H>(p, c, tOld[0], p.oldv[0], __fadd_rn(pIs[0], 0.5f), pIs[1], pIs[2]);
        old[1] = tile_samplec_f32<GS_T, GS_H>(p, c, tOld[1], p.oldv[1], pIs[0], __fadd_rn(pIs[1], 0.5f), pIs[2]);
        old[2] = tile_samplec_f32<GS_T, GS_H>(p, c, tOld[2], p.oldv[2], pIs[0], pIs[1], __fadd_rn(pIs[2], 0.5f));
        float flip = __fsub_rn(1.0f, p.picMin);
        float pls = p.hasLiquid ? tile_box_f64<GS_T, GS_H>(p, c, tLsdf, p.lsdf, p.lsdfBg, (double)pIs[0], (double)pIs[1], (double)pIs[2]) : p.lsdfBg;
        float t_coef = 1.f;
        if (pls < 0.f && pls >= -p.surfacedist) {
            t_coef = __fdiv_rn(pls, -p.surfacedist);
            t_coef = fminf(fmaxf(t_coef, 0.0f), 1.0f);
        }
        if (pls >= 0.f) t_coef = 0.f;
        if (p.surfacedist > 0.f)
            flip = __fadd_rn(__fmul_rn(t_coef, flip), __fmul_rn(__fsub_rn(1.0f, t_coef), fminf(__fsub_rn(1.0f, p.picMax), flip)));
        float pss = tile_box_f64_solid(p, c, tSolid, (double)__fadd_rn(pIs[0], 0.5f), (double)__fadd_rn(pIs[1], 0.5f), (double)__fadd_rn(pIs[2], 0.5f));
        if (pss >= 0.f && (double)pss <= 2.0 * (double)dx) {
            float scoef = __fdiv_rn(pss, __fmul_rn(2.0f, dx));
            flip = __fadd_rn(__fmul_rn(scoef, flip), __fmul_rn(__fsub_rn(1.0f, scoef), 1.0f));
        }
        if (SAME) { carried[0] = adv[0]; carried[1] = adv[1]; carried[2] = adv[2]; }
        else {
            carried[0] = tile_samplec_f32<GS_T, GS_H>(p, c, tCarr[0], p.carr[0], __fadd_rn(pIs[0], 0.5f), pIs[1], pIs[2]);
            carried[1] = tile_samplec_f32<GS_T, GS_H>(p, c, tCarr[1], p.carr[1], pIs[0], __fadd_rn(pIs[1], 0.5f), pIs[2]);
            carried[2] = tile_samplec_f32<GS_T, GS_H>(p, c, tCarr[2], p.carr[2], pIs[0], pIs[1], __fadd_rn(pIs[2], 0.5f));
        }
#pragma unroll
        for (int a = 0; a < 3; a++) pvel[a] = __fadd_rn(carried[a], __fmul_rn(flip, __fadd_rn(-old[a], pvel[a])));
        float pIt[3] = {pIs[0], pIs[1], pIs[2]};
        {   // custom_integrator (FF/FLIP_vdb.cpp:169-214), same expressions as integrate()
            const int order = (pls >= -p.surfacedist) ? 1 : p.rkOrder;
            float q[3], V1[3], V2[3], V3[3];
            if (order == 2) {
#pragma unroll
                for (int a = 0; a < 3; a++) q[a] = __fadd_rn(pIt[a], __fmul_rn(__fmul_rn(0.5f, adv[a]), dtinvx));
                vel_f64(q, V1);
#pragma unroll
                for (int a = 0; a < 3; a++) pIt[a] = __fadd_rn(pIt[a], __fmul_rn(V1[a], dtinvx));
            } else if (order == 3) {
#pragma unroll
                for (int a = 0; a < 3; a++) q[a] = __fadd_rn(pIt[a], __fmul_rn(__fmul_rn(0.5f, adv[a]), dtinvx));
                vel_f64(q, V1);
#pragma unroll
                for (int a = 0; a < 3; a++) q[a] = __fadd_rn(pIt[a], __fmul_rn(dtinvx, __fsub_rn(__fmul_rn(2.0f, V1[a]), adv[a])));
                vel_f64(q, V2);
#pragma unroll
                for (int a = 0; a < 3; a++)
                    pIt[a] = __fadd_rn(pIt[a], __fmul_rn(__fmul_rn(dtinvx, __fadd_rn(__fadd_rn(adv[a], __fmul_rn(4.0f, V1[a])), V2[a])), (1.0f / 6.0f)));
            } else if (order == 4) {
#pragma unroll
                for (int a = 0; a < 3; a++) q[a] = __fadd_rn(pIt[a], __fmul_rn(__fmul_rn(0.5f, adv[a]), dtinvx));
                vel_f64(q, V1);
#pragma unroll
                for (int a = 0; a < 3; a++) q[a] = __fadd_rn(pIt[a], __fmul_rn(__fmul_rn(0.5f, V1[a]), dtinvx));
                vel_f64(q, V2);
#pragma unroll
                for (int a = 0; a < 3; a++) q[a] = __fadd_rn(pIt[a], __fmul_rn(V2[a], dtinvx));
                vel_f64(q, V3);
#pragma unroll
                for (int a = 0; a < 3; a++)
                    pIt[a] = __fadd_rn(pIt[a], __fmul_rn(__fmul_rn(dtinvx, __fadd_rn(__fadd_rn(adv[a], __fmul_rn(2.0f, __fadd_rn(V1[a], V2[a]))), V3[a])), (1.0f / 6.0f)));
            } else {
#pragma unroll
                for (int a = 0; a < 3; a++) pIt[a] = __fadd_rn(pIt[a], __fmul_rn(adv[a], dtinvx));
            }
        }
        int pt[3];
#pragma unroll
        for (int a = 0; a < 3; a++) pt[a] = (int)floor((double)__fadd_rn(pIt[a], 0.5f));
        float nps = tile_box_f64_solid(p, c, tSolid, (double)__fadd_rn(pIt[0], 0.5f), (double)__fadd_rn(pIt[1], 0.5f), (double)__fadd_rn(pIt[2], 0.5f));
        bool dropped = false;
        if (nps < 0.f) {
            if (nps < deep_threshold) dropped = true;
            else {
                float sn[3];
                solid_normal_at_slow(p, pt[0], pt[1], pt[2], sn);
#pragma unroll
                for (int a = 0; a < 3; a++) pIt[a] = __fsub_rn(pIt[a], __fmul_rn(__fmul_rn(__fmul_rn(nps, sn[a]), invdx), 1.0f));
#pragma unroll
                for (int a = 0; a < 3; a++) pt[a] = (int)floor((double)__fadd_rn(pIt[a], 0.5f));
                float vnv = 0.f;
                if (p.hasSolidVel) {
                    float n2[3];
                    if (solid_normal_at_slow(p, pt[0], pt[1], pt[2], n2)) {
                        int l = topo_find(p.t, pt[0], pt[1], pt[2]);  // on => inside the pool
                        size_t k = (size_t)l * LEAF + voxel_off(pt[0], pt[1], pt[2]);
                        vnv = __fadd_rn(__fadd_rn(__fmul_rn(p.svelView[0][k], n2[0]), __fmul_rn(p.svelView[1][k], n2[1])), __fmul_rn(p.svelView[2][k], n2[2]));
                    }
                }
                float dot = __fadd_rn(__fadd_rn(__fmul_rn(sn[0], pvel[0]), __fmul_rn(sn[1], pvel[1])), __fmul_rn(sn[2], pvel[2]));
                float coef = __fsub_rn(vnv, dot);
#pragma unroll
                for (int a = 0; a < 3; a++) pvel[a] = __fadd_rn(pvel[a], __fmul_rn(coef, sn[a]));
            }
        }
        if (dropped) {
            p.alive[gi] = 0;
            p.ijkOut[gi] = make_int3(vx, vy, vz);
            if (p.preAlive) p.preAlive[gi] = 0;
            continue;
        }
        uint32_t P[3], V[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float local = __double2float_rn(__dsub_rn((double)pIt[a], (double)pt[a]));
            P[a] = fx_encode(local);
            V[a] = h_encode(pvel[a]);
        }
        p.w0[gi] = P[0] | (P[1] << 16);
        p.w1[gi] = P[2] | (V[0] << 16);
        p.w2[gi] = V[1] | (V[2] << 16);
        p.alive[gi] = 1;
        p.ijkOut[gi] = make_int3(pt[0], pt[1], pt[2]);
        if (p.preAlive) {
            p.preAlive[gi] = 1;
#pragma unroll
            for (int a = 0; a < 3; a++) { p.prePos[3 * (size_t)gi + a] = pIt[a]; p.preVel[3 * (size_t)gi + a] = pvel[a]; }
        }
    }
}
}  // namespace

namespace {
__global__ void fill_const_kernel(float* __restrict__ p, float v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void fill_const(World* w, float* p, float v, size_t n) {
    if (!n) return;
    FB_LAUNCH(w, "fill", n * 4) fill_const_kernel<<<(unsigned)((n + 255) / 256), 256, 0, w->stream>>>(p, v, n);
    check_launch("fill_const");
}
}  // namespace

// G2PAdvectorSheet::apply (FF/nosys/SheetG2PAdvector.cpp:15-54)
void g2p_advect_sheetty(World* w, float dt, float dx, int surfaceSize, int rkOrder, float picMin, float picMax, int flags) {
    FB_REQUIRE(w->pts.topo != nullptr, FLIPB200_ERR_STATE, "G2PAdvectorSheetty: no particles");
    // flags & 2: the plain G2P_Advector (FF/nosys/G2P_Advector.cpp -> FLIP_vdb::Advect, FF/FLIP_vdb.cpp:3209-3219): no liquid SDF
    // (the sampled value is the background dx >= 0, so every particle takes the Euler step and the FLIP factor is 1 - pic_smoothness),
    // no solids, velocity carried = velocity advected. The Sheetty node clamps pic_min to pic_max (SheetG2PAdvector.cpp:49-52), Advect does not.
    const bool plain = (flags & 2) != 0;
    if (!plain) picMin = picMin > picMax ? picMax : picMin;
    const bool same = (flags & 1) != 0 || plain;
    if (same) ensure_pool(w, {FLIPB200_VELOCITY, FLIPB200_POSTADV_VELOCITY, FLIPB200_LIQUID_SDF}, true);
    else ensure_pool(w, {FLIPB200_VELOCITY, FLIPB200_POSTADV_VELOCITY, FLIPB200_VISCOUS_VELOCITY, FLIPB200_LIQUID_SDF}, true);
    refresh_solid_views(w);
    TopoPtr pool = w->pool;
    const uint64_t n = w->pts.n;
    const int nl = pool->n;
    GridV& vel = w->V(FLIPB200_VELOCITY);
    GridV& oldv = w->V(FLIPB200_POSTADV_VELOCITY);
    GridV& carr = same ? vel : w->V(FLIPB200_VISCOUS_VELOCITY);
    GridF& lsdf = w->F(FLIPB200_LIQUID_SDF);

    FB_PHASE(w, "g2p total");
    // K8 support: dilate5(liquid sdf topology), 26-neighbourhood (FF/FLIP_vdb.cpp:3277-3279)
    DBuf<uint64_t> nm((size_t)nl * 8 + 1, w->stream), nm2((size_t)nl * 8 + 1, w->stream);
    if (w->hasSolidSDF && nl && !plain) {
        mask_dilate(w, *pool, lsdf.mask.p, nm.p, true);
        for (int i = 1; i < 5; i++) { mask_dilate(w, *pool, nm.p, nm2.p, true); std::swap(nm, nm2); }
    }
    DBuf<int3> ijk(n + 1, w->stream);
    DBuf<uint8_t> alive(n + 1, w->stream);
    if (w->capturePreCodec) {
        w->preCodecPos.alloc(3 * n + 1, w->stream);
        w->preCodecVel.alloc(3 * n + 1, w->stream);
        w->preCodecAlive.alloc(n + 1, w->stream);
        w->preCodecN = n;
    }
    G2PParams p;
    p.t = pool->view();
    p.voxelStart = w->pts.voxelStart.p;
    p.w0 = w->pts.w0.p; p.w1 = w->pts.w1.p; p.w2 = w->pts.w2.p;
    for (int c = 0; c < 3; c++) { p.vel[c] = vel.val[c].p; p.oldv[c] = oldv.val[c].p; p.carr[c] = carr.val[c].p; p.svelView[c] = w->solidVelView[c].p; }
    p.lsdf = lsdf.val.p; p.lsdfBg = lsdf.bg; p.hasLiquid = 1;
    p.solidView = w->solidSdfView.p;
    p.solidBg = w->hasSolidSDF ? w->F(FLIPB200_SOLID_SDF).bg : 3.0f * dx;
    p.hasSolid = w->hasSolidSDF ? 1 : 0;
    GridF& ss = w->F(FLIPB200_SOLID_SDF);
    p.st = (w->hasSolidSDF && ss.topo) ? ss.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
    p.solidStatic = ss.val.p;
    p.hasSolidVel = (w->hasSolidVel && w->V(FLIPB200_SOLID_VELOCITY).leaves() > 0) ? 1 : 0;
    p.nmask = nm.p;
    DBuf<float> noSolid;
    if (plain) {
        // everything the kernel can read about liquid depth and solids is the background
        p.hasLiquid = 0; p.lsdfBg = dx;
        noSolid.alloc((size_t)nl * LEAF + 1, w->stream);
        const float bg3 = 3.0f * dx;
        fill_const(w, noSolid.p, bg3, (size_t)nl * LEAF);
        p.solidView = noSolid.p; p.solidBg = bg3; p.hasSolid = 0; p.hasSolidVel = 0;
        p.st = TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
    }
    p.dx = dx; p.dt = dt; p.picMin = picMin; p.picMax = picMax; p.surfacedist = (float)surfaceSize * dx;
    p.rkOrder = rkOrder; p.sameField = same ? 1 : 0;
    p.ijkOut = ijk.p; p.alive = alive.p;
    p.prePos = w->capturePreCodec ? w->preCodecPos.p : nullptr;
    p.preVel = w->capturePreCodec ? w->preCodecVel.p : nullptr;
    p.preAlive = w->capturePreCodec ? w->preCodecAlive.p : nullptr;
    // slab decomposition: only the owned leaves are advected (ghost particles are re-imported below)
    int leafLo = 0, leafHi = nl;
    const bool dd = dd_on(w);
    if (dd) dd_owned_slots(w, &leafLo, &leafHi);
    p.leaf0 = leafLo;
    if (leafHi > leafLo && n) {
        FB_PHASE(w, "g2p kernel");
        // compulsory traffic: 12 B read + 12 B write per particle + the band grids once
        FB_LAUNCH(w, "g2p_advect", n * 24 + (size_t)nl * LEAF * 28 + (size_t)nl * LEAF * 4)
            if (!getenv("FLIPB200_G2P_OLD")) {
                const size_t smem = (size_t)(4 * GT_N + (same ? 4 : 7) * GS_N) * sizeof(float);
                static bool attr = false;
                if (!attr) {
                    FB_CUDA(cudaFuncSetAttribute(g2p_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(4 * GT_N + 4 * GS_N) * sizeof(float))));
                    FB_CUDA(cudaFuncSetAttribute(g2p_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(4 * GT_N + 7 * GS_N) * sizeof(float))));
                    attr = true;
                }
                if (same) g2p_tile_kernel<true><<<leafHi - leafLo, G2P_THREADS, smem, w->stream>>>(p);
                else g2p_tile_kernel<false><<<leafHi - leafLo, G2P_THREADS, smem, w->stream>>>(p);
            } else {
                static const int occ = getenv("FLIPB200_G2P_OCC") ? atoi(getenv("FLIPB200_G2P_OCC")) : 5;   // 48 registers, 5 CTAs/SM: 2.23 ms against 2.37 ms at 80 registers / 3 CTAs (B200, 16.8 M particles)
                if (occ >= 5) g2p_advect_kernel<5><<<leafHi - leafLo, G2P_THREADS, 0, w->stream>>>(p);
                else if (occ == 4) g2p_advect_kernel<4><<<leafHi - leafLo, G2P_THREADS, 0, w->stream>>>(p);
                else g2p_advect_kernel<3><<<leafHi - leafLo, G2P_THREADS, 0, w->stream>>>(p);
            }
        check_launch("g2p_advect");
    }
    DBuf<uint32_t> i0 = std::move(w->pts.w0), i1 = std::move(w->pts.w1), i2 = std::move(w->pts.w2);
    if (dd) {
        // migrants and ghost copies cross the slab faces; the merged set replaces the local one
        uint32_t range[2] = {0, 0};
        if (nl) {
            FB_CUDA(cudaMemcpyAsync(&range[0], w->pts.voxelStart.p + (size_t)leafLo * LEAF, 4, cudaMemcpyDeviceToHost, w->stream));
            FB_CUDA(cudaMemcpyAsync(&range[1], w->pts.voxelStart.p + (size_t)leafHi * LEAF, 4, cudaMemcpyDeviceToHost, w->stream));
            sync(w);
        }
        DBuf<uint32_t> m0, m1, m2;
        DBuf<int3> mijk;
        uint64_t nm = 0;
        { FB_PHASE(w, "g2p dd_migrate"); dd_migrate(w, range[0], range[1], i0.p, i1.p, i2.p, ijk.p, alive.p, m0, m1, m2, mijk, &nm); }
        TopoPtr newPool;
        { FB_PHASE(w, "g2p topo");
          newPool = topo_from_origins_dev(w, mijk.p, (int)nm, true); }   // the topology kernels only use (coordinate >> 3): voxel coordinates do
        DBuf<uint32_t> keys(nm + 1, w->stream);
        FB_PHASE(w, "g2p rebin");
        keys_from_ijk(w, newPool, mijk.p, nullptr, nm, keys.p);
        rebin_particles(w, newPool, keys.p, nm, m0, m1, m2);
        w->pool = newPool;
        return;
    }
    // K2: new pool from the target leaves (+ring), keys, stable counting sort with the voxel cap
    TopoPtr newPool;
    { FB_PHASE(w, "g2p topo");
      newPool = topo_from_origins_dev(w, ijk.p, (int)n, true); }   // the topology kernels only use (coordinate >> 3): voxel coordinates do
    FB_PHASE(w, "g2p rebin");
    DBuf<uint32_t> keys(n + 1, w->stream);
    keys_from_ijk(w, newPool, ijk.p, alive.p, n, keys.p);
    rebin_particles(w, newPool, keys.p, n, i0, i1, i2);
    w->pool = newPool;
}

}  // namespace fb
