// libflipb200 -- per-thread bodies of the kernels behind the nodes beyond the substep chain.
// They live in a header so that tests/emul/ can execute exactly this device code on the CPU (plain C++ with the CUDA
// intrinsics mapped to their IEEE host equivalents) and compare it with the oracle before the kernels ever run on a GPU.
#pragma once
#include "common.cuh"

namespace fb {
namespace nextk {

// openvdb::tools::BoxSampler::sample, double weights (tools/Interpolation.h:712-737,763-778): a + float((b - a) * w)
__device__ __forceinline__ float ip64(float a, float b, double w) {
    return __fadd_rn(a, __double2float_rn(__dmul_rn((double)__fsub_rn(b, a), w)));
}
__device__ __forceinline__ float box_sample_f64(const TopoView& t, const float* __restrict__ val, float bg, double x, double y, double z) {
    const int bx = (int)floor(x), by = (int)floor(y), bz = (int)floor(z);
    const double u = __dsub_rn(x, (double)bx), v = __dsub_rn(y, (double)by), w = __dsub_rn(z, (double)bz);
    float d[8];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) d[i * 4 + j * 2 + k] = grid_get(t, val, bg, bx + i, by + j, bz + k);
    return ip64(ip64(ip64(d[0], d[1], w), ip64(d[2], d[3], w), v), ip64(ip64(d[4], d[5], w), ip64(d[6], d[7], w), v), u);
}

constexpr uint32_t KEY_DROPPED_NEXT = 0xffffffffu;

// kill_particles_inside (FF/nosys/KillParticles.cpp:13-149) for particle gi of store leaf `leaf`; sStart = the leaf's 513
// voxel prefix entries. key = the particle's own voxel if it survives, dropped otherwise; survivors' positions go through
// decode -> encode once (the reference rewrites them through the attribute write handle, :138-141)
__device__ __forceinline__ void kill_keys_one(const TopoView& pt, const uint32_t* sStart, uint32_t* __restrict__ w0, uint32_t* __restrict__ w1,
                                              const TopoView& st, const float* __restrict__ sval, float sbg, int keep,
                                              uint32_t* __restrict__ keys, int leaf, uint32_t gi) {
    int lo = 0, hi = LEAF;   // voxel of this particle: largest off with sStart[off] <= gi
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sStart[mid] <= gi) lo = mid; else hi = mid; }
    const int off = lo;
    const int3 o = pt.origin[leaf];
    const uint32_t a0 = w0[gi], a1 = w1[gi];
    const float px = fx_decode(a0 & 0xffffu), py = fx_decode(a0 >> 16), pz = fx_decode(a1 & 0xffffu);
    const float x = __fadd_rn((float)(o.x + (off >> 6)), px), y = __fadd_rn((float)(o.y + ((off >> 3) & 7)), py),
                z = __fadd_rn((float)(o.z + (off & 7)), pz);
    const float s = box_sample_f64(st, sval, sbg, (double)x, (double)y, (double)z);
    const bool alive = keep ? (s <= 0.f) : (s >= 0.f);
    if (alive) {
        w0[gi] = fx_encode(px) | (fx_encode(py) << 16);
        w1[gi] = fx_encode(pz) | (a1 & 0xffff0000u);
    }
    keys[gi] = alive ? (uint32_t)leaf * LEAF + (uint32_t)off : KEY_DROPPED_NEXT;
}

// FLIP_vdb::point_integrate_vector, channel "vel" (FF/FLIP_vdb.cpp:3526-3532): half -> double, + dv, -> float -> half
__device__ __forceinline__ void add_dv_one(uint32_t* __restrict__ w1, uint32_t* __restrict__ w2, uint64_t i, double dx, double dy, double dz) {
    const uint32_t b = w1[i], c = w2[i];
    const float vx = __double2float_rn(__dadd_rn((double)h_decode(b >> 16), dx));
    const float vy = __double2float_rn(__dadd_rn((double)h_decode(c & 0xffffu), dy));
    const float vz = __double2float_rn(__dadd_rn((double)h_decode(c >> 16), dz));
    w1[i] = (b & 0xffffu) | (h_encode(vx) << 16);
    w2[i] = h_encode(vy) | (h_encode(vz) << 16);
}

// VDBErodeSDF: active voxels += d
__device__ __forceinline__ void add_active_one(const uint64_t* __restrict__ mask, float* __restrict__ val, int leaf, int off, float d) {
    if (mask_get(mask, leaf, off)) { const size_t i = (size_t)leaf * LEAF + off; val[i] = __fadd_rn(val[i], d); }
}

// One Euler stage of openvdb::tools::LevelSetTracker's Normalizer (tools/LevelSetTracker.h:631-675) with the first-order
// upwind Godunov norm (math/Operators.h:249-260, math/FiniteDifference.h:326-347): an ACTIVE voxel of `cur` gets
// alpha*phi0 + beta*v (v alone when useAlpha == 0); an inactive voxel keeps its value; the stencil reads `cur` wherever it lands.
__device__ __forceinline__ void renorm_stage_one(const TopoView& t, const uint64_t* __restrict__ mask, const float* __restrict__ cur,
                                                 const float* __restrict__ phi0, float* __restrict__ out, float bg, float dt, float invDx,
                                                 float alpha, float beta, int useAlpha, int leaf, int off) {
    const size_t i = (size_t)leaf * LEAF + off;
    const float c = cur[i];
    float r = c;
    if (mask_get(mask, leaf, off)) {
        const int3 o = t.origin[leaf];
        const int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
        const float ux = __fsub_rn(grid_get(t, cur, bg, x + 1, y, z), c), uy = __fsub_rn(grid_get(t, cur, bg, x, y + 1, z), c),
                    uz = __fsub_rn(grid_get(t, cur, bg, x, y, z + 1), c);
        const float dxm = __fsub_rn(c, grid_get(t, cur, bg, x - 1, y, z)), dym = __fsub_rn(c, grid_get(t, cur, bg, x, y - 1, z)),
                    dzm = __fsub_rn(c, grid_get(t, cur, bg, x, y, z - 1));
        float n2;
        if (c > 0.f) {
            float a = fmaxf(dxm, 0.f), b = fminf(ux, 0.f);
            n2 = fmaxf(__fmul_rn(a, a), __fmul_rn(b, b));
            a = fmaxf(dym, 0.f); b = fminf(uy, 0.f);
            n2 = __fadd_rn(n2, fmaxf(__fmul_rn(a, a), __fmul_rn(b, b)));
            a = fmaxf(dzm, 0.f); b = fminf(uz, 0.f);
            n2 = __fadd_rn(n2, fmaxf(__fmul_rn(a, a), __fmul_rn(b, b)));
        } else {
            float a = fminf(dxm, 0.f), b = fmaxf(ux, 0.f);
            n2 = fmaxf(__fmul_rn(a, a), __fmul_rn(b, b));
            a = fminf(dym, 0.f); b = fmaxf(uy, 0.f);
            n2 = __fadd_rn(n2, fmaxf(__fmul_rn(a, a), __fmul_rn(b, b)));
            a = fminf(dzm, 0.f); b = fmaxf(uz, 0.f);
            n2 = __fadd_rn(n2, fmaxf(__fmul_rn(a, a), __fmul_rn(b, b)));
        }
        float v = __fdiv_rn(c, __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(c, c), n2)), 1e-8f));
        v = __fsub_rn(c, __fmul_rn(__fmul_rn(dt, v), __fsub_rn(__fmul_rn(__fsqrt_rn(n2), invDx), 1.0f)));
        r = useAlpha ? __fadd_rn(__fmul_rn(alpha, phi0[i]), __fmul_rn(beta, v)) : v;
    }
    out[i] = r;
}

// One pass of openvdb::tools::Filter's separable box filter (tools/Filter.h:503-514,744-768, no alpha mask): an ACTIVE voxel of `cur`
// gets (sum over k = -w..w of cur at the voxel shifted by k along `axis`, added in that order) * frac; the taps read `cur` wherever
// they land (inactive voxels and the background included); an inactive voxel keeps its value.
__device__ __forceinline__ void box_avg_one(const TopoView& t, const uint64_t* __restrict__ mask, const float* __restrict__ cur,
                                            float* __restrict__ out, float bg, int axis, int w, float frac, int leaf, int off) {
    const size_t i = (size_t)leaf * LEAF + off;
    float r = cur[i];
    if (mask_get(mask, leaf, off)) {
        const int3 o = t.origin[leaf];
        int c[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
        const int centre = c[axis];
        float sum = 0.f;
        for (int k = -w; k <= w; k++) {
            c[axis] = centre + k;
            sum = __fadd_rn(sum, grid_get(t, cur, bg, c[0], c[1], c[2]));
        }
        r = __fmul_rn(sum, frac);
    }
    out[i] = r;
}

}  // namespace nextk
}  // namespace fb
