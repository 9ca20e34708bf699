// TEST INFRASTRUCTURE ONLY -- executes the DEVICE code of zeno_b200/csrc/next_kernels.cuh on the CPU.
//
// The per-thread bodies of the kernels behind KillParticlesInSDF, ParticleAddDV, VDBRenormalizeSDF and VDBErodeSDF are
// compiled here with plain g++: the CUDA round-to-nearest intrinsics they use are mapped to the IEEE host operations they
// stand for (this file is built with -ffp-contract=off, SSE arithmetic), __ldg is a load, fp16 conversions come from
// cuda_fp16.h's own host implementations. tests/test_next_kernels_emul_cpu.py drives every (leaf, thread) of a launch
// through these bodies and compares with the oracle bit for bit -- the arithmetic and indexing of the kernels is checked
// before their first GPU run. What this cannot check: launch configuration, shared-memory staging, stream ordering.
#include <cmath>
#include <cstdint>
#include <cstring>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __double2float_rn(double a) { return (float)a; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

#include "../../zeno_b200/csrc/next_kernels.cuh"

using namespace fb;

static TopoView view(int n, const int* dmin, const int* ddim, const int* dir, const int* origins) {
    return TopoView{n, make_int3(dmin[0], dmin[1], dmin[2]), make_int3(ddim[0], ddim[1], ddim[2]), dir,
                    reinterpret_cast<const int3*>(origins), nullptr};
}

extern "C" {
// one full launch of renorm_stage_kernel: grid = n leaves, block = 512 threads
void emul_renorm_stage(int n, const int* dmin, const int* ddim, const int* dir, const int* origins, const uint64_t* mask,
                       const float* cur, const float* phi0, float* out, float bg, float dt, float invDx, float alpha, float beta, int useAlpha) {
    const TopoView t = view(n, dmin, ddim, dir, origins);
    for (int leaf = 0; leaf < n; leaf++)
        for (int off = 0; off < LEAF; off++) nextk::renorm_stage_one(t, mask, cur, phi0, out, bg, dt, invDx, alpha, beta, useAlpha, leaf, off);
}
void emul_box_avg(int n, const int* dmin, const int* ddim, const int* dir, const int* origins, const uint64_t* mask, const float* cur,
                  float* out, float bg, int axis, int w, float frac) {
    const TopoView t = view(n, dmin, ddim, dir, origins);
    for (int leaf = 0; leaf < n; leaf++)
        for (int off = 0; off < LEAF; off++) nextk::box_avg_one(t, mask, cur, out, bg, axis, w, frac, leaf, off);
}
void emul_add_active(int n, const uint64_t* mask, float* val, float d) {
    for (int leaf = 0; leaf < n; leaf++)
        for (int off = 0; off < LEAF; off++) nextk::add_active_one(mask, val, leaf, off, d);
}
void emul_add_dv(uint32_t* w1, uint32_t* w2, uint64_t n, double dx, double dy, double dz) {
    for (uint64_t i = 0; i < n; i++) nextk::add_dv_one(w1, w2, i, dx, dy, dz);
}
// one full launch of kill_keys_kernel: grid = n store leaves; the CTA's shared prefix is the leaf's slice of voxelStart
void emul_kill_keys(int n, const int* dmin, const int* ddim, const int* dir, const int* origins, const uint32_t* voxelStart,
                    uint32_t* w0, uint32_t* w1, int sn, const int* sdmin, const int* sddim, const int* sdir, const int* sorigins,
                    const float* sval, float sbg, int keep, uint32_t* keys) {
    const TopoView pt = view(n, dmin, ddim, dir, origins), st = view(sn, sdmin, sddim, sdir, sorigins);
    for (int leaf = 0; leaf < n; leaf++) {
        const uint32_t* sStart = voxelStart + (size_t)leaf * LEAF;
        for (uint32_t gi = sStart[0]; gi < sStart[LEAF]; gi++) nextk::kill_keys_one(pt, sStart, w0, w1, st, sval, sbg, keep, keys, leaf, gi);
    }
}
}
