// libflipb200 -- matrix-free multigrid-preconditioned CG pressure projection (K9-K13).
// Follows FF/simd_vdb_poisson_uaamg.cpp: variational ghost-fluid 7-point Laplacian on 8^3
// leaves (BuildFinestMatrix :623-747), Galerkin coarsening x 1/8 x 1/2 (:409-620), default-leaf
// trimming (:1905-1960), red-black SOR smoothing w = 1.2 (:1109-1150), piecewise-constant
// restriction / prolongation (:1773-1903), mu = 2 cycle preconditioner (:1993-2126) and
// McAdams-style PCG on the L-inf residual (:2332-2403).
//
// Every level is a set of dense [leaf][512] fp32 arrays on a Topo; coefficient leaves that the
// reference would trim are overwritten with the defaults (so reads are identical) and flagged,
// and the stencil kernels take a constant-coefficient path on flagged leaves (no coefficient
// traffic). The stencil arithmetic keeps the reference's association and its explicit FMAs.
#include "world.cuh"
#include <climits>
#include "levelset.cuh"
#include <algorithm>
#include <cstring>

namespace fb {
namespace {

constexpr int NB_XM = 4, NB_XP = 22, NB_YM = 10, NB_YP = 16, NB_ZM = 12, NB_ZP = 14;  // nbr27 indices
constexpr int MAX_COARSEST = 4000;  // uaamg.cpp:1965
constexpr int RED_THREADS = 256;

// everything a stencil CTA needs to know about its leaf, in one 32-byte record (one load instead of the
// dependent mask -> nbr27 -> flags[nbr] chain): the six face-neighbour slots and
// bit0 = leaf has a DOF, bit1 = all face coefficients this leaf reads are the default (-term)
struct __align__(128) LeafInfo { int nb[6]; uint32_t flags; uint32_t pad; uint64_t mask[8]; uint64_t pad2[4]; };
enum { LI_ANY = 1, LI_CONST = 2, LI_DIAG = 4, LI_OCT_SHIFT = 8 };  // LI_DIAG: diag / invdiag read as the default on the whole leaf

struct Level {
    TopoPtr topo;
    int n = 0;
    float dx = 0.f, term = 0.f;
    int numDof = 0;
    DBuf<uint64_t> dof;
    DBuf<float> diag, invdiag, xe, ye, ze;
    DBuf<uint8_t> flags;   // bit0 diag, bit1 x, bit2 y, bit3 z read as the default
    bool hasParent = false;   // LeafInfo::pad / octant bits are set (leaf_parent_kernel)
    DBuf<LeafInfo> info;
    DBuf<float> x, b;
    // candidate leaves of the next coarser level, listed speculatively while the DOF count is being read back
    // (one host wait per level instead of two)
    DBuf<int3> cand;
    int candCount = 0;
    int ownLo = 0, ownHi = -1;   // slab decomposition: leaves whose DOFs this rank owns (reductions); -1 = all
};
struct LevelView {
    TopoView t;
    const uint64_t* dof;
    const float *diag, *invdiag, *xe, *ye, *ze;
    const uint8_t* flags;
    const LeafInfo* info;
    float term;
    int ownLo, ownHi;   // leaves that contribute to dot products / norms (everything on a single GPU)
};
LevelView view_of(const Level& L) {
    return LevelView{L.topo->view(), L.dof.p, L.diag.p, L.invdiag.p, L.xe.p, L.ye.p, L.ze.p, L.flags.p, L.info.p, L.term,
                     L.ownLo, L.ownHi < 0 ? L.n : L.ownHi};
}
__device__ __forceinline__ bool owned_leaf(const LevelView& L, int leaf) { return leaf >= L.ownLo && leaf < L.ownHi; }

// ---------------------------------------------------------------- reductions (deterministic)
// block-wide sum / max of one value per thread (512 threads), result valid in thread 0
__device__ __forceinline__ float block_sum_512(float v, float* sm16) {
    for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0) sm16[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sm16[threadIdx.x] : 0.f;
        for (int d = 16; d > 0; d >>= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, d));
    }
    return r;
}
__device__ __forceinline__ float block_absmax_512(float v, float* sm16) {
    // non-finite values propagate (FF/openvdb_grid_math_op.h:17-25)
    float a = isfinite(v) ? fabsf(v) : v;
    for (int d = 16; d > 0; d >>= 1) {
        float o = __shfl_xor_sync(0xffffffffu, a, d);
        a = (isfinite(a) ? (isfinite(o) ? fmaxf(a, o) : o) : a);
    }
    if ((threadIdx.x & 31) == 0) sm16[threadIdx.x >> 5] = a;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sm16[threadIdx.x] : 0.f;
        for (int d = 16; d > 0; d >>= 1) {
            float o = __shfl_xor_sync(0xffffffffu, r, d);
            r = (isfinite(r) ? (isfinite(o) ? fmaxf(r, o) : o) : r);
        }
    }
    return r;
}

// ---------------------------------------------------------------- matrix construction
// BuildFinestMatrix::operator() (uaamg.cpp:651-735); DOF = phi active, phi < 0, diagonal != 0
__global__ void __launch_bounds__(512) build_finest_kernel(TopoView t, const float* __restrict__ phi, float phiBg,
                                                           const uint64_t* __restrict__ phiMask,
                                                           const float* __restrict__ fw0, const float* __restrict__ fw1,
                                                           const float* __restrict__ fw2, float dtOverDxSqr,
                                                           uint64_t* __restrict__ dof, float* __restrict__ diag,
                                                           float* __restrict__ xe, float* __restrict__ ye, float* __restrict__ ze) {
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    bool isDof = false;
    float dv = __fmul_rn(6.f, dtOverDxSqr), xv = -dtOverDxSqr, yv = -dtOverDxSqr, zv = -dtOverDxSqr;
    if (mask_get(phiMask, leaf, off)) {
        float phiHere = phi[i];
        if (phiHere < 0.f) {
            int3 o = t.origin[leaf];
            int g[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
            float diagonal = 0.f;
            float xyz[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int f = 0; f < 6; f++) {
                int comp = f >> 1;
                bool pos = (f & 1) == 0;
                const float* fw = comp == 0 ? fw0 : (comp == 1 ? fw1 : fw2);
                int c[3] = {g[0], g[1], g[2]};
                float weight, phiOther;
                if (pos) {
                    c[comp] += 1;
                    weight = grid_get(t, fw, 1.0f, c[0], c[1], c[2]);
                    phiOther = grid_get(t, phi, phiBg, c[0], c[1], c[2]);
                } else {
                    weight = fw[i];
                    c[comp] -= 1;
                    phiOther = grid_get(t, phi, phiBg, c[0], c[1], c[2]);
                }
                float term = __fmul_rn(weight, dtOverDxSqr);
                if (phiOther < 0.f) {
                    diagonal = __fadd_rn(diagonal, term);
                    if (!pos) xyz[comp] = -term;
                } else {
                    float theta = fraction_inside2(phiHere, phiOther);
                    if (theta < 0.02f) theta = 0.02f;
                    diagonal = __fadd_rn(diagonal, __fdiv_rn(term, theta));
                }
            }
            if (diagonal != 0.f) { isDof = true; dv = diagonal; xv = xyz[0]; yv = xyz[1]; zv = xyz[2]; }
        }
    }
    diag[i] = dv; xe[i] = xv; ye[i] = yv; ze[i] = zv;
    unsigned b = __ballot_sync(0xffffffffu, isDof);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<uint32_t*>(dof)[(size_t)leaf * 16 + (threadIdx.x >> 5)] = b;
}

// trimDefaultNodes (uaamg.cpp:1905-1960) + initInvDiagonal (:1222-1253). One CTA per leaf.
__global__ void __launch_bounds__(512) trim_kernel(const uint64_t* __restrict__ dof, float* __restrict__ diag,
                                                   float* __restrict__ xe, float* __restrict__ ye, float* __restrict__ ze,
                                                   float* __restrict__ invdiag, uint8_t* __restrict__ flags, float term) {
    __shared__ float sm16[16];
    __shared__ int sFlags;
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    bool on = mask_get(dof, leaf, off);
    const float defDiag = __fmul_rn(6.0f, term), defFace = -term;
    // epsilon = |default * 1e-5| (uaamg.cpp:1909-1916)
    const float epsD = fabsf(__fmul_rn(__fmul_rn(-6.0f, -term), 1e-5f)), epsF = fabsf(__fmul_rn(-term, 1e-5f));
    float* arr[4] = {diag, xe, ye, ze};
    int fl = 0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
        float def = a == 0 ? defDiag : defFace;
        float err = on ? fabsf(__fsub_rn(arr[a][i], def)) : 0.f;
        float m = block_absmax_512(err, sm16);
        if (threadIdx.x == 0) { if (m <= (a == 0 ? epsD : epsF)) fl |= (1 << a); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sFlags = fl; flags[leaf] = (uint8_t)fl; }
    __syncthreads();
    fl = sFlags;
#pragma unroll
    for (int a = 0; a < 4; a++)
        if (fl & (1 << a)) arr[a][i] = a == 0 ? defDiag : defFace;
    float inv;
    const float defInv = __fdiv_rn(1.0f, defDiag);
    if (fl & 1) inv = defInv;
    else if (on) { float d = diag[i]; inv = d == 0.f ? 0.f : __fdiv_rn(1.0f, d); }
    else inv = defInv;
    invdiag[i] = inv;
}

__global__ void dof_leaf_origins_kernel(TopoView t, const uint64_t* __restrict__ dof, int3* __restrict__ out,
                                        uint32_t* __restrict__ counter) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= t.n) return;
    uint64_t a = 0;
    for (int k = 0; k < 8; k++) a |= dof[(size_t)l * 8 + k];
    if (!a) return;
    int3 o = t.origin[l];
    // TouchCoarseLeafReducer: touchLeaf(origin / 2) (FF/SIMD_UAAMG_Ops.h:17-21); origins are multiples of 8
    out[atomicAdd(counter, 1u)] = make_int3((o.x / 2) & ~7, (o.y / 2) & ~7, (o.z / 2) & ~7);
}
__device__ __forceinline__ bool dof_on(const TopoView& t, const uint64_t* dof, int x, int y, int z) {
    int l = topo_find(t, x, y, z);
    return l >= 0 && mask_get(dof, l, voxel_off(x, y, z));
}
// initializeFromFineLevel (uaamg.cpp:436-616): coarse DOF mask + Galerkin coefficients
__global__ void __launch_bounds__(512) coarsen_kernel(LevelView F, TopoView ct, float cterm, uint64_t* __restrict__ cdof,
                                                      float* __restrict__ cdiag, float* __restrict__ cx,
                                                      float* __restrict__ cy, float* __restrict__ cz) {
    int leaf = blockIdx.x, off = threadIdx.x;
    int3 o = ct.origin[leaf];
    int gx = o.x + (off >> 6), gy = o.y + ((off >> 3) & 7), gz = o.z + (off & 7);
    float diag = 0.f, x = 0.f, y = 0.f, z = 0.f;
    bool any = false;
#pragma unroll
    for (int ii = 0; ii < 2; ii++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++)
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
                int fx = 2 * gx + ii, fy = 2 * gy + jj, fz = 2 * gz + kk;
                int fl = topo_find(F.t, fx, fy, fz);
                if (fl < 0) continue;
                int fo = voxel_off(fx, fy, fz);
                if (!mask_get(F.dof, fl, fo)) continue;
                any = true;
                size_t fi = (size_t)fl * LEAF + fo;
                diag = __fadd_rn(diag, F.diag[fi]);
                if (dof_on(F.t, F.dof, fx - 1, fy, fz)) {
                    if (ii == 0) x = __fadd_rn(x, F.xe[fi]); else diag = __fadd_rn(diag, __fmul_rn(2.f, F.xe[fi]));
                }
                if (dof_on(F.t, F.dof, fx, fy - 1, fz)) {
                    if (jj == 0) y = __fadd_rn(y, F.ye[fi]); else diag = __fadd_rn(diag, __fmul_rn(2.f, F.ye[fi]));
                }
                if (dof_on(F.t, F.dof, fx, fy, fz - 1)) {
                    if (kk == 0) z = __fadd_rn(z, F.ze[fi]); else diag = __fadd_rn(diag, __fmul_rn(2.f, F.ze[fi]));
                }
            }
    size_t i = (size_t)leaf * LEAF + off;
    const float factor = 0.5f * (1.0f / 8.0f);
    if (any) {
        cdiag[i] = __fmul_rn(diag, factor); cx[i] = __fmul_rn(x, factor); cy[i] = __fmul_rn(y, factor); cz[i] = __fmul_rn(z, factor);
    } else {
        cdiag[i] = __fmul_rn(6.0f, cterm); cx[i] = -cterm; cy[i] = -cterm; cz[i] = -cterm;
    }
    unsigned b = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<uint32_t*>(cdof)[(size_t)leaf * 16 + (threadIdx.x >> 5)] = b;
}

// BuildPoissonRhs (uaamg.cpp:44-86)
// the surface-tension terms of the right-hand side (enabled by SurfaceTension > 0): liquid SDF on the pool, curvature on its own topology
struct TensionArgs { int on; float tension, dtOverDxSqr; const float* phi; float phiBg; TopoView ct; const float* curv; float curvBg; };
__global__ void __launch_bounds__(512) rhs_kernel(TopoView t, const uint64_t* __restrict__ dof,
                                                  const float* __restrict__ fw0, const float* __restrict__ fw1, const float* __restrict__ fw2,
                                                  const float* __restrict__ v0, const float* __restrict__ v1, const float* __restrict__ v2,
                                                  const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                                                  float invdx, float* __restrict__ rhs, TensionArgs T) {
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    float r = 0.f;
    if (mask_get(dof, leaf, off)) {
        int3 o = t.origin[leaf];
        int g[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
        float weightSum = 0.f;
        bool nonZero = false;
        const float phiThis = T.on ? T.phi[i] : 0.f;
        const float curvThis = T.on ? grid_get(T.ct, T.curv, T.curvBg, g[0], g[1], g[2]) : 0.f;
#pragma unroll
        for (int f = 0; f < 6; f++) {
            int ch = f >> 1;
            bool pos = (f & 1) == 0;
            const float* fw = ch == 0 ? fw0 : (ch == 1 ? fw1 : fw2);
            const float* vv = ch == 0 ? v0 : (ch == 1 ? v1 : v2);
            const float* sv = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
            float weight, vel, svel;
            if (pos) {
                int c[3] = {g[0], g[1], g[2]};
                c[ch] += 1;
                int nl = topo_find(t, c[0], c[1], c[2]);
                if (nl >= 0) { size_t k = (size_t)nl * LEAF + voxel_off(c[0], c[1], c[2]); weight = fw[k]; vel = vv[k]; svel = sv[k]; }
                else { weight = 1.0f; vel = 0.f; svel = 0.f; }
            } else { weight = fw[i]; vel = vv[i]; svel = sv[i]; }
            weightSum = __fadd_rn(weightSum, weight);
            if (weight != 0.f) nonZero = true;
            float flux = __fmul_rn(invdx, __fadd_rn(__fmul_rn(weight, vel), __fmul_rn(__fsub_rn(1.0f, weight), svel)));
            if (pos) r = __fsub_rn(r, flux); else r = __fadd_rn(r, flux);
            if (T.on) {   // BuildPoissonRhs_withTension (uaamg.cpp:187-192): the cell across this face is air
                int c[3] = {g[0], g[1], g[2]};
                c[ch] += pos ? 1 : -1;
                const float phiOther = grid_get(t, T.phi, T.phiBg, c[0], c[1], c[2]);
                if (phiThis < 0.f && phiOther >= 0.f) {
                    const float curvOther = grid_get(T.ct, T.curv, T.curvBg, c[0], c[1], c[2]);
                    float theta = fraction_inside2(phiThis, phiOther);
                    if (theta < 0.02f) theta = 0.02f;
                    const float mix = __fadd_rn(__fmul_rn(theta, curvOther), __fmul_rn(__fsub_rn(1.f, theta), curvThis));
                    r = __fadd_rn(r, __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(T.dtOverDxSqr, weight), T.tension), mix), theta));
                }
            }
        }
        if (!nonZero || (double)weightSum < 0.1) r = 0.f;
    }
    rhs[i] = r;
}

// ---------------------------------------------------------------- stencil device functions
// Written once as per-leaf device functions and used by two kinds of kernels: one CTA per leaf for the
// large levels, and mg_bottom_kernel, a single resident CTA that runs the whole mu-cycle below a level
// (every sweep / residual / restriction / prolongation / the coarsest CG) without returning to the host.
struct Nbr { int xm, xp, ym, yp, zm, zp; };
// how the iterate x and the right-hand side b are read:
//   M_LEAF  one CTA per leaf, one kernel per pass: x plain, b through the read-only path
//   M_GRID  mg_cycle_kernel, many CTAs in one launch separated by grid barriers: x and b were written by other
//           SMs inside this launch, so they are read from L2 (ld.global.cg), never from this SM's L1
//   M_LOCAL one CTA owns the data for the whole launch (shared memory or its own L1): plain loads
//   M_GRID_L1  like M_GRID but with plain (L1-cached) loads. Sound because every op of the launch ends in grid_barrier(), whose
//           ld.acquire.gpu orders all later loads of the CTA after the other SMs' released stores (the hardware invalidates the
//           SM's L1 there), and inside a colour pass a thread reads only values no other thread writes in that pass (the six
//           neighbours have the other colour). What it buys: the seven x loads of a voxel touch the same few sectors; served by
//           L1 they cost one L2 sector read instead of up to seven (a level-0 pass moved ~84 MB through L2 -> SM).
enum { M_LEAF = 0, M_GRID = 1, M_LOCAL = 2, M_GRID_L1 = 3 };
// L2-coherent load (LDG.E.STRONG.GPU): never served from this SM's L1
__device__ __forceinline__ float ld_l2(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
template <int M> __device__ __forceinline__ float ldx(const float* p) { return M == M_GRID ? ld_l2(p) : *p; }
template <int M> __device__ __forceinline__ float ldb(const float* p) { return M == M_GRID ? ld_l2(p) : (M == M_LEAF ? __ldg(p) : *p); }
__device__ __forceinline__ LeafInfo load_info(const LevelView& L, int leaf) {
    const int4* p = reinterpret_cast<const int4*>(L.info + leaf);
    int4 a = __ldg(p), b = __ldg(p + 1);
    LeafInfo li;
    li.nb[0] = a.x; li.nb[1] = a.y; li.nb[2] = a.z; li.nb[3] = a.w; li.nb[4] = b.x; li.nb[5] = b.y;
    li.flags = (uint32_t)b.z;
    return li;
}
__device__ __forceinline__ Nbr nbr_of(const LeafInfo& li) { return Nbr{li.nb[0], li.nb[1], li.nb[2], li.nb[3], li.nb[4], li.nb[5]}; }
// off-diagonal sum with the reference's association (uaamg.cpp:1044-1049):
// ((x+ c_x+ + x- c_x-) + (y+ c_y+ + y- c_y-)) + (z+ c_z+ + z- c_z-)
template <bool CONST_COEF, int M>
__device__ __forceinline__ float offdiag(const LevelView& L, const float* x, int leaf, int off, const Nbr& nb) {
    // branch-free: every neighbour is (leaf slot, offset) picked with selects, a missing neighbour leaf reads
    // slot 0 and is masked afterwards, so the twelve loads issue back to back (the first version branched per
    // face and the loads sat behind reconvergence points)
    const int X = off >> 6, Y = (off >> 3) & 7, Z = off & 7;
    const float def = -L.term;
    const int lxp = X < 7 ? leaf : nb.xp, oxp = X < 7 ? off + 64 : off - 448;
    const int lxm = X > 0 ? leaf : nb.xm, oxm = X > 0 ? off - 64 : off + 448;
    const int lyp = Y < 7 ? leaf : nb.yp, oyp = Y < 7 ? off + 8 : off - 56;
    const int lym = Y > 0 ? leaf : nb.ym, oym = Y > 0 ? off - 8 : off + 56;
    const int lzp = Z < 7 ? leaf : nb.zp, ozp = Z < 7 ? off + 1 : off - 7;
    const int lzm = Z > 0 ? leaf : nb.zm, ozm = Z > 0 ? off - 1 : off + 7;
    const size_t ixp = (size_t)max(lxp, 0) * LEAF + oxp, ixm = (size_t)max(lxm, 0) * LEAF + oxm;
    const size_t iyp = (size_t)max(lyp, 0) * LEAF + oyp, iym = (size_t)max(lym, 0) * LEAF + oym;
    const size_t izp = (size_t)max(lzp, 0) * LEAF + ozp, izm = (size_t)max(lzm, 0) * LEAF + ozm;
    float xp = ldx<M>(&x[ixp]), xm = ldx<M>(&x[ixm]), yp = ldx<M>(&x[iyp]), ym = ldx<M>(&x[iym]), zp = ldx<M>(&x[izp]), zm = ldx<M>(&x[izm]);
    float cxp = def, cxm = def, cyp = def, cym = def, czp = def, czm = def;
    if (!CONST_COEF) {
        const size_t i = (size_t)leaf * LEAF + off;
        const float a = __ldg(&L.xe[ixp]), b2 = __ldg(&L.ye[iyp]), c2 = __ldg(&L.ze[izp]);
        cxm = __ldg(&L.xe[i]); cym = __ldg(&L.ye[i]); czm = __ldg(&L.ze[i]);
        cxp = lxp < 0 ? def : a; cyp = lyp < 0 ? def : b2; czp = lzp < 0 ? def : c2;
    }
    xp = lxp < 0 ? 0.f : xp; xm = lxm < 0 ? 0.f : xm;
    yp = lyp < 0 ? 0.f : yp; ym = lym < 0 ? 0.f : ym;
    zp = lzp < 0 ? 0.f : zp; zm = lzm < 0 ? 0.f : zm;
    float fx = __fadd_rn(__fmul_rn(xp, cxp), __fmul_rn(xm, cxm));
    float fy = __fadd_rn(__fmul_rn(yp, cyp), __fmul_rn(ym, cym));
    float fz = __fadd_rn(__fmul_rn(zp, czp), __fmul_rn(zm, czm));
    return __fadd_rn(__fadd_rn(fx, fy), fz);
}
__device__ __forceinline__ bool dof_bit(const LevelView& L, int leaf, int off) {
    return (__ldg(&L.info[leaf].mask[off >> 6]) >> (off & 63)) & 1ull;
}
// one colour of red-black SOR on one leaf, in place (uaamg.cpp:1109-1150):
// x <- fma(x, 1-w, ((b - off) * invdiag) * w). t in [0,256) owns one voxel of the colour;
// colour 0 = red = (x+y+z) even.
template <int M>
__device__ __forceinline__ void rbgs_leaf(const LevelView& L, float* x, const float* b, int leaf, int t, int colour,
                                          float w, float oneMinusW) {
    const int X = t >> 5, Y = (t >> 2) & 7;
    const int Z = ((t & 3) << 1) | ((X + Y + colour) & 1);
    const int off = (X << 6) | (Y << 3) | Z;
    const size_t i = (size_t)leaf * LEAF + off;
    const LeafInfo li = load_info(L, leaf);
    if (!(li.flags & LI_ANY)) return;   // uniform over the 256 threads of the leaf
    // everything below is issued unconditionally (the arrays cover every voxel of every leaf); only the store
    // is predicated on the DOF bit, so no load waits behind a branch
    const float bi = ldb<M>(&b[i]), inv = __ldg(&L.invdiag[i]), xi = ldx<M>(&x[i]);
    const bool on = dof_bit(L, leaf, off);
    const Nbr nb = nbr_of(li);
    const float od = (li.flags & LI_CONST) ? offdiag<true, M>(L, x, leaf, off, nb) : offdiag<false, M>(L, x, leaf, off, nb);
    const float tt = __fmul_rn(__fmul_rn(__fsub_rn(bi, od), inv), w);
    if (on) x[i] = __fmaf_rn(xi, oneMinusW, tt);
}
// The same pass with the leaf record read from shared memory (mg_cycle_kernel keeps the records of the chunks a CTA visits,
// which never change inside a launch): one L2 round trip per chunk (the operands) instead of two (record, then operands).
struct CachedInfo { int nb[6]; uint32_t flags, pad; uint64_t mask[8]; };   // the first 96 bytes of a LeafInfo
__device__ __forceinline__ LeafInfo cached_info(const CachedInfo& ci) {
    const int4 a = *reinterpret_cast<const int4*>(&ci.nb[0]);
    const int4 b = *reinterpret_cast<const int4*>(&ci.nb[4]);
    LeafInfo li;
    li.nb[0] = a.x; li.nb[1] = a.y; li.nb[2] = a.z; li.nb[3] = a.w; li.nb[4] = b.x; li.nb[5] = b.y;
    li.flags = (uint32_t)b.z;
    return li;
}
template <int M>
__device__ __forceinline__ void rbgs_leaf_cached(const LevelView& L, float* x, const float* b, int leaf, int t, int colour,
                                                 float w, float oneMinusW, const CachedInfo& ci, float invDefault) {
    const int X = t >> 5, Y = (t >> 2) & 7;
    const int Z = ((t & 3) << 1) | ((X + Y + colour) & 1);
    const int off = (X << 6) | (Y << 3) | Z;
    const size_t i = (size_t)leaf * LEAF + off;
    const LeafInfo li = cached_info(ci);
    if (!(li.flags & LI_ANY)) return;
    // a leaf whose diagonal is the default everywhere (LI_DIAG) holds 1 / (6 term) in every invdiag entry: not read
    const float bi = ldb<M>(&b[i]), inv = (li.flags & LI_DIAG) ? invDefault : __ldg(&L.invdiag[i]), xi = ldx<M>(&x[i]);
    const bool on = (ci.mask[X] >> (off & 63)) & 1ull;
    const Nbr nb = nbr_of(li);
    const float od = (li.flags & LI_CONST) ? offdiag<true, M>(L, x, leaf, off, nb) : offdiag<false, M>(L, x, leaf, off, nb);
    const float tt = __fmul_rn(__fmul_rn(__fsub_rn(bi, od), inv), w);
    if (on) x[i] = __fmaf_rn(xi, oneMinusW, tt);
}
// zero_red_leaf with the record in shared memory: t in [0,256) owns the z pair 2t, 2t + 1 (one red, one black voxel)
template <int M>
__device__ __forceinline__ void zero_red_pair_cached(const LevelView& L, float* x, const float* b, int leaf, int t, float w,
                                                     const CachedInfo& ci, float invDefault) {
    const int X = t >> 5, Y = (t >> 2) & 7, Zp = t & 3;
    const int h = (X + Y) & 1;                       // the red voxel of the pair: z = 2 Zp + h
    const size_t i = (size_t)leaf * LEAF + 2 * t;
    float v = 0.f;
    if ((ci.mask[X] >> ((Y << 3) | (Zp << 1) | h)) & 1ull) {
        const float inv = (ci.flags & LI_DIAG) ? invDefault : __ldg(&L.invdiag[i + h]);
        v = __fmul_rn(__fmul_rn(ldb<M>(&b[i + h]), inv), w);
    }
    *reinterpret_cast<float2*>(x + i) = h ? make_float2(0.f, v) : make_float2(v, 0.f);
}
// prolongation of one fine leaf with its record in shared memory: t in [0,256) owns the z-adjacent voxels 2t, 2t + 1, which
// share their coarse parent; the parent leaf and octant come from the record (leaf_parent_kernel). Arithmetic = prolong_voxel.
template <int M>
__device__ __forceinline__ void prolong_pair_cached(const LevelView& C, float* fine, const float* coarse, int leaf, int t, float alpha,
                                                    const CachedInfo& ci) {
    const int cl = (int)ci.pad;
    if (cl < 0 || !(ci.flags & LI_ANY)) return;
    const int X = t >> 5, Y = (t >> 2) & 7, Zp = t & 3;
    const unsigned bits = (unsigned)(ci.mask[X] >> ((Y << 3) | (Zp << 1))) & 3u;
    if (!bits) return;
    const uint32_t oct = ci.flags >> LI_OCT_SHIFT;
    const int co = ((((oct >> 2) & 1) * 4 + (X >> 1)) << 6) | ((((oct >> 1) & 1) * 4 + (Y >> 1)) << 3) | ((oct & 1) * 4 + Zp);
    if (!dof_bit(C, cl, co)) return;
    const float cv = __fmul_rn(alpha, ldx<M>(&coarse[(size_t)cl * LEAF + co]));
    float2* fp = reinterpret_cast<float2*>(fine + (size_t)leaf * LEAF + 2 * t);
    float2 f = *fp;
    if (bits & 1u) f.x = __fadd_rn(f.x, cv);
    if (bits & 2u) f.y = __fadd_rn(f.y, cv);
    *fp = f;
}
// residual + restriction of one fine leaf without shared memory or block barriers: the 256 threads of the leaf are laid out
// so that a 2 x 2 x 2 block of fine voxels sits in one warp (lane bit 4 = x parity, bit 2 = y parity, a thread owns the z pair);
// the thread with even x and y collects the eight residuals by shuffle and adds the active ones in the reference's (ii, jj, kk)
// order (uaamg.cpp:1773-1833). Arithmetic = the tile version above / residual_restrict_kernel.
template <int M>
__device__ __forceinline__ void resid_restrict_cached(const LevelView& F, const float* x, const float* b, bool bReadOnly, float* coarse,
                                                      int leaf, int t, const CachedInfo& ci) {
    const LeafInfo li = cached_info(ci);
    if (!(li.flags & LI_ANY)) return;   // uniform over the leaf's 256 threads (whole warps)
    const int wq = t >> 5, l = t & 31;
    const int X = ((wq >> 1) << 1) | (l >> 4), Y = ((wq & 1) << 2) | ((l >> 2) & 3), Zp = l & 3;
    const int off0 = (X << 6) | (Y << 3) | (Zp << 1);
    const size_t i0 = (size_t)leaf * LEAF + off0;
    const unsigned bits = (unsigned)(ci.mask[X] >> ((Y << 3) | (Zp << 1))) & 3u;
    const float2 bv = bReadOnly ? __ldg(reinterpret_cast<const float2*>(b + i0)) : *reinterpret_cast<const float2*>(b + i0);
    const Nbr nb = nbr_of(li);
    float r[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int off = off0 + h;
        const float od = (li.flags & LI_CONST) ? offdiag<true, M>(F, x, leaf, off, nb) : offdiag<false, M>(F, x, leaf, off, nb);
        const float dg = (li.flags & LI_DIAG) ? __fmul_rn(6.0f, F.term) : __ldg(&F.diag[i0 + h]);
        const float ax = __fmaf_rn(ldx<M>(&x[i0 + h]), dg, od);
        r[h] = ((bits >> h) & 1u) ? __fsub_rn(h == 0 ? bv.x : bv.y, ax) : 0.f;
    }
    // (ii, jj) = (0,0) own, (0,1) lane ^ 4, (1,0) lane ^ 16, (1,1) lane ^ 20
    float rr[4][2];
    unsigned bb[4];
    rr[0][0] = r[0]; rr[0][1] = r[1]; bb[0] = bits;
#pragma unroll
    for (int q = 1; q < 4; q++) {
        const int m = ((q & 1) ? 4 : 0) | ((q & 2) ? 16 : 0);
        rr[q][0] = __shfl_xor_sync(0xffffffffu, r[0], m);
        rr[q][1] = __shfl_xor_sync(0xffffffffu, r[1], m);
        bb[q] = __shfl_xor_sync(0xffffffffu, bits, m);
    }
    if ((l & 20) != 0) return;   // odd x or odd y: contributed through the shuffles
    float sum = 0.f;
    bool any = false;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
            if ((bb[q] >> kk) & 1u) { sum = __fadd_rn(sum, rr[q][kk]); any = true; }
    const int cl = (int)ci.pad;
    if (!any || cl < 0) return;
    const uint32_t oct = li.flags >> LI_OCT_SHIFT;
    const int co = ((((oct >> 2) & 1) * 4 + (X >> 1)) << 6) | ((((oct >> 1) & 1) * 4 + (Y >> 1)) << 3) | ((oct & 1) * 4 + Zp);
    coarse[(size_t)cl * LEAF + co] = __fmul_rn(sum, 0.125f);
}
// the first red pass of a sweep that starts from a zero guess (setGridToResultAfterFirstRBGS,
// uaamg.cpp:1665-1731): every neighbour is 0, so off = ((0*c + 0*c) + ...) = 0 and
// x_red = fma(0, 1-w, ((b - 0) * invdiag) * w); black voxels are set to 0. No neighbour traffic.
template <int M>
__device__ __forceinline__ void zero_red_leaf(const LevelView& L, float* x, const float* b, int leaf, int off, float w) {
    const size_t i = (size_t)leaf * LEAF + off;
    float v = 0.f;
    const int X = off >> 6, Y = (off >> 3) & 7, Z = off & 7;
    if (((X + Y + Z) & 1) == 0 && dof_bit(L, leaf, off)) v = __fmul_rn(__fmul_rn(ldb<M>(&b[i]), __ldg(&L.invdiag[i])), w);
    x[i] = v;
}
// A x on one voxel (uaamg.cpp:1085-1106)
template <int M>
__device__ __forceinline__ float ax_voxel(const LevelView& L, const LeafInfo& li, const float* x, int leaf, int off) {
    const Nbr nb = nbr_of(li);
    const float od = (li.flags & LI_CONST) ? offdiag<true, M>(L, x, leaf, off, nb) : offdiag<false, M>(L, x, leaf, off, nb);
    const size_t i = (size_t)leaf * LEAF + off;
    return __fmaf_rn(ldx<M>(&x[i]), __ldg(&L.diag[i]), od);
}
// restriction of one fine leaf's residual held in shared memory (uaamg.cpp:1773-1833): coarse = 1/8 sum of
// the active fine 2^3, same (ii,jj,kk) order. q in [0,64) is the coarse cell inside the fine leaf's octant.
__device__ __forceinline__ void restrict_leaf(const LevelView& F, const LevelView& C, const float* sres, int fineLeaf, int q,
                                              float* __restrict__ coarse) {
    const int cx = q >> 4, cy = (q >> 2) & 3, cz = q & 3;
    const int fb = (cx << 7) | (cy << 4) | (cz << 1);
    float sum = 0.f;
    bool any = false;
#pragma unroll
    for (int ii = 0; ii < 2; ii++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++)
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
                int fo = fb + 64 * ii + 8 * jj + kk;
                if (dof_bit(F, fineLeaf, fo)) { sum = __fadd_rn(sum, sres[fo]); any = true; }
            }
    if (!any) return;
    const int3 o = F.t.origin[fineLeaf];
    const int gx = (o.x >> 1) + cx, gy = (o.y >> 1) + cy, gz = (o.z >> 1) + cz;
    const int cl = topo_find(C.t, gx, gy, gz);
    if (cl < 0) return;
    coarse[(size_t)cl * LEAF + voxel_off(gx, gy, gz)] = __fmul_rn(sum, 0.125f);
}
// prolongation<inplace_add> (uaamg.cpp:1835-1903), gathered per fine voxel: fine += alpha * coarse(parent)
template <int M>
__device__ __forceinline__ void prolong_voxel(const LevelView& F, const LevelView& C, float* fine, const float* coarse, int leaf,
                                              int off, float alpha) {
    if (!dof_bit(F, leaf, off)) return;
    const int3 o = F.t.origin[leaf];
    const int cx = (o.x + (off >> 6)) >> 1, cy = (o.y + ((off >> 3) & 7)) >> 1, cz = (o.z + (off & 7)) >> 1;
    const int cl = topo_find(C.t, cx, cy, cz);
    if (cl < 0) return;
    const int co = voxel_off(cx, cy, cz);
    if (!dof_bit(C, cl, co)) return;
    const size_t i = (size_t)leaf * LEAF + off;
    fine[i] = __fadd_rn(ldx<M>(&fine[i]), __fmul_rn(alpha, ldx<M>(&coarse[(size_t)cl * LEAF + co])));
}

// the coarse leaf a fine leaf restricts into / prolongs from (the fine leaf is one octant of it): slot in LeafInfo::pad,
// octant (x, y, z half) in flags bits 8..10. One directory lookup per leaf and solve instead of one per voxel and transfer.
__global__ void leaf_parent_kernel(TopoView ft, TopoView ct, LeafInfo* __restrict__ info) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= ft.n) return;
    const int3 o = ft.origin[l];
    const int cx = o.x >> 1, cy = o.y >> 1, cz = o.z >> 1;
    info[l].pad = (uint32_t)topo_find(ct, cx, cy, cz);
    info[l].flags |= (uint32_t)((((cx >> 2) & 1) << 2) | (((cy >> 2) & 1) << 1) | ((cz >> 2) & 1)) << LI_OCT_SHIFT;
}
// ---------------------------------------------------------------- per-leaf kernels (large levels)
__global__ void leaf_info_kernel(TopoView t, const uint64_t* __restrict__ dof, const uint8_t* __restrict__ flags,
                                 LeafInfo* __restrict__ info) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= t.n) return;
    const int* nb = t.nbr27 + (size_t)l * 27;
    LeafInfo li;
    li.nb[0] = nb[NB_XM]; li.nb[1] = nb[NB_XP]; li.nb[2] = nb[NB_YM]; li.nb[3] = nb[NB_YP]; li.nb[4] = nb[NB_ZM]; li.nb[5] = nb[NB_ZP];
    uint64_t any = 0;
    for (int k = 0; k < 8; k++) any |= dof[(size_t)l * 8 + k];
    // constant path: own and upper neighbours' face leaves all read as the default
    bool c = (flags[l] & 14) == 14;
    if (li.nb[1] >= 0 && !(flags[li.nb[1]] & 2)) c = false;
    if (li.nb[3] >= 0 && !(flags[li.nb[3]] & 4)) c = false;
    if (li.nb[5] >= 0 && !(flags[li.nb[5]] & 8)) c = false;
    li.flags = (any ? LI_ANY : 0) | (c ? LI_CONST : 0) | ((flags[l] & 1) ? LI_DIAG : 0);
    li.pad = 0;
    for (int k = 0; k < 8; k++) li.mask[k] = dof[(size_t)l * 8 + k];
    for (int k = 0; k < 4; k++) li.pad2[k] = 0;
    info[l] = li;
}
__global__ void __launch_bounds__(256) rbgs_kernel(LevelView L, float* x, const float* __restrict__ b, int colour, float w, float oneMinusW) {
    rbgs_leaf<M_LEAF>(L, x, b, blockIdx.x, threadIdx.x, colour, w, oneMinusW);
}
// ---- one colour pass with a Z-ROW PER THREAD (the large levels of the tiled path)
// rbgs_kernel spends ~100 issued instructions on one update (six neighbour index selects, 64-bit address arithmetic, one 4-byte
// load each). Here a thread owns the row (X, Y) of its leaf: the row, its four x / y neighbour rows, b, invdiag and the
// coefficient rows arrive as 128-bit loads and feed the row's four updates of the colour. The 64 threads of a leaf are arranged
// so that a warp holds the rows of ONE parity of X + Y: which half of a row is being updated (P = first z) is then warp-uniform
// and selected by a branch, not by per-element selects. Same arithmetic and association as rbgs_leaf / offdiag.
struct Row8 { float v[8]; };
__device__ __forceinline__ Row8 row_ld(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    return Row8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ Row8 row_ldg(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    return Row8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ Row8 row_const(float c) { return Row8{{c, c, c, c, c, c, c, c}}; }
template <int P, bool CONST_COEF>
__device__ __forceinline__ void rbgs_row(const LevelView& L, float* x, const float* __restrict__ b, int leaf, int X, int Y, const LeafInfo& li,
                                         float w, float oneMinusW) {
    const int off = (X << 6) | (Y << 3);
    const size_t base = (size_t)leaf * LEAF + off;
    const float def = -L.term;
    const int lxp = X < 7 ? leaf : li.nb[1], lxm = X > 0 ? leaf : li.nb[0], lyp = Y < 7 ? leaf : li.nb[3], lym = Y > 0 ? leaf : li.nb[2];
    const int oxp = X < 7 ? off + 64 : off - 448, oxm = X > 0 ? off - 64 : off + 448, oyp = Y < 7 ? off + 8 : off - 56, oym = Y > 0 ? off - 8 : off + 56;
    // (a missing neighbour leaf reads slot 0 and is masked afterwards: every load is issued unconditionally)
    Row8 own = row_ld(x + base);
    Row8 xp = row_ld(x + (size_t)max(lxp, 0) * LEAF + oxp), xm = row_ld(x + (size_t)max(lxm, 0) * LEAF + oxm);
    Row8 yp = row_ld(x + (size_t)max(lyp, 0) * LEAF + oyp), ym = row_ld(x + (size_t)max(lym, 0) * LEAF + oym);
    const int lz = P == 0 ? li.nb[4] : li.nb[5];
    float zh = x[(size_t)max(lz, 0) * LEAF + off + (P == 0 ? 7 : 0)];   // the one z neighbour outside the row
    const Row8 bb = row_ldg(b + base);
    const uint32_t dof = (uint32_t)(__ldg(&L.info[leaf].mask[X]) >> (Y << 3)) & 0xffu;
    Row8 inv, cxp, cxm, cyp, cym, cz;
    float cz8 = def;
    if (CONST_COEF) { cxp = cxm = cyp = cym = cz = row_const(def); }
    else {
        cxm = row_ldg(L.xe + base); cym = row_ldg(L.ye + base); cz = row_ldg(L.ze + base);
        cxp = row_ldg(L.xe + (size_t)max(lxp, 0) * LEAF + oxp); cyp = row_ldg(L.ye + (size_t)max(lyp, 0) * LEAF + oyp);
        if (P == 1) cz8 = __ldg(&L.ze[(size_t)max(lz, 0) * LEAF + off]);
        if (lxp < 0) cxp = row_const(def);
        if (lyp < 0) cyp = row_const(def);
        if (lz < 0) cz8 = def;
    }
    if (li.flags & LI_DIAG) inv = row_const(__fdiv_rn(1.0f, __fmul_rn(6.0f, L.term)));
    else inv = row_ldg(L.invdiag + base);
    if (lxp < 0) xp = row_const(0.f);
    if (lxm < 0) xm = row_const(0.f);
    if (lyp < 0) yp = row_const(0.f);
    if (lym < 0) ym = row_const(0.f);
    if (lz < 0) zh = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int Z = 2 * k + P;
        const float zpv = Z < 7 ? own.v[Z < 7 ? Z + 1 : 7] : zh, zmv = Z > 0 ? own.v[Z > 0 ? Z - 1 : 0] : zh;
        const float czp = Z < 7 ? cz.v[Z < 7 ? Z + 1 : 7] : cz8, czm = cz.v[Z];
        const float fx = __fadd_rn(__fmul_rn(xp.v[Z], cxp.v[Z]), __fmul_rn(xm.v[Z], cxm.v[Z]));
        const float fy = __fadd_rn(__fmul_rn(yp.v[Z], cyp.v[Z]), __fmul_rn(ym.v[Z], cym.v[Z]));
        const float fz = __fadd_rn(__fmul_rn(zpv, czp), __fmul_rn(zmv, czm));
        const float od = __fadd_rn(__fadd_rn(fx, fy), fz);
        const float tt = __fmul_rn(__fmul_rn(__fsub_rn(bb.v[Z], od), inv.v[Z]), w);
        if ((dof >> Z) & 1u) x[base + Z] = __fmaf_rn(own.v[Z], oneMinusW, tt);
    }
}
__global__ void __launch_bounds__(256) rbgs_rows_kernel(LevelView L, float* x, const float* __restrict__ b, int colour, float w, float oneMinusW) {
    const int leaf = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (leaf >= L.t.n) return;
    const LeafInfo li = load_info(L, leaf);
    if (!(li.flags & LI_ANY)) return;
    const int l = threadIdx.x & 31, parity = (threadIdx.x >> 5) & 1;
    const int X = l >> 2, Y = ((l & 3) << 1) | ((X + parity) & 1);
    const bool cc = (li.flags & LI_CONST) != 0;
    if ((parity ^ colour) == 0) {
        if (cc) rbgs_row<0, true>(L, x, b, leaf, X, Y, li, w, oneMinusW); else rbgs_row<0, false>(L, x, b, leaf, X, Y, li, w, oneMinusW);
    } else {
        if (cc) rbgs_row<1, true>(L, x, b, leaf, X, Y, li, w, oneMinusW); else rbgs_row<1, false>(L, x, b, leaf, X, Y, li, w, oneMinusW);
    }
}
__global__ void __launch_bounds__(512) zero_red_kernel(LevelView L, float* __restrict__ x, const float* __restrict__ b, float w) {
    zero_red_leaf<M_LEAF>(L, x, b, blockIdx.x, threadIdx.x, w);
}
// r = b - A x on a fine leaf, restricted straight into the coarse right-hand side (no residual round trip)
__global__ void __launch_bounds__(512) residual_restrict_kernel(LevelView F, LevelView C, const float* x, const float* __restrict__ b,
                                                                float* __restrict__ coarse, int leaf0) {
    __shared__ float sres[LEAF];
    const int leaf = blockIdx.x + leaf0, off = threadIdx.x;
    const LeafInfo li = load_info(F, leaf);
    if (!(li.flags & LI_ANY)) return;
    float r = 0.f;
    if (dof_bit(F, leaf, off)) r = __fsub_rn(b[(size_t)leaf * LEAF + off], ax_voxel<M_LEAF>(F, li, x, leaf, off));
    sres[off] = r;
    __syncthreads();
    if (off < 64) restrict_leaf(F, C, sres, leaf, off, coarse);
}
__global__ void __launch_bounds__(512) prolong_kernel(LevelView F, LevelView C, float* fine, const float* __restrict__ coarse, float alpha) {
    prolong_voxel<M_LEAF>(F, C, fine, coarse, blockIdx.x, threadIdx.x, alpha);
}

// ---------------------------------------------------------------- level-0 PCG kernels with fused reductions
// scalars live on the device: s[0]=rho s[1]=sigma s[2]=alpha s[3]=beta s[4]=nu s[5]=rho_new.
// Every producing kernel writes one partial per leaf; the LAST CTA to finish (threadfence + counter) folds the
// partials in index order, so the result does not depend on which CTA that is, and applies the scalar update
// that follows the reduction in solveMultigridPCG (uaamg.cpp:2332-2403). No separate fold / scalar launches.
// FIN_DEFER (slab decomposition): the fold only stores this rank's partial in s[6]; after the all-reduce of s[6]
// scalar_fin_kernel applies the same update.
enum { FIN_NU = 0, FIN_SIGMA_ALPHA = 1, FIN_RHO_INIT = 2, FIN_RHO_BETA = 3, FIN_DEFER = 8 };
__device__ __forceinline__ void apply_fin(float* s, int fin, float r) {
    if (fin == FIN_NU) s[4] = r;
    else if (fin == FIN_SIGMA_ALPHA) { s[1] = r; s[2] = __fdiv_rn(s[0], r); }
    else if (fin == FIN_RHO_INIT) s[0] = r;
    else { s[5] = r; s[3] = __fdiv_rn(r, s[0]); s[0] = r; }
}
__global__ void scalar_fin_kernel(float* s, int fin) { apply_fin(s, fin, s[6]); }
// The per-leaf partials are already in partial[0, nPartials); the last CTA to get here folds them in index order.
// (One arrival per CTA on ONE counter: with a CTA per leaf that was ~5800 same-address atomics, ~25 us of every
// reduction kernel at 16.8 M particles; the kernels now stride over the leaves with a few hundred CTAs.)
template <bool IS_MAX>
__device__ __forceinline__ void finish_reduction(float* partial, int nPartials, unsigned* counter, float* s, int fin, float* sm16) {
    __shared__ bool sLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        sLast = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    float a = 0.f;
    for (int i = threadIdx.x; i < nPartials; i += blockDim.x) {
        float v = __ldcg(&partial[i]);
        if (IS_MAX) a = (isfinite(a) ? (isfinite(v) ? fmaxf(a, v) : v) : a);
        else a = __fadd_rn(a, v);
    }
    float r = IS_MAX ? block_absmax_512(a, sm16) : block_sum_512(a, sm16);
    if (threadIdx.x == 0) {
        if (fin & FIN_DEFER) s[6] = r;
        else apply_fin(s, fin, r);
        *counter = 0;
    }
}
constexpr int RED_GRID = 4 * 148;   // CTAs of the level-0 reduction kernels
// Every thread folds its own voxel over the leaves its CTA visits (ascending), ONE block reduction per CTA at the end, the last CTA
// folds the gridDim.x CTA partials in index order: a fixed order for a given leaf count (the reference's own order is whatever TBB
// does, FF/openvdb_grid_math_op.h:63-70). Round 1 reduced every leaf on its own -- two bar.sync per leaf, ten leaves per CTA in a row,
// no loads in flight across them (pcg_laplacian_dot 50 us against 22 us for two colour passes over the same data).
__device__ __forceinline__ float absmax_acc(float a, float v) {   // non-finite values propagate, as in block_absmax_512
    const float av = isfinite(v) ? fabsf(v) : v;
    return isfinite(a) ? (isfinite(av) ? fmaxf(a, av) : av) : a;
}
enum { MODE_LAPLACIAN = 0, MODE_RESIDUAL = 1 };
// y = A x (+ sigma = x.y, alpha) or y = b - A x (+ nu = |y|_inf)   (uaamg.cpp:1085-1106)
template <int MODE>
__global__ void __launch_bounds__(512) apply_kernel(LevelView L, const float* x, const float* __restrict__ b, float* __restrict__ y,
                                                    float* partial, unsigned* counter, float* s, int defer) {
    __shared__ float sm16[16];
    const int off = threadIdx.x;
    float acc = 0.f;
    for (int leaf = blockIdx.x; leaf < L.t.n; leaf += gridDim.x) {
        const LeafInfo li = load_info(L, leaf);
        float red = 0.f;
        if (li.flags & LI_ANY) {  // the reference skips empty rows; vectors stay zero there
            float out = 0.f;
            if (dof_bit(L, leaf, off)) {
                const size_t i = (size_t)leaf * LEAF + off;
                float ax = ax_voxel<M_LEAF>(L, li, x, leaf, off);
                if (MODE == MODE_RESIDUAL) { out = __fsub_rn(b[i], ax); red = out; }
                else { out = ax; red = __fmul_rn(x[i], ax); }
            }
            y[(size_t)leaf * LEAF + off] = out;
        }
        if (!owned_leaf(L, leaf)) red = 0.f;
        acc = MODE == MODE_RESIDUAL ? absmax_acc(acc, red) : __fadd_rn(acc, red);
    }
    const float r = MODE == MODE_RESIDUAL ? block_absmax_512(acc, sm16) : block_sum_512(acc, sm16);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
    finish_reduction<MODE == MODE_RESIDUAL>(partial, gridDim.x, counter, s, (MODE == MODE_RESIDUAL ? FIN_NU : FIN_SIGMA_ALPHA) | defer, sm16);
}
// r -= alpha z ; nu = |r|_inf (levelAlphaXPlusY + levelAbsMax, uaamg.cpp:2447-2479)
__global__ void __launch_bounds__(512) axpy_absmax_kernel(LevelView L, float* s, const float* __restrict__ z, float* __restrict__ r,
                                                          float* partial, unsigned* counter, int defer) {
    __shared__ float sm16[16];
    const int off = threadIdx.x;
    const float alpha = s[2];
    float acc = 0.f;
    for (int leaf = blockIdx.x; leaf < L.t.n; leaf += gridDim.x) {
        const size_t i = (size_t)leaf * LEAF + off;
        float v = 0.f;
        if (dof_bit(L, leaf, off)) { v = __fadd_rn(r[i], __fmul_rn(-alpha, z[i])); r[i] = v; }
        if (!owned_leaf(L, leaf)) v = 0.f;
        acc = absmax_acc(acc, v);
    }
    const float m = block_absmax_512(acc, sm16);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
    finish_reduction<true>(partial, gridDim.x, counter, s, FIN_NU | defer, sm16);
}
__global__ void __launch_bounds__(512) dot_kernel(LevelView L, const float* __restrict__ a, const float* __restrict__ b, float* partial,
                                                  unsigned* counter, float* s, int fin) {
    __shared__ float sm16[16];
    const int off = threadIdx.x;
    float acc = 0.f;
    for (int leaf = blockIdx.x; leaf < L.t.n; leaf += gridDim.x) {
        const size_t i = (size_t)leaf * LEAF + off;
        const float v = (dof_bit(L, leaf, off) && owned_leaf(L, leaf)) ? __fmul_rn(a[i], b[i]) : 0.f;
        acc = __fadd_rn(acc, v);
    }
    const float m = block_sum_512(acc, sm16);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
    finish_reduction<false>(partial, gridDim.x, counter, s, fin, sm16);
}
// x += alpha p ; p = z + beta p   (uaamg.cpp:2396-2397); final=1: only the x update (:2375)
__global__ void __launch_bounds__(512) update_kernel(LevelView L, const float* __restrict__ s, float* __restrict__ x, float* __restrict__ p,
                                                     const float* __restrict__ z, int final) {
    const int leaf = blockIdx.x, off = threadIdx.x;
    if (!dof_bit(L, leaf, off)) return;
    const size_t i = (size_t)leaf * LEAF + off;
    float pv = p[i];
    x[i] = __fadd_rn(x[i], __fmul_rn(s[2], pv));
    if (!final) p[i] = __fadd_rn(z[i], __fmul_rn(s[3], pv));
}

// ---------------------------------------------------------------- coarsest level
// Compact ELL form of the coarsest matrix (getTriplets, uaamg.cpp:278-355) and a single-CTA
// Jacobi-preconditioned CG, <= 10 iterations, tolerance float epsilon, zero initial guess
// (Eigen::ConjugateGradient defaults, uaamg.cpp:2291-2303,2019-2023; Eigen itself is not in
// the reference tree -> this follows Eigen's published algorithm, parity unpinned at the bit level).
__global__ void __launch_bounds__(512) ell_build_kernel(LevelView L, const uint32_t* __restrict__ leafStart,
                                                        int ndofPad, int* __restrict__ cols, float* __restrict__ vals,
                                                        int* __restrict__ rowOfVoxel) {
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    const uint64_t* m = L.dof + (size_t)leaf * 8;
    bool on = (m[off >> 6] >> (off & 63)) & 1ull;
    // row index = leaf prefix + number of DOF bits below off
    int below = 0;
    for (int w = 0; w < (off >> 6); w++) below += __popcll(m[w]);
    below += __popcll(m[off >> 6] & ((1ull << (off & 63)) - 1ull));
    int row = (int)leafStart[leaf] + below;
    rowOfVoxel[i] = on ? row : -1;
    if (!on) return;
    int3 o = L.t.origin[leaf];
    int g[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
    cols[row] = row;
    vals[row] = L.diag[i];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float* e = ch == 0 ? L.xe : (ch == 1 ? L.ye : L.ze);
        int nc[3] = {g[0], g[1], g[2]}, pc[3] = {g[0], g[1], g[2]};
        nc[ch] -= 1; pc[ch] += 1;
        int k0 = (1 + 2 * ch) * ndofPad + row, k1 = (2 + 2 * ch) * ndofPad + row;
        cols[k0] = -1; vals[k0] = 0.f; cols[k1] = -1; vals[k1] = 0.f;
        int nl = topo_find(L.t, nc[0], nc[1], nc[2]);
        if (nl >= 0 && mask_get(L.dof, nl, voxel_off(nc[0], nc[1], nc[2]))) {
            cols[k0] = -2 - (int)((size_t)nl * LEAF + voxel_off(nc[0], nc[1], nc[2]));  // resolved to a row below
            vals[k0] = e[i];
        }
        int pl = topo_find(L.t, pc[0], pc[1], pc[2]);
        if (pl >= 0 && mask_get(L.dof, pl, voxel_off(pc[0], pc[1], pc[2]))) {
            size_t pi = (size_t)pl * LEAF + voxel_off(pc[0], pc[1], pc[2]);
            cols[k1] = -2 - (int)pi;
            vals[k1] = e[pi];
        }
    }
}
__global__ void ell_resolve_kernel(int* __restrict__ cols, int total, const int* __restrict__ rowOfVoxel) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = cols[i];
    if (c <= -2) cols[i] = rowOfVoxel[-2 - c];
}
constexpr int BOT_THREADS = 1024;
__device__ __forceinline__ float cg_block_sum(float v, float* sm33) {
    for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();  // protect sm33 reuse
    if ((threadIdx.x & 31) == 0) sm33[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = threadIdx.x < 32 ? sm33[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
        for (int d = 16; d > 0; d >>= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, d));
        if (threadIdx.x == 0) sm33[32] = r;
    }
    __syncthreads();
    return sm33[32];
}
struct CoarseELL {
    int ndof, ndofPad, nVoxels;
    int inSmem;   // mg_cycle_kernel: the matrix is staged in CTA 0's shared memory (coarse_cg_smem)
    const int* cols; const float* vals; const int* rowOfVoxel;
};
// one CTA of BOT_THREADS threads; sm = 5*ndofPad floats of dynamic shared memory
__device__ void coarse_cg(const CoarseELL& E, const float* rhsGrid, float* lhsGrid, float* sm, float* sm33) {
    const int ndof = E.ndof, ndofPad = E.ndofPad;
    float* X = sm; float* R = X + ndofPad; float* P = R + ndofPad; float* T = P + ndofPad; float* DI = T + ndofPad;
    const int tid = threadIdx.x;
    for (int v = tid; v < E.nVoxels; v += BOT_THREADS) { int r = E.rowOfVoxel[v]; if (r >= 0) R[r] = rhsGrid[v]; }
    for (int r = tid; r < ndof; r += BOT_THREADS) { X[r] = 0.f; float d = E.vals[r]; DI[r] = d != 0.f ? __fdiv_rn(1.0f, d) : 1.0f; }
    __syncthreads();
    float acc = 0.f;
    for (int r = tid; r < ndof; r += BOT_THREADS) acc = __fadd_rn(acc, __fmul_rn(R[r], R[r]));
    float rhsNorm2 = cg_block_sum(acc, sm33);
    if (rhsNorm2 != 0.f) {
        const float tol = 1.1920929e-07f;
        float threshold = fmaxf(__fmul_rn(__fmul_rn(tol, tol), rhsNorm2), 1.17549435e-38f);
        float residualNorm2 = rhsNorm2;
        if (!(residualNorm2 < threshold)) {
            acc = 0.f;
            for (int r = tid; r < ndof; r += BOT_THREADS) { float pv = __fmul_rn(DI[r], R[r]); P[r] = pv; acc = __fadd_rn(acc, __fmul_rn(R[r], pv)); }
            float absNew = cg_block_sum(acc, sm33);
            for (int it = 0; it < 10; it++) {
                acc = 0.f;
                for (int r = tid; r < ndof; r += BOT_THREADS) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < 7; k++) { int c = E.cols[k * ndofPad + r]; if (c >= 0) s = __fadd_rn(s, __fmul_rn(E.vals[k * ndofPad + r], P[c])); }
                    T[r] = s;
                    acc = __fadd_rn(acc, __fmul_rn(P[r], s));
                }
                float pt = cg_block_sum(acc, sm33);
                float alpha = __fdiv_rn(absNew, pt);
                acc = 0.f;
                for (int r = tid; r < ndof; r += BOT_THREADS) {
                    X[r] = __fadd_rn(X[r], __fmul_rn(alpha, P[r]));
                    float rv = __fsub_rn(R[r], __fmul_rn(alpha, T[r]));
                    R[r] = rv;
                    acc = __fadd_rn(acc, __fmul_rn(rv, rv));
                }
                residualNorm2 = cg_block_sum(acc, sm33);
                if (residualNorm2 < threshold) break;
                acc = 0.f;
                for (int r = tid; r < ndof; r += BOT_THREADS) { float zv = __fmul_rn(DI[r], R[r]); T[r] = zv; acc = __fadd_rn(acc, __fmul_rn(R[r], zv)); }
                float absOld = absNew;
                absNew = cg_block_sum(acc, sm33);
                float beta = __fdiv_rn(absNew, absOld);
                for (int r = tid; r < ndof; r += BOT_THREADS) P[r] = __fadd_rn(T[r], __fmul_rn(beta, P[r]));
                __syncthreads();
            }
        }
    }
    __syncthreads();
    for (int v = tid; v < E.nVoxels; v += BOT_THREADS) { int r = E.rowOfVoxel[v]; if (r >= 0) lhsGrid[v] = X[r]; }
    __syncthreads();
}

// ---------------------------------------------------------------- the resident bottom of the mu-cycle
// The W-like cycle (mu = 2) visits level l 2^l times; below the first few levels every visit is a handful of
// leaves and the work is pure launch latency. mg_bottom_kernel executes the complete recursion below level
// `first` as a flat op list on ONE resident CTA (1024 threads, block-level barriers between dependent passes,
// data in L1/L2, the coarsest CG in shared memory) -- one launch per visit of level `first`.
enum { OP_ZERO_RED = 0, OP_RED = 1, OP_BLACK = 2, OP_RESID_RESTRICT = 3, OP_PROLONG = 4, OP_COARSE = 5 };
constexpr int BOT_MAX_LEVELS = 6;
constexpr int BOT_MAX_OPS = 400;
struct BottomLevel { LevelView v; float* x; float* b; int n; int xoff, boff; int bReadOnly; int hasParent; };  // offsets (floats) into dynamic smem, -1 = global
struct BottomParams {
    BottomLevel lv[BOT_MAX_LEVELS];
    CoarseELL ell;
    int nOps, nLevels;
    int loadX;                    // the first level's x holds a guess that must be read (no ZERO_RED first)
    float w, oneMinusW, prolongAlpha;
    uint8_t op[BOT_MAX_OPS];      // low 3 bits: op, high bits: level index inside the bottom
};
__global__ void __launch_bounds__(BOT_THREADS) mg_bottom_kernel(const __grid_constant__ BottomParams P) {
    extern __shared__ float dynsm[];
    __shared__ float sm33[33];
    __shared__ float sres[2 * LEAF];
    const int tid = threadIdx.x;
    // stage the first level's iterate in shared memory (x of every level and b of the lower levels live there
    // for the whole cycle; only coefficients, masks and the first level's b are read from global memory)
    if (P.lv[0].xoff >= 0 && P.loadX) {
        float* xs = dynsm + P.lv[0].xoff;
        for (int i = tid; i < P.lv[0].n * LEAF; i += BOT_THREADS) xs[i] = P.lv[0].x[i];
        __syncthreads();
    }
    for (int k = 0; k < P.nOps; k++) {
        const int code = P.op[k] & 7, li = P.op[k] >> 3;
        const BottomLevel& B = P.lv[li];
        float* x = B.xoff >= 0 ? dynsm + B.xoff : B.x;
        const float* b = B.boff >= 0 ? dynsm + B.boff : B.b;
        if (code == OP_ZERO_RED) {
            for (int base = 0; base < B.n; base += 2) { int leaf = base + (tid >> 9); if (leaf < B.n) zero_red_leaf<M_LOCAL>(B.v, x, b, leaf, tid & 511, P.w); }
        } else if (code == OP_RED || code == OP_BLACK) {
#pragma unroll 2
            for (int base = 0; base < B.n; base += 4) { int leaf = base + (tid >> 8); if (leaf < B.n) rbgs_leaf<M_LOCAL>(B.v, x, b, leaf, tid & 255, code == OP_RED ? 0 : 1, P.w, P.oneMinusW); }
        } else if (code == OP_RESID_RESTRICT) {
            const BottomLevel& C = P.lv[li + 1];
            float* cb = C.boff >= 0 ? dynsm + C.boff : C.b;
            for (int base = 0; base < B.n; base += 2) {
                int leaf = base + (tid >> 9), off = tid & 511;
                bool live = false;
                if (leaf < B.n) {
                    const LeafInfo info = load_info(B.v, leaf);
                    live = (info.flags & LI_ANY) != 0;
                    float r = 0.f;
                    if (live && dof_bit(B.v, leaf, off)) r = __fsub_rn(b[(size_t)leaf * LEAF + off], ax_voxel<M_LOCAL>(B.v, info, x, leaf, off));
                    sres[tid] = r;
                }
                __syncthreads();
                if (live && off < 64) restrict_leaf(B.v, C.v, sres + (tid >> 9) * LEAF, leaf, off, cb);
                __syncthreads();
            }
        } else if (code == OP_PROLONG) {
            const BottomLevel& C = P.lv[li + 1];
            const float* cx = C.xoff >= 0 ? dynsm + C.xoff : C.x;
            for (int base = 0; base < B.n; base += 2) { int leaf = base + (tid >> 9); if (leaf < B.n) prolong_voxel<M_LOCAL>(B.v, C.v, x, cx, leaf, tid & 511, P.prolongAlpha); }
        } else {
            coarse_cg(P.ell, b, x, dynsm, sm33);
        }
        __syncthreads();
    }
    if (P.lv[0].xoff >= 0) {
        const float* xs = dynsm + P.lv[0].xoff;
        for (int i = tid; i < P.lv[0].n * LEAF; i += BOT_THREADS) P.lv[0].x[i] = xs[i];
    }
}
// ---------------------------------------------------------------- shared-memory leaf tiles (mg_cycle_kernel)
// In the one-launch cycle the iterate is written by other SMs between passes, so it cannot be read through L1.
// Reading the 7-point stencil straight from L2 fetches every line six times (measured: 1.4 us per 4-leaf chunk,
// SM<->L2 bandwidth bound). Instead each group of 256 threads stages its leaf plus the six neighbour faces once,
// coalesced, in a padded 10x10x10 tile, and the stencil reads shared memory.
constexpr int TILE = 1000;
__device__ __forceinline__ int tidx(int x, int y, int z) { return (x + 1) * 100 + (y + 1) * 10 + (z + 1); }
__device__ __forceinline__ float2 ld_l2_f2(const float* p) {
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// t in [0,256): loads the leaf (two z-adjacent values per thread) and 6 x 64 face values (1.5 per thread)
__device__ __forceinline__ void tile_load(const float* x, int leaf, const LeafInfo& li, float* T, int t) {
    {
        const float2 v = ld_l2_f2(x + (size_t)leaf * LEAF + 2 * t);
        const int X = t >> 5, Y = (t >> 2) & 7, Z = (t & 3) << 1;
        T[tidx(X, Y, Z)] = v.x; T[tidx(X, Y, Z + 1)] = v.y;
    }
    const int f = t >> 6, q = t & 63, a = q >> 3, b = q & 7;   // face f, cell (a,b) of the face
    // round 1: faces x-, x+, y-, y+ ; round 2 (threads 0..127): z-, z+
    {
        int nb, off, ti;
        if (f == 0) { nb = li.nb[0]; off = 448 + q; ti = tidx(-1, a, b); }
        else if (f == 1) { nb = li.nb[1]; off = q; ti = tidx(8, a, b); }
        else if (f == 2) { nb = li.nb[2]; off = (a << 6) + 56 + b; ti = tidx(a, -1, b); }
        else { nb = li.nb[3]; off = (a << 6) + b; ti = tidx(a, 8, b); }
        T[ti] = nb >= 0 ? ld_l2(x + (size_t)nb * LEAF + off) : 0.f;
    }
    if (f < 2) {
        int nb, off, ti;
        if (f == 0) { nb = li.nb[4]; off = (a << 6) + (b << 3) + 7; ti = tidx(a, b, -1); }
        else { nb = li.nb[5]; off = (a << 6) + (b << 3); ti = tidx(a, b, 8); }
        T[ti] = nb >= 0 ? ld_l2(x + (size_t)nb * LEAF + off) : 0.f;
    }
}
// off-diagonal sum from a tile, same association as offdiag()
template <bool CONST_COEF>
__device__ __forceinline__ float offdiag_tile(const LevelView& L, const float* T, int leaf, int X, int Y, int Z, const LeafInfo& li) {
    const int off = (X << 6) | (Y << 3) | Z;
    const size_t base = (size_t)leaf * LEAF;
    const float def = -L.term;
    float cxp, cxm, cyp, cym, czp, czm;
    if (CONST_COEF) { cxp = cxm = cyp = cym = czp = czm = def; }
    else {
        cxp = X < 7 ? __ldg(&L.xe[base + off + 64]) : (li.nb[1] < 0 ? def : __ldg(&L.xe[(size_t)li.nb[1] * LEAF + off - 448]));
        cxm = __ldg(&L.xe[base + off]);
        cyp = Y < 7 ? __ldg(&L.ye[base + off + 8]) : (li.nb[3] < 0 ? def : __ldg(&L.ye[(size_t)li.nb[3] * LEAF + off - 56]));
        cym = __ldg(&L.ye[base + off]);
        czp = Z < 7 ? __ldg(&L.ze[base + off + 1]) : (li.nb[5] < 0 ? def : __ldg(&L.ze[(size_t)li.nb[5] * LEAF + off - 7]));
        czm = __ldg(&L.ze[base + off]);
    }
    const int c = tidx(X, Y, Z);
    float fx = __fadd_rn(__fmul_rn(T[c + 100], cxp), __fmul_rn(T[c - 100], cxm));
    float fy = __fadd_rn(__fmul_rn(T[c + 10], cyp), __fmul_rn(T[c - 10], cym));
    float fz = __fadd_rn(__fmul_rn(T[c + 1], czp), __fmul_rn(T[c - 1], czm));
    return __fadd_rn(__fadd_rn(fx, fy), fz);
}

// ---------------------------------------------------------------- compact rows for the bottom of the cycle
// Below the first few levels a leaf holds a few hundred DOFs at most, and the mu = 2 cycle visits those levels
// 8-16 times per application. They are stored as COMPACT ROWS (one row per DOF, 16-bit neighbour rows, the three
// "minus" face coefficients; the coefficient towards a "plus" neighbour is that neighbour's "minus" coefficient)
// that stay in CTA 0's shared memory for the whole launch: a colour pass is one shared-memory sweep + one
// bar.sync instead of an L2 round trip + a device-wide barrier. Row n (the first padding row) is a zero dummy:
// x = 0 and every coefficient 0, so a missing / non-DOF neighbour contributes 0 * 0 where the leaf form
// contributes 0 * c -- the same value in the same place of the reference's association.
struct BlobLayout { int diag, inv, minus, cols, parent, child, bytes; };
__host__ __device__ inline BlobLayout blob_layout(int np, bool hasChild) {
    BlobLayout b;
    int o = 0;
    b.diag = o; o += 4 * np;
    b.inv = o; o += 4 * np;
    b.minus = o; o += 12 * np;
    b.cols = o; o += 12 * np;     // u16 [6][np]: x-, x+, y-, y+, z-, z+
    b.parent = o; o += 2 * np;    // u16 [np]: row in the next coarser compact level
    b.child = o; if (hasChild) o += 16 * np;  // u16 [8][np]: rows of the 2^3 children in the finer compact level
    b.bytes = o;
    return b;
}
inline int compact_np(int n) { return (n + 1 + 31) & ~31; }
// Rows are numbered red first: a red DOF's row = (red DOFs in earlier leaves) + (red DOF bits below it in its
// leaf), a black DOF's row = nRed + the same over black bits. A colour pass is then a contiguous row range.
// red = (x+y+z) even; bit (y<<3|z) of mask word x
__device__ __forceinline__ uint64_t colour_mask(int word, int colour) {
    const uint64_t red = (word & 1) ? 0x55AA55AA55AA55AAull : 0xAA55AA55AA55AA55ull;
    return colour == 0 ? red : ~red;
}
__global__ void leaf_colour_count_kernel(const uint64_t* __restrict__ dof, int n, uint32_t* __restrict__ red, uint32_t* __restrict__ black) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    uint32_t r = 0, b = 0;
    for (int k = 0; k < 8; k++) { const uint64_t m = dof[(size_t)l * 8 + k]; r += __popcll(m & colour_mask(k, 0)); b += __popcll(m & colour_mask(k, 1)); }
    red[l] = r; black[l] = b;
}
struct RowMap { const uint64_t* dof; const uint32_t* redStart; const uint32_t* blackStart; int nRed; };
__device__ __forceinline__ int dof_row(const RowMap& R, int leaf, int off) {   // -1 if not a DOF
    const uint64_t* m = R.dof + (size_t)leaf * 8;
    const int X = off >> 6, bit = off & 63;
    const uint64_t w = m[X];
    if (!((w >> bit) & 1ull)) return -1;
    const int colour = (X + (bit >> 3) + (bit & 7)) & 1;
    int below = __popcll(w & colour_mask(X, colour) & ((1ull << bit) - 1ull));
    for (int k = 0; k < X; k++) below += __popcll(m[k] & colour_mask(k, colour));
    return colour == 0 ? (int)R.redStart[leaf] + below : R.nRed + (int)R.blackStart[leaf] + below;
}
__device__ __forceinline__ int dof_row_at(const TopoView& t, const RowMap& R, int gx, int gy, int gz) {
    const int l = topo_find(t, gx, gy, gz);
    return l < 0 ? -1 : dof_row(R, l, voxel_off(gx, gy, gz));
}
// one CTA per leaf of level L. F (finer) / C (coarser) compact neighbours are optional (RowMap.dof == null).
__global__ void __launch_bounds__(512) compact_build_kernel(LevelView L, RowMap R, int n, int np,
                                                            TopoView ft, RowMap FR, int fn, TopoView ct, RowMap CR,
                                                            uint8_t* __restrict__ blob, uint32_t* __restrict__ voxelOfRow) {
    const bool hasF = FR.dof != nullptr, hasC = CR.dof != nullptr;
    const BlobLayout B = blob_layout(np, hasF);
    float* diag = reinterpret_cast<float*>(blob + B.diag);
    float* inv = reinterpret_cast<float*>(blob + B.inv);
    float* minus = reinterpret_cast<float*>(blob + B.minus);
    uint16_t* cols = reinterpret_cast<uint16_t*>(blob + B.cols);
    uint16_t* parent = reinterpret_cast<uint16_t*>(blob + B.parent);
    uint16_t* child = reinterpret_cast<uint16_t*>(blob + B.child);
    const int leaf = blockIdx.x, off = threadIdx.x;
    if (leaf == 0 && off < np - n) {   // padding rows, the first of them is the zero dummy
        const int r = n + off;
        diag[r] = 1.f; inv[r] = 0.f;
        for (int ch = 0; ch < 3; ch++) minus[ch * np + r] = 0.f;
        for (int k = 0; k < 6; k++) cols[k * np + r] = (uint16_t)n;
        parent[r] = 0;
        if (hasF) for (int k = 0; k < 8; k++) child[k * np + r] = (uint16_t)fn;
        voxelOfRow[r] = 0;
    }
    const int row = dof_row(R, leaf, off);
    if (row < 0) return;
    const size_t i = (size_t)leaf * LEAF + off;
    const int3 o = L.t.origin[leaf];
    const int gx = o.x + (off >> 6), gy = o.y + ((off >> 3) & 7), gz = o.z + (off & 7);
    diag[row] = L.diag[i];
    inv[row] = L.invdiag[i];
    minus[row] = L.xe[i]; minus[np + row] = L.ye[i]; minus[2 * np + row] = L.ze[i];
    voxelOfRow[row] = (uint32_t)i;
    const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int r = dof_row_at(L.t, R, gx + d[k][0], gy + d[k][1], gz + d[k][2]);
        cols[k * np + row] = (uint16_t)(r < 0 ? n : r);
    }
    int pr = 0;
    if (hasC) { pr = dof_row_at(ct, CR, gx >> 1, gy >> 1, gz >> 1); if (pr < 0) pr = 0; }
    parent[row] = (uint16_t)pr;
    if (hasF) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int r = dof_row_at(ft, FR, 2 * gx + (k >> 2), 2 * gy + ((k >> 1) & 1), 2 * gz + (k & 1));
            child[k * np + row] = (uint16_t)(r < 0 ? fn : r);
        }
    }
}
struct CompactDev {        // one compact level, as the cycle kernel sees it
    int n, np, nRed, hasChild;
    int oInv, oMinus, oCols;      // byte offsets into dynamic shared memory, always resident
    int oDiag, oParent, oChild;   // byte offset into dynamic shared memory, or -1: read from the global blob
    int xOff, bOff;               // scratch
    const uint8_t* blob;          // global blob (blob_layout)
    const uint32_t* voxelOfRow;
};
struct CompactSm {
    int n, np, nRed;
    float *x, *b;
    const float *diag, *inv, *minus;
    const uint16_t *cols, *parent, *child;
};
__device__ __forceinline__ CompactSm compact_sm(const CompactDev& c, unsigned char* sm) {
    const BlobLayout B = blob_layout(c.np, c.hasChild != 0);
    CompactSm S;
    S.n = c.n; S.np = c.np; S.nRed = c.nRed;
    S.x = reinterpret_cast<float*>(sm + c.xOff);
    S.b = reinterpret_cast<float*>(sm + c.bOff);
    S.inv = reinterpret_cast<const float*>(sm + c.oInv);
    S.minus = reinterpret_cast<const float*>(sm + c.oMinus);
    S.cols = reinterpret_cast<const uint16_t*>(sm + c.oCols);
    S.diag = reinterpret_cast<const float*>(c.oDiag >= 0 ? sm + c.oDiag : c.blob + B.diag);
    S.parent = reinterpret_cast<const uint16_t*>(c.oParent >= 0 ? sm + c.oParent : c.blob + B.parent);
    S.child = reinterpret_cast<const uint16_t*>(c.oChild >= 0 ? sm + c.oChild : c.blob + B.child);
    return S;
}
// off-diagonal sum of row r, same association as offdiag()
__device__ __forceinline__ float compact_offdiag(const CompactSm& S, const float* x, int r) {
    const int np = S.np;
    const unsigned xm = S.cols[r], xp = S.cols[np + r], ym = S.cols[2 * np + r], yp = S.cols[3 * np + r],
                   zm = S.cols[4 * np + r], zp = S.cols[5 * np + r];
    const float fx = __fadd_rn(__fmul_rn(x[xp], S.minus[xp]), __fmul_rn(x[xm], S.minus[r]));
    const float fy = __fadd_rn(__fmul_rn(x[yp], S.minus[np + yp]), __fmul_rn(x[ym], S.minus[np + r]));
    const float fz = __fadd_rn(__fmul_rn(x[zp], S.minus[2 * np + zp]), __fmul_rn(x[zm], S.minus[2 * np + r]));
    return __fadd_rn(__fadd_rn(fx, fy), fz);
}
// sum over the CTA (1024 threads), same tree as cg_block_sum but one bar.sync: every warp folds the 32 warp
// partials itself; red = 2 x 32 floats used alternately
__device__ __forceinline__ float cta_sum_1024(float v, float* red, unsigned& phase) {
    for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
    float* buf = red + (phase & 1u) * 32;
    phase++;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = buf[threadIdx.x & 31];
    for (int d = 16; d > 0; d >>= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, d));
    return r;
}
// Jacobi-preconditioned CG on the compact coarsest level (same arithmetic as coarse_cg_smem); x = S.x, the
// residual lives in S.b, P and T in `pt` (2 x np floats)
__device__ __forceinline__ float2 cta_sum2_1024(float a, float b, float2* red2, unsigned& phase) {
    for (int d = 16; d > 0; d >>= 1) { a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, d)); b = __fadd_rn(b, __shfl_xor_sync(0xffffffffu, b, d)); }
    float2* buf = red2 + (phase & 1u) * 32;
    phase++;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = make_float2(a, b);
    __syncthreads();
    float2 r = buf[threadIdx.x & 31];
    for (int d = 16; d > 0; d >>= 1) { r.x = __fadd_rn(r.x, __shfl_xor_sync(0xffffffffu, r.x, d)); r.y = __fadd_rn(r.y, __shfl_xor_sync(0xffffffffu, r.y, d)); }
    return r;
}
__device__ void compact_cg(const CompactSm& S, float* pt, float* red, float2* red2, unsigned& phase) {
    const int n = S.n, tid = threadIdx.x;
    float* X = S.x; float* R = S.b; float* P = pt; float* T = pt + S.np;
    auto dinv = [&](int r) { float d = S.diag[r]; return d != 0.f ? __fdiv_rn(1.0f, d) : 1.0f; };
    for (int r = tid; r < S.np; r += BOT_THREADS) { X[r] = 0.f; P[r] = 0.f; if (r >= n) R[r] = 0.f; }
    float acc = 0.f;
    for (int r = tid; r < n; r += BOT_THREADS) acc = __fadd_rn(acc, __fmul_rn(R[r], R[r]));
    const float rhsNorm2 = cta_sum_1024(acc, red, phase);
    if (rhsNorm2 != 0.f) {
        const float tol = 1.1920929e-07f;
        const float threshold = fmaxf(__fmul_rn(__fmul_rn(tol, tol), rhsNorm2), 1.17549435e-38f);
        float residualNorm2 = rhsNorm2;
        if (!(residualNorm2 < threshold)) {
            acc = 0.f;
            for (int r = tid; r < n; r += BOT_THREADS) { float pv = __fmul_rn(dinv(r), R[r]); P[r] = pv; acc = __fadd_rn(acc, __fmul_rn(R[r], pv)); }
            float absNew = cta_sum_1024(acc, red, phase);   // the bar.sync inside also publishes P
            const int np = S.np;
            for (int it = 0; it < 10; it++) {
                acc = 0.f;
                for (int r = tid; r < n; r += BOT_THREADS) {
                    float s = __fadd_rn(0.f, __fmul_rn(S.diag[r], P[r]));
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const unsigned cm = S.cols[(2 * ch) * np + r], cp = S.cols[(2 * ch + 1) * np + r];
                        if (cm != (unsigned)n) s = __fadd_rn(s, __fmul_rn(S.minus[ch * np + r], P[cm]));
                        if (cp != (unsigned)n) s = __fadd_rn(s, __fmul_rn(S.minus[ch * np + cp], P[cp]));
                    }
                    T[r] = s;
                    acc = __fadd_rn(acc, __fmul_rn(P[r], s));
                }
                const float pt2 = cta_sum_1024(acc, red, phase);
                const float alpha = __fdiv_rn(absNew, pt2);
                acc = 0.f;
                float acc2 = 0.f;
                for (int r = tid; r < n; r += BOT_THREADS) {
                    X[r] = __fadd_rn(X[r], __fmul_rn(alpha, P[r]));
                    const float rv = __fsub_rn(R[r], __fmul_rn(alpha, T[r]));
                    R[r] = rv;
                    acc = __fadd_rn(acc, __fmul_rn(rv, rv));
                    // |r|^2 and r.z (z = D^-1 r) go through one reduction: both only need the new r
                    const float zv = __fmul_rn(dinv(r), rv);
                    T[r] = zv;
                    acc2 = __fadd_rn(acc2, __fmul_rn(rv, zv));
                }
                const float2 both = cta_sum2_1024(acc, acc2, red2, phase);   // all SpMV reads of P are behind this bar.sync
                residualNorm2 = both.x;
                if (residualNorm2 < threshold) break;
                const float absOld = absNew;
                absNew = both.y;
                const float beta = __fdiv_rn(absNew, absOld);
                for (int r = tid; r < n; r += BOT_THREADS) P[r] = __fadd_rn(T[r], __fmul_rn(beta, P[r]));
                __syncthreads();
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------- the whole preconditioner in one launch
// mg_cycle_kernel: a persistent cooperative kernel, one CTA of 1024 threads per SM, that executes the COMPLETE
// mu-cycle (every level) as a flat op list. Ops on the large levels are spread over all CTAs (leaf chunks
// strided by the grid) and separated by a device-wide barrier; runs of ops on the bottom levels are executed by
// CTA 0 alone out of its shared memory (as in mg_bottom_kernel) while the others wait at the next barrier.
// One launch replaces ~370 kernel launches per preconditioner application at 512^3 -- the host could not
// issue those faster than ~4.5 us each, which was the solver's real bound.
constexpr int CYC_MAX_LEVELS = 10;
constexpr int CYC_GRID_SCRATCH = (4 * TILE + 4 * LEAF) * 4;   // bytes: four 10^3 tiles + four leaf residuals
struct CycleParams {
    BottomLevel lv[CYC_MAX_LEVELS];   // leaf form (global memory), used on levels < compactFirst
    CompactDev cl[CYC_MAX_LEVELS];    // compact rows (CTA 0's shared memory), levels >= compactFirst
    const uint8_t* prog;              // op | level << 3
    int nOps, nLevels, compactFirst;
    int scratchOff, cgOff;            // byte offsets into dynamic shared memory
    float w, oneMinusW, prolongAlpha;
    unsigned* barrier;                // arrival counter, zero at launch
    unsigned long long* trace;        // optional: %globaltimer at the start of every op (CTA 0), nOps + 1 entries
    int gridOnly;                     // the program holds ops of levels < compactFirst only (hybrid path): no compact staging
    int cacheInfo;                    // keep the leaf records of the first grid levels in shared memory (CY_CACHE_*)
    int l1Loads;                      // colour passes read x and b through L1 (M_GRID_L1)
};
constexpr int CY_CACHE_LEVELS = 3, CY_CACHE_SLOTS = 48;
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Device-wide barrier for the co-resident grid: one arrival counter, release by fence + atomic, acquire by polling.
// Measured on B200 with 148 CTAs x 1024 threads (tools/micro/barrier_bench.cu), us per barrier, no work / with work:
//   this (atomic counter, acquire poll)          1.41 / 2.03      cooperative_groups grid.sync   1.26 / 2.02
//   one flag per CTA, every CTA polls all flags   3.54 / 3.83      master gather + private release 3.28 / 3.64
// (the flag variants win only below ~32 CTAs: 0.97 us). The barrier, not bandwidth, bounds every op on the small
// levels of the cycle.
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& passed) {
    __syncthreads();
    passed++;
    if (threadIdx.x == 0) {
        const unsigned target = passed * gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire(ctr) < target) { }
    }
    __syncthreads();
}
// ops of the compact bottom, executed by CTA 0 between two device-wide barriers
__device__ void compact_run(const CycleParams& P, unsigned char* sm, int k, int kEnd, float* red, float2* red2, unsigned& phase) {
    const int tid = threadIdx.x;
    const CompactDev& TD = P.cl[P.compactFirst];
    const BottomLevel& TG = P.lv[P.compactFirst];
    // every x starts as zero (row n of each level must read 0 for the whole run; the scratch region is shared
    // with the leaf-tile staging of the grid ops and holds garbage here)
    for (int l = P.compactFirst; l < P.nLevels; l++) {
        const CompactSm S = compact_sm(P.cl[l], sm);
        for (int r = tid; r < S.np; r += BOT_THREADS) { S.x[r] = 0.f; S.b[r] = 0.f; }
    }
    __syncthreads();
    {
        const CompactSm S = compact_sm(TD, sm);
        const int first = P.prog[k] & 7;
        const bool needX = first != OP_ZERO_RED && first != OP_COARSE;
        for (int r = tid; r < S.n; r += BOT_THREADS) {
            const uint32_t v = __ldg(&TD.voxelOfRow[r]);
            S.b[r] = ld_l2(&TG.b[v]);
            if (needX) S.x[r] = ld_l2(&TG.x[v]);
        }
    }
    __syncthreads();
    for (int q = k; q < kEnd; q++) {
        const int cd = P.prog[q] & 7, l = P.prog[q] >> 3;
        if (P.trace && tid == 0) P.trace[q] = globaltimer();
        const CompactSm S = compact_sm(P.cl[l], sm);
        if (cd == OP_ZERO_RED) {
            for (int r = tid; r < S.n; r += BOT_THREADS)
                S.x[r] = r < S.nRed ? __fmul_rn(__fmul_rn(S.b[r], S.inv[r]), P.w) : 0.f;
        } else if (cd == OP_RED || cd == OP_BLACK) {
            const int r0 = cd == OP_RED ? 0 : S.nRed, r1 = cd == OP_RED ? S.nRed : S.n;
            for (int r = r0 + tid; r < r1; r += BOT_THREADS) {
                const float od = compact_offdiag(S, S.x, r);
                const float tt = __fmul_rn(__fmul_rn(__fsub_rn(S.b[r], od), S.inv[r]), P.w);
                S.x[r] = __fmaf_rn(S.x[r], P.oneMinusW, tt);
            }
        } else if (cd == OP_RESID_RESTRICT) {
            // item = coarse row * 8 + child slot; the eight residuals of a coarse row sit in eight adjacent lanes
            // and are added in the reference's (ii,jj,kk) order (uaamg.cpp:1773-1833)
            const CompactSm C = compact_sm(P.cl[l + 1], sm);
            const int items = (C.n * 8 + 31) & ~31;
            // child row and its diagonal are fetched one iteration ahead (either may live in global memory)
            auto child_of = [&](int w) { const int crow = w >> 3; return (w < items && crow < C.n) ? (int)C.child[(w & 7) * C.np + crow] : S.n; };
            int w = tid;
            int f = child_of(w);
            float dg = f < S.n ? S.diag[f] : 0.f;
            while (w < items) {
                const int wn = w + BOT_THREADS;
                const int fn = child_of(wn);
                const float dgn = fn < S.n ? S.diag[fn] : 0.f;
                const int crow = w >> 3, slot = w & 7;
                const bool present = f < S.n;
                float res = 0.f;
                if (present) res = __fsub_rn(S.b[f], __fmaf_rn(S.x[f], dg, compact_offdiag(S, S.x, f)));
                const unsigned pm = __ballot_sync(0xffffffffu, present);
                const int base = (tid & 31) & ~7;
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float v = __shfl_sync(0xffffffffu, res, base + j);
                    if ((pm >> (base + j)) & 1u) sum = __fadd_rn(sum, v);
                }
                if (slot == 0 && crow < C.n) C.b[crow] = __fmul_rn(sum, 0.125f);
                w = wn; f = fn; dg = dgn;
            }
        } else if (cd == OP_PROLONG) {
            const CompactSm C = compact_sm(P.cl[l + 1], sm);
            for (int r = tid; r < S.n; r += BOT_THREADS)
                S.x[r] = __fadd_rn(S.x[r], __fmul_rn(P.prolongAlpha, C.x[S.parent[r]]));
        } else {
            compact_cg(S, reinterpret_cast<float*>(sm + P.cgOff), red, red2, phase);
        }
        __syncthreads();
    }
    {
        const CompactSm S = compact_sm(TD, sm);
        for (int r = tid; r < S.n; r += BOT_THREADS) TG.x[__ldg(&TD.voxelOfRow[r])] = S.x[r];
    }
}
__global__ void __launch_bounds__(BOT_THREADS) mg_cycle_kernel(const __grid_constant__ CycleParams P) {
    extern __shared__ __align__(16) unsigned char cycsm[];
    unsigned char* dynsm = cycsm;
    __shared__ float red[64];
    __shared__ float2 red2[64];
    float* tiles = reinterpret_cast<float*>(dynsm + P.scratchOff);
    float* sresGrid = tiles + 4 * TILE;
    const int tid = threadIdx.x, G = gridDim.x, c = blockIdx.x;
    unsigned passed = 0, phase = 0;
    int k = 0;
    if (c == 0 && !P.gridOnly) {   // stage the compact matrices once
        for (int l = P.compactFirst; l < P.nLevels; l++) {
            const CompactDev& D = P.cl[l];
            const BlobLayout B = blob_layout(D.np, D.hasChild != 0);
            const int src[6] = {B.inv, B.minus, B.cols, B.diag, B.parent, B.child};
            const int dst[6] = {D.oInv, D.oMinus, D.oCols, D.oDiag, D.oParent, D.hasChild ? D.oChild : -1};
            const int len[6] = {4 * D.np, 12 * D.np, 12 * D.np, 4 * D.np, 2 * D.np, 16 * D.np};
            for (int q = 0; q < 6; q++) {
                if (dst[q] < 0) continue;
                const uint4* sp = reinterpret_cast<const uint4*>(D.blob + src[q]);
                uint4* dp = reinterpret_cast<uint4*>(dynsm + dst[q]);
                for (int i = tid; i < (len[q] >> 4); i += BOT_THREADS) dp[i] = __ldg(&sp[i]);
            }
        }
        __syncthreads();
    }
    // the leaf records of the chunks this CTA visits on the first grid levels (chunk it of a pass = leaves c*4 + it*G*4 + 0..3)
    __shared__ CachedInfo sInfo[CY_CACHE_LEVELS][CY_CACHE_SLOTS];
    if (P.cacheInfo) {
        const int ncl = P.compactFirst < CY_CACHE_LEVELS ? P.compactFirst : CY_CACHE_LEVELS;
        for (int l = 0; l < ncl; l++) {
            const BottomLevel& B = P.lv[l];
            for (int i = tid; i < CY_CACHE_SLOTS * 6; i += BOT_THREADS) {
                const int slot = i / 6, q = i - slot * 6;
                const int leaf = c * 4 + (slot >> 2) * G * 4 + (slot & 3);
                if (leaf < B.n) reinterpret_cast<int4*>(&sInfo[l][slot])[q] = __ldg(reinterpret_cast<const int4*>(B.v.info + leaf) + q);
            }
        }
        __syncthreads();
    }
    while (k < P.nOps) {
        const int code = P.prog[k] & 7, li = P.prog[k] >> 3;
        if (li < P.compactFirst) {
            if (P.trace && c == 0 && tid == 0) P.trace[k] = globaltimer();
            const BottomLevel& B = P.lv[li];
            const bool cached = P.cacheInfo && li < CY_CACHE_LEVELS;
            if (code == OP_ZERO_RED && cached) {
                const float invDefault = __fdiv_rn(1.0f, __fmul_rn(6.0f, B.v.term));
                int slot = tid >> 8, base = c * 4;
                for (; base < B.n && slot < CY_CACHE_SLOTS; base += G * 4, slot += 4) {
                    const int leaf = base + (tid >> 8);
                    if (leaf < B.n) zero_red_pair_cached<M_GRID_L1>(B.v, B.x, B.b, leaf, tid & 255, P.w, sInfo[li][slot], invDefault);
                }
                for (int leaf = base + (tid >> 8); leaf < B.n; leaf += G * 4) { zero_red_leaf<M_GRID>(B.v, B.x, B.b, leaf, (tid & 255) * 2, P.w); zero_red_leaf<M_GRID>(B.v, B.x, B.b, leaf, (tid & 255) * 2 + 1, P.w); }
            } else if (code == OP_ZERO_RED) {
                for (int base = c * 2; base < B.n; base += G * 2) { int leaf = base + (tid >> 9); if (leaf < B.n) zero_red_leaf<M_GRID>(B.v, B.x, B.b, leaf, tid & 511, P.w); }
            } else if (code == OP_RED || code == OP_BLACK) {
                // no block barrier inside a pass: warps run ahead into the next chunk, which is what keeps loads in
                // flight (a shared-memory tile version with two bar.sync per chunk measured 25 us vs 15 us at level 0)
                int slot = tid >> 8;
                const float invDefault = __fdiv_rn(1.0f, __fmul_rn(6.0f, B.v.term));
                for (int base = c * 4; base < B.n; base += G * 4, slot += 4) {
                    const int leaf = base + (tid >> 8);
                    if (leaf >= B.n) continue;
                    if (cached && slot < CY_CACHE_SLOTS) {
                        if (P.l1Loads) rbgs_leaf_cached<M_GRID_L1>(B.v, B.x, B.b, leaf, tid & 255, code == OP_RED ? 0 : 1, P.w, P.oneMinusW, sInfo[li][slot], invDefault);
                        else rbgs_leaf_cached<M_GRID>(B.v, B.x, B.b, leaf, tid & 255, code == OP_RED ? 0 : 1, P.w, P.oneMinusW, sInfo[li][slot], invDefault);
                    } else rbgs_leaf<M_GRID>(B.v, B.x, B.b, leaf, tid & 255, code == OP_RED ? 0 : 1, P.w, P.oneMinusW);
                }
            } else if (code == OP_RESID_RESTRICT) {
                // 256 threads per leaf, two z-adjacent voxels each; residual into shared memory, then restrict
                const BottomLevel& C = P.lv[li + 1];
                const int g = tid >> 8, t = tid & 255;
                float* T = tiles + g * TILE;
                float* R = sresGrid + g * LEAF;
                int slot = g;
                int base = c * 4;
                if (cached && B.hasParent) {   // chunks whose records are in shared memory: no tiles, no block barriers
                    for (; base < B.n && slot < CY_CACHE_SLOTS; base += G * 4, slot += 4) {
                        const int leaf = base + g;
                        if (leaf < B.n) resid_restrict_cached<M_GRID_L1>(B.v, B.x, B.b, B.bReadOnly != 0, C.b, leaf, t, sInfo[li][slot]);
                    }
                }
                for (; base < B.n; base += G * 4, slot += 4) {
                    const int leaf = base + g;
                    LeafInfo info;
                    bool live = false;
                    float2 bv = make_float2(0.f, 0.f);
                    if (leaf < B.n) {
                        info = (cached && slot < CY_CACHE_SLOTS) ? cached_info(sInfo[li][slot]) : load_info(B.v, leaf);
                        live = (info.flags & LI_ANY) != 0;
                        if (live) {
                            tile_load(B.x, leaf, info, T, t);
                            bv = B.bReadOnly ? __ldg(reinterpret_cast<const float2*>(B.b + (size_t)leaf * LEAF + 2 * t)) : ld_l2_f2(B.b + (size_t)leaf * LEAF + 2 * t);
                        }
                    }
                    __syncthreads();
                    if (live) {
                        const int X = t >> 5, Y = (t >> 2) & 7;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int Z = ((t & 3) << 1) | h;
                            const int off = (X << 6) | (Y << 3) | Z;
                            float r = 0.f;
                            if (dof_bit(B.v, leaf, off)) {
                                const float od = (info.flags & LI_CONST) ? offdiag_tile<true>(B.v, T, leaf, X, Y, Z, info) : offdiag_tile<false>(B.v, T, leaf, X, Y, Z, info);
                                const float dg = (info.flags & LI_DIAG) ? __fmul_rn(6.0f, B.v.term) : __ldg(&B.v.diag[(size_t)leaf * LEAF + off]);
                                r = __fsub_rn(h == 0 ? bv.x : bv.y, __fmaf_rn(T[tidx(X, Y, Z)], dg, od));
                            }
                            R[off] = r;
                        }
                    }
                    __syncthreads();
                    if (live && t < 64) restrict_leaf(B.v, C.v, R, leaf, t, C.b);
                    __syncthreads();
                }
            } else if (code == OP_PROLONG) {
                const BottomLevel& C = P.lv[li + 1];
                if (cached && B.hasParent) {
                    int slot = tid >> 8, base = c * 4;
                    for (; base < B.n && slot < CY_CACHE_SLOTS; base += G * 4, slot += 4) {
                        const int leaf = base + (tid >> 8);
                        if (leaf < B.n) prolong_pair_cached<M_GRID_L1>(C.v, B.x, C.x, leaf, tid & 255, P.prolongAlpha, sInfo[li][slot]);
                    }
                    // what the cache does not hold (levels with more than CY_CACHE_SLOTS leaves per CTA): per voxel, through the directory
                    for (int leaf = base + (tid >> 8); leaf < B.n; leaf += G * 4) { prolong_voxel<M_GRID>(B.v, C.v, B.x, C.x, leaf, (tid & 255) * 2, P.prolongAlpha); prolong_voxel<M_GRID>(B.v, C.v, B.x, C.x, leaf, (tid & 255) * 2 + 1, P.prolongAlpha); }
                } else
                for (int base = c * 2; base < B.n; base += G * 2) { int leaf = base + (tid >> 9); if (leaf < B.n) prolong_voxel<M_GRID>(B.v, C.v, B.x, C.x, leaf, tid & 511, P.prolongAlpha); }
            }
            k++;
            grid_barrier(P.barrier, passed);
            continue;
        }
        // a run of compact ops [k, kEnd): CTA 0 alone, shared-memory resident
        int kEnd = k;
        while (kEnd < P.nOps && (P.prog[kEnd] >> 3) >= P.compactFirst) kEnd++;
        if (c == 0) compact_run(P, dynsm, k, kEnd, red, red2, phase);
        k = kEnd;
        if (k < P.nOps) grid_barrier(P.barrier, passed);
    }
    if (P.trace && c == 0 && tid == 0) P.trace[P.nOps] = globaltimer();
}
__global__ void leaf_popcount_kernel(const uint64_t* __restrict__ dof, int n, uint32_t* __restrict__ out) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    uint32_t c = 0;
    for (int k = 0; k < 8; k++) c += __popcll(dof[(size_t)l * 8 + k]);
    out[l] = c;
}
__global__ void warm_start_kernel(TopoView t, const uint64_t* __restrict__ dof, TopoView ot, const float* __restrict__ oldP,
                                  float oldBg, float* __restrict__ x) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!mask_get(dof, leaf, off)) return;
    int3 o = t.origin[leaf];
    float v = ot.n > 0 ? grid_get(ot, oldP, oldBg, o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)) : oldBg;
    if (isfinite(v)) x[(size_t)leaf * LEAF + off] = v;
}

// ---------------------------------------------------------------- slab decomposition: global level 1
// Every rank coarsens its own pool; a level-1 cell whose eight children it owns is exact there (the coefficients read
// the children and the DOF bits one fine voxel below them, which the ghost layer holds). This kernel writes the
// owned cells of the GLOBAL level-1 leaves (defaults where the rank has no such leaf, zero for cells of other ranks);
// an all-reduce (sum) then gives every rank the complete level.
__global__ void __launch_bounds__(512) dd_scatter_coarse_kernel(TopoView gt, TopoView lt, const uint64_t* __restrict__ ldof,
                                                                const float* __restrict__ ldiag, const float* __restrict__ lx,
                                                                const float* __restrict__ ly, const float* __restrict__ lz,
                                                                int xlo, int xhi, float cterm, uint64_t* __restrict__ gdof,
                                                                float* __restrict__ gdiag, float* __restrict__ gx,
                                                                float* __restrict__ gy, float* __restrict__ gz) {
    const int leaf = blockIdx.x, off = threadIdx.x;
    const int3 o = gt.origin[leaf];
    const int X = o.x + (off >> 6), Y = o.y + ((off >> 3) & 7), Z = o.z + (off & 7);
    float d = 0.f, a = 0.f, b = 0.f, c = 0.f;
    bool on = false;
    if (X >= xlo && X < xhi) {
        const int l = topo_find(lt, X, Y, Z);
        if (l >= 0) {
            const size_t i = (size_t)l * LEAF + off;
            d = ldiag[i]; a = lx[i]; b = ly[i]; c = lz[i];
            on = mask_get(ldof, l, off);
        } else {
            d = __fmul_rn(6.0f, cterm); a = -cterm; b = -cterm; c = -cterm;
        }
    }
    const size_t i = (size_t)leaf * LEAF + off;
    gdiag[i] = d; gx[i] = a; gy[i] = b; gz[i] = c;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<uint32_t*>(gdof)[(size_t)leaf * 16 + (threadIdx.x >> 5)] = bal;
}

#include "mg_cluster.cuh"

// ---------------------------------------------------------------- host-side solver
// The bottom runs on ONE SM: its coefficient set (7 arrays x 2 KB per leaf) has to stay inside that SM's
// L1 + shared memory (256 KB) or every pass turns into a serial chain of L2 round trips (measured: a 64-leaf
// level inside the bottom cost 183 us per visit). 12 leaves = 168 KB.
constexpr int BOTTOM_MAX_TOTAL_LEAVES = 12;
constexpr size_t BOTTOM_SMEM_CAP = 190 * 1024;   // dynamic shared memory of mg_bottom_kernel (x, lower b, coarsest CG)

struct Solver {
    World* w;
    std::vector<std::unique_ptr<Level>> levels;
    float dt = 0.f;
    // coarsest ELL
    int ndof = 0, ndofPad = 0;
    DBuf<int> ellCols, rowOfVoxel;
    DBuf<float> ellVals;
    DBuf<float> partial, scal;  // [max leaves], [8]
    DBuf<unsigned> counter;
    // mg_cycle_kernel (the whole preconditioner in one cooperative launch)
    struct CompactHost {
        int n = 0, np = 0;
        bool hasChild = false;
        int nRed = 0;
        DBuf<uint32_t> redStart, blackStart, voxelOfRow;
        DBuf<uint8_t> blob;
        int oInv = 0, oMinus = 0, oCols = 0, oDiag = -1, oParent = -1, oChild = -1, xOff = 0, bOff = 0;
    };
    std::vector<CompactHost> compact;   // index = level - compactFirst
    int compactFirst = 0, scratchOff = 0, cgOff = 0;
    bool cycleReady = false;
    DBuf<uint8_t> cycleProg, cycleProg2;   // visit from a zero guess / from the current iterate
    std::vector<uint8_t> hostOps, hostOps2;
    int cycleOps = 0, cycleOps2 = 0, cycleGrid = 1;
    // slab decomposition (dd.cu): level 0 lives on this rank's pool, levels >= 1 are assembled globally and
    // replicated on every rank inside `coarse`
    bool dd = false;
    bool coarseOnly = false;   // this object is the replicated coarse part of a decomposed solve
    std::unique_ptr<Solver> coarse;
    static constexpr int cycleGridMax = 192;
    size_t cycleSmem = 0;
    DBuf<unsigned> cycleBarrier;
    // mg_cluster_kernel (mg_cluster.cuh): levels >= clFirst run inside one thread-block cluster
    struct ClusterHost { DBuf<uint32_t> assign; DBuf<ClMeta> meta; int nN = 0, nC = 0, xbPer = 0, coefPer = 0, xbOff = 0, coefOff = 0, metaOff = 0; };
    std::vector<ClusterHost> clHost;   // index = level
    DBuf<int> clCounts;
    bool tiles = true;                 // FLIPB200_MG_PATH=cycle selects the round-1 one-launch cycle kernel instead
    bool clusterReady = false;
    int clFirst = 0, clSize = 0, clCandFirst = 0;
    size_t clSmem = 0;
    int clXbBytes = 0, clCgOff = 0, clSresOff = 0, clProgOff = 0, clDumpOff = 0;
    // hybrid: levels < clFirst as segments of the one-launch cycle kernel, the cluster kernel between them
    bool hybridReady = false;
    DBuf<uint8_t> hybridProg;
    std::vector<uint8_t> hybridHost;
    std::vector<std::pair<int, int>> hybridSeg, hybridSeg2;   // (offset, ops): visit from a zero guess / from the current iterate
    DBuf<uint8_t> clProg[3];           // one visit from a zero guess / one visit from the current iterate / both in a row
    std::vector<uint8_t> clHostOps[3];
    int clOps[3] = {0, 0, 0};
    int bottomFirst = 0;        // first level handled by mg_bottom_kernel
    bool bottomInSmem = false;  // x (and the lower levels' b) of the bottom live in shared memory
    size_t bottomSmem = 0;
    int bottom_leaves(int first) const {
        int t = 0;
        for (int l = first; l < (int)levels.size(); l++) t += levels[l]->n;
        return t;
    }
    size_t bottom_need(int first) const {
        size_t need = (size_t)5 * ndofPad * sizeof(float);
        for (int l = first; l < (int)levels.size(); l++) need += (size_t)levels[l]->n * LEAF * sizeof(float) * 2;
        return need;
    }

    void alloc_vectors(Level& L) {
        size_t n = (size_t)L.n * LEAF;
        L.x.alloc(n, w->stream); L.b.alloc(n, w->stream);
        L.x.zero(); L.b.zero();
    }
    void finish_level(Level& L) {
        L.invdiag.alloc((size_t)L.n * LEAF, w->stream);
        L.flags.alloc(L.n, w->stream);
        L.info.alloc(L.n, w->stream);
        FB_LAUNCH(w, "mg_trim", (size_t)L.n * LEAF * 24) trim_kernel<<<L.n, 512, 0, w->stream>>>(L.dof.p, L.diag.p, L.xe.p, L.ye.p, L.ze.p, L.invdiag.p, L.flags.p, L.term);
        check_launch("trim");
        FB_LAUNCH(w, "mg_leaf_info", (size_t)L.n * 140) leaf_info_kernel<<<(L.n + 127) / 128, 128, 0, w->stream>>>(L.topo->view(), L.dof.p, L.flags.p, L.info.p);
        check_launch("leaf_info");
        // DOF count and the coarser level's candidate leaves in one read-back
        DBuf<unsigned long long> c(2, w->stream);
        c.zero();
        mask_count_async(w, L.dof.p, L.n, c.p);
        L.cand.alloc(L.n + 1, w->stream);
        FB_LAUNCH(w, "mg_coarse_leaves", (size_t)L.n * 76) dof_leaf_origins_kernel<<<(L.n + 127) / 128, 128, 0, w->stream>>>(L.topo->view(), L.dof.p, L.cand.p, reinterpret_cast<uint32_t*>(c.p + 1));
        check_launch("dof_leaf_origins");
        unsigned long long h[2] = {0, 0};
        read_back(w, h, c.p, 16);
        L.numDof = (int)h[0];
        L.candCount = (int)(h[1] & 0xffffffffull);
    }
    std::unique_ptr<Level> coarsen_raw() {
        const Level& F = *levels.back();
        auto Lp = std::make_unique<Level>();
        Level& L = *Lp;
        // coarse leaf coordinate = floor(fine leaf coordinate / 2): the fine directory bounds give the coarse ones
        const Topo& ft = *F.topo;
        const int bb[6] = {ft.dmin.x >> 1, ft.dmin.y >> 1, ft.dmin.z >> 1,
                           (ft.dmin.x + ft.ddim.x - 1) >> 1, (ft.dmin.y + ft.ddim.y - 1) >> 1, (ft.dmin.z + ft.ddim.z - 1) >> 1};
        L.topo = topo_from_origins_dev(w, F.cand.p, F.candCount, false, bb);
        L.n = L.topo->n;
        L.dx = 2.0f * F.dx;
        L.term = dt / (L.dx * L.dx);
        size_t n = (size_t)L.n * LEAF;
        L.dof.alloc((size_t)L.n * 8, w->stream);
        L.diag.alloc(n, w->stream); L.xe.alloc(n, w->stream); L.ye.alloc(n, w->stream); L.ze.alloc(n, w->stream);
        FB_LAUNCH(w, "mg_coarsen", (size_t)L.n * LEAF * 16 + (size_t)F.n * LEAF * 16) coarsen_kernel<<<L.n, 512, 0, w->stream>>>(view_of(F), L.topo->view(), L.term, L.dof.p, L.diag.p, L.xe.p, L.ye.p, L.ze.p);
        check_launch("coarsen");
        return Lp;
    }
    void coarsen() {
        std::unique_ptr<Level> Lp = coarsen_raw();
        finish_level(*Lp);
        alloc_vectors(*Lp);
        Level& F = *levels.back();
        FB_LAUNCH(w, "mg_leaf_parent", (size_t)F.n * 24) leaf_parent_kernel<<<(F.n + 127) / 128, 128, 0, w->stream>>>(F.topo->view(), Lp->topo->view(), F.info.p);
        check_launch("leaf_parent");
        F.hasParent = true;
        levels.push_back(std::move(Lp));
    }
    // slab decomposition: assemble level 1 globally, then build and keep the rest of the hierarchy replicated
    void dd_build_coarse() {
        FB_PHASE(w, "ppe dd_build_coarse");
        FB_REQUIRE(levels[0]->numDof > MAX_COARSEST, FLIPB200_ERR_DOMAIN,
                   "slab decomposition needs at least two multigrid levels (more than 4000 pressure DOFs)");
        std::unique_ptr<Level> loc = coarsen_raw();
        const int R = w->nRanks, me = w->rank;
        auto ph = std::make_unique<PhaseTimer>(w, "ppe coarse: gather leaf lists");
        DBuf<int> cnt(R, w->stream);
        cnt.zero();
        const int mine = loc->n;
        FB_CUDA(cudaMemcpyAsync(cnt.p + me, &mine, 4, cudaMemcpyHostToDevice, w->stream));
        comm_allreduce(w, cnt.p, R, CT_I32, false);
        std::vector<int> counts(R);
        FB_CUDA(cudaMemcpyAsync(counts.data(), cnt.p, 4 * R, cudaMemcpyDeviceToHost, w->stream));
        sync(w);
        int maxCnt = 1, total = 0;
        for (int c : counts) { maxCnt = std::max(maxCnt, c); total += c; }
        DBuf<int> gath((size_t)R * maxCnt * 3, w->stream);
        gath.zero();
        if (mine) FB_CUDA(cudaMemcpyAsync(gath.p + (size_t)me * maxCnt * 3, loc->topo->origin.p, (size_t)mine * 12, cudaMemcpyDeviceToDevice, w->stream));
        comm_allreduce(w, gath.p, (size_t)R * maxCnt * 3, CT_I32, false);
        DBuf<int3> cand(total + 1, w->stream);
        for (int r = 0, at = 0; r < R; at += counts[r], r++)
            if (counts[r]) FB_CUDA(cudaMemcpyAsync(cand.p + at, gath.p + (size_t)r * maxCnt * 3, (size_t)counts[r] * 12, cudaMemcpyDeviceToDevice, w->stream));
        ph.reset(); ph = std::make_unique<PhaseTimer>(w, "ppe coarse: global level-1 topology + scatter");
        auto Gp = std::make_unique<Level>();
        Level& G = *Gp;
        G.topo = topo_from_origins_dev(w, cand.p, total, false);
        G.n = G.topo->n;
        G.dx = loc->dx; G.term = loc->term;
        const size_t nv = (size_t)G.n * LEAF;
        G.dof.alloc((size_t)G.n * 8, w->stream);
        {   // the four arrays of the previous solve are reused when they are large enough (dd_scatter_coarse_kernel writes every element)
            DBuf<float>* dst[4] = {&G.diag, &G.xe, &G.ye, &G.ze};
            for (int k = 0; k < 4; k++) {
                if (w->ddCoarse[k].n >= nv) *dst[k] = std::move(w->ddCoarse[k]);
                else dst[k]->alloc(nv + nv / 4, w->stream);
            }
        }
        int lo, hi;
        dd_owned_coords(w, &lo, &hi);
        const int xlo = me > 0 ? 4 * lo : INT_MIN, xhi = me < R - 1 ? 4 * hi : INT_MAX;   // level-1 cells: fine voxel / 2
        TopoView lt = loc->topo->view();
        FB_LAUNCH(w, "dd_scatter_coarse", nv * 40) dd_scatter_coarse_kernel<<<G.n, 512, 0, w->stream>>>(G.topo->view(), lt, loc->dof.p, loc->diag.p, loc->xe.p, loc->ye.p, loc->ze.p,
                                                                                                       xlo, xhi, G.term, G.dof.p, G.diag.p, G.xe.p, G.ye.p, G.ze.p);
        check_launch("dd_scatter_coarse");
        ph.reset(); ph = std::make_unique<PhaseTimer>(w, "ppe coarse: all-reduce coefficients");
        // every cell has exactly one non-zero contributor: sum == assembly. The four coefficient arrays travel as ONE message
        // through a persistent staging buffer (five all-reduces of freshly allocated arrays took 5.8 ms per solve at N = 8,
        // profiles/r03_phase_trace_n8.txt)
        static const bool stageAr = !(getenv("FLIPB200_DD_AR_STAGE") && atoi(getenv("FLIPB200_DD_AR_STAGE")) == 0);
        if (stageAr) {
            if (w->ddStage.n < 4 * nv) w->ddStage.alloc(4 * nv + nv / 2, w->stream);
            float* st = w->ddStage.p;
            float* src[4] = {G.diag.p, G.xe.p, G.ye.p, G.ze.p};
            for (int k = 0; k < 4; k++) FB_CUDA(cudaMemcpyAsync(st + (size_t)k * nv, src[k], nv * 4, cudaMemcpyDeviceToDevice, w->stream));
            comm_allreduce(w, st, 4 * nv, CT_F32, false);
            for (int k = 0; k < 4; k++) FB_CUDA(cudaMemcpyAsync(src[k], st + (size_t)k * nv, nv * 4, cudaMemcpyDeviceToDevice, w->stream));
        } else {
            comm_allreduce(w, G.diag.p, nv, CT_F32, false);
            comm_allreduce(w, G.xe.p, nv, CT_F32, false);
            comm_allreduce(w, G.ye.p, nv, CT_F32, false);
            comm_allreduce(w, G.ze.p, nv, CT_F32, false);
        }
        comm_allreduce(w, G.dof.p, (size_t)G.n * 16, CT_U32, false);   // disjoint bits: sum == or
        ph.reset(); ph = std::make_unique<PhaseTimer>(w, "ppe coarse: replicated hierarchy set-up");
        coarse = std::make_unique<Solver>();
        coarse->w = w; coarse->dt = dt; coarse->coarseOnly = true; coarse->tiles = tiles;
        coarse->finish_level(G);
        coarse->alloc_vectors(G);
        coarse->levels.push_back(std::move(Gp));
        while (coarse->levels.back()->numDof > MAX_COARSEST) coarse->coarsen();
        coarse->build_coarsest();
        if (!getenv("FLIPB200_NO_CYCLE_KERNEL")) coarse->prepare_cycle(4);
        ph.reset();
    }
    void build_coarsest() {
        Level& L = *levels.back();
        ndof = L.numDof;
        ndofPad = (ndof + 31) & ~31;
        if (ndofPad == 0) ndofPad = 32;
        DBuf<uint32_t> leafCnt(L.n + 1, w->stream);
        leafCnt.zero();
        FB_LAUNCH(w, "mg_leaf_popcount", (size_t)L.n * 68) leaf_popcount_kernel<<<(L.n + 127) / 128, 128, 0, w->stream>>>(L.dof.p, L.n, leafCnt.p);
        check_launch("leaf_popcount");
        exclusive_scan_u32(w, leafCnt.p, leafCnt.p, L.n + 1, nullptr);
        ellCols.alloc((size_t)7 * ndofPad, w->stream);
        ellVals.alloc((size_t)7 * ndofPad, w->stream);
        ellCols.fill_bytes(0xff);
        ellVals.zero();
        rowOfVoxel.alloc((size_t)L.n * LEAF, w->stream);
        FB_LAUNCH(w, "mg_ell_build", (size_t)L.n * LEAF * 24) ell_build_kernel<<<L.n, 512, 0, w->stream>>>(view_of(L), leafCnt.p, ndofPad, ellCols.p, ellVals.p, rowOfVoxel.p);
        check_launch("ell_build");
        int total = 7 * ndofPad;
        FB_LAUNCH(w, "mg_ell_resolve", (size_t)total * 8) ell_resolve_kernel<<<(total + 255) / 256, 256, 0, w->stream>>>(ellCols.p, total, rowOfVoxel.p);
        check_launch("ell_resolve");
        // the bottom starts at the first level from which everything fits in shared memory and the op list fits
        const int nl = (int)levels.size();
        bottomFirst = nl - 1;
        bottomInSmem = bottom_need(nl - 1) <= BOTTOM_SMEM_CAP;
        if (bottomInSmem) {
            while (bottomFirst > 0 && nl - (bottomFirst - 1) <= 5 && bottom_need(bottomFirst - 1) <= BOTTOM_SMEM_CAP &&
                   bottom_leaves(bottomFirst - 1) <= BOTTOM_MAX_TOTAL_LEAVES) bottomFirst--;
            bottomSmem = bottom_need(bottomFirst);
        } else {
            while (bottomFirst > 0 && bottom_leaves(bottomFirst - 1) <= BOTTOM_MAX_TOTAL_LEAVES && nl - (bottomFirst - 1) <= 5) bottomFirst--;
            bottomSmem = (size_t)5 * ndofPad * sizeof(float);
        }
        bottomSmem = std::max<size_t>(bottomSmem, 1024);
        FB_CUDA(cudaFuncSetAttribute(mg_bottom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bottomSmem));
    }
    // op list of one visit of bottom level li (index inside the bottom): muCyclePreconditioner<2,skip_first>
    // (uaamg.cpp:1993-2126) when precond, muCycleIterative<2> (:2127-2288) otherwise
    void emit(std::vector<uint8_t>& ops, int li, int nBottom, int n, bool skipFirst, bool precond, int postSmooth, bool isTop) {
        auto put = [&](int code) { ops.push_back((uint8_t)(code | (li << 3))); };
        const bool coarsest = li == nBottom - 1;
        if (precond) {
            if (coarsest) { put(OP_COARSE); return; }
            if (skipFirst) { put(OP_ZERO_RED); put(OP_BLACK); }
            for (int i = (skipFirst ? 1 : 0); i < n; i++) { put(OP_RED); put(OP_BLACK); }
            put(OP_RESID_RESTRICT);
            emit(ops, li + 1, nBottom, n, true, true, 0, false);
            emit(ops, li + 1, nBottom, n, false, true, 0, false);
            put(OP_PROLONG);
            for (int i = 0; i < n; i++) { put(OP_BLACK); put(OP_RED); }
        } else {
            if (coarsest) { for (int i = 0; i < 10 * n; i++) { put(OP_RED); put(OP_BLACK); } return; }
            for (int i = 0; i < n; i++) { put(OP_RED); put(OP_BLACK); }
            put(OP_RESID_RESTRICT);
            for (int mu = 0; mu < 2; mu++) emit(ops, li + 1, nBottom, n, false, false, 0, false);
            put(OP_PROLONG);
            for (int i = 0; i < n; i++) { put(OP_BLACK); put(OP_RED); }
            for (int i = 0; i < postSmooth && isTop; i++) { put(OP_BLACK); put(OP_RED); }
        }
    }
    // op list of muCyclePreconditioner<2, true>(level 0, n) over ALL levels (level index = absolute level)
    void emit_cycle(std::vector<uint8_t>& ops, int level, int n, bool skipFirst) {
        const int nl = (int)levels.size();
        auto put = [&](int code) { ops.push_back((uint8_t)(code | (level << 3))); };
        // the second visit of the coarsest level solves the same right-hand side from a zero guess again
        // (uaamg.cpp:2019-2023: Eigen solve(), not solveWithGuess) -- same bits, so it is issued once
        if (level == nl - 1) { if (skipFirst) put(OP_COARSE); return; }
        if (skipFirst) { put(OP_ZERO_RED); put(OP_BLACK); }
        for (int i = (skipFirst ? 1 : 0); i < n; i++) { put(OP_RED); put(OP_BLACK); }
        put(OP_RESID_RESTRICT);
        emit_cycle(ops, level + 1, n, true);
        emit_cycle(ops, level + 1, n, false);
        put(OP_PROLONG);
        for (int i = 0; i < n; i++) { put(OP_BLACK); put(OP_RED); }
    }
    // ---- mg_cycle_kernel: decide which levels run as compact rows in CTA 0's shared memory and build them
    // mandatory shared memory with levels >= first compact: inv, minus, cols, x, b per row (36 B), the coarsest
    // level's diag and the CG's P and T; diag / parent / child of the other levels are resident only if they fit
    size_t compact_need(int first) const {
        const int nl = (int)levels.size();
        size_t persistent = 0, scratch = 0;
        for (int l = first; l < nl; l++) {
            const int np = compact_np(levels[l]->numDof);
            persistent += (size_t)28 * np;
            scratch += (size_t)8 * np;
        }
        const int npc = compact_np(levels[nl - 1]->numDof);
        persistent += (size_t)4 * npc;
        scratch += (size_t)8 * npc;
        return persistent + std::max<size_t>(scratch, CYC_GRID_SCRATCH);
    }
    void prepare_cycle(int n) {
        const int nl = (int)levels.size();
        cycleReady = false;
        if (nl > CYC_MAX_LEVELS || nl > 31) return;
        int dev = 0, sms = 0, perSm = 0, coop = 0, optin = 0;
        FB_CUDA(cudaGetDevice(&dev));
        FB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        FB_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cudaFuncAttributes fa;
        FB_CUDA(cudaFuncGetAttributes(&fa, mg_cycle_kernel));
        const size_t cap = (size_t)optin - fa.sharedSizeBytes - 256;
        const int maxRows = 65000;   // 16-bit row indices
        if (levels[nl - 1]->numDof > maxRows || compact_need(nl - 1) > cap) return;
        compactFirst = nl - 1;
        while (compactFirst > 0 && levels[compactFirst - 1]->numDof <= maxRows && compact_need(compactFirst - 1) <= cap) compactFirst--;
        if (const char* e = getenv("FLIPB200_COMPACT_FIRST")) compactFirst = std::max(compactFirst, std::min(nl - 1, atoi(e)));
        // rows: per-leaf red / black DOF prefixes of every compact level, then the blobs
        compact.clear();
        compact.resize(nl - compactFirst);
        for (int l = compactFirst; l < nl; l++) {
            Level& L = *levels[l];
            CompactHost& H = compact[l - compactFirst];
            H.n = L.numDof; H.np = compact_np(H.n); H.hasChild = l > compactFirst;
            H.redStart.alloc(L.n + 1, w->stream); H.blackStart.alloc(L.n + 1, w->stream);
            H.redStart.zero(); H.blackStart.zero();
            FB_LAUNCH(w, "mg_leaf_popcount", (size_t)L.n * 72) leaf_colour_count_kernel<<<(L.n + 127) / 128, 128, 0, w->stream>>>(L.dof.p, L.n, H.redStart.p, H.blackStart.p);
            exclusive_scan_u32(w, H.redStart.p, H.redStart.p, L.n + 1, nullptr);
            exclusive_scan_u32(w, H.blackStart.p, H.blackStart.p, L.n + 1, nullptr);
            H.blob.alloc(blob_layout(H.np, H.hasChild).bytes, w->stream);
            H.blob.zero();
            H.voxelOfRow.alloc(H.np, w->stream);
        }
        cluster_rank_leaves();   // its counts come back with the same host wait
        {   // red row counts of all compact levels (= the last entry of each exclusive scan) in one host wait
            uint32_t* nr = reinterpret_cast<uint32_t*>(w->hostScratch);
            for (size_t i = 0; i < compact.size(); i++)
                d2h_words(w, &nr[i], compact[i].redStart.p + levels[compactFirst + i]->n, 4);
            sync(w);
            for (size_t i = 0; i < compact.size(); i++) compact[i].nRed = (int)nr[i];
        }
        auto rowmap = [&](int l) {
            const CompactHost& H = compact[l - compactFirst];
            return RowMap{levels[l]->dof.p, H.redStart.p, H.blackStart.p, H.nRed};
        };
        const RowMap none{nullptr, nullptr, nullptr, 0};
        for (int l = compactFirst; l < nl; l++) {
            Level& L = *levels[l];
            CompactHost& H = compact[l - compactFirst];
            const bool hasF = l > compactFirst, hasC = l + 1 < nl;
            const LevelView lv = view_of(L);
            FB_LAUNCH(w, "mg_compact_build", (size_t)L.n * LEAF * 24)
                compact_build_kernel<<<L.n, 512, 0, w->stream>>>(lv, rowmap(l), H.n, H.np,
                    hasF ? levels[l - 1]->topo->view() : lv.t, hasF ? rowmap(l - 1) : none, hasF ? compact[l - 1 - compactFirst].n : 0,
                    hasC ? levels[l + 1]->topo->view() : lv.t, hasC ? rowmap(l + 1) : none,
                    H.blob.p, H.voxelOfRow.p);
        }
        check_launch("compact_build");
        // shared-memory map: [resident sections][scratch: x, b of every compact level, CG P/T | leaf tiles of the grid ops]
        size_t cursor = 0;
        for (auto& H : compact) {
            H.oInv = (int)cursor; cursor += (size_t)4 * H.np;
            H.oMinus = (int)cursor; cursor += (size_t)12 * H.np;
            H.oCols = (int)cursor; cursor += (size_t)12 * H.np;
            H.oDiag = H.oParent = H.oChild = -1;
        }
        { CompactHost& C = compact.back(); C.oDiag = (int)cursor; cursor += (size_t)4 * C.np; }
        size_t scratch = 0;
        for (auto& H : compact) scratch += (size_t)8 * H.np;
        scratch += (size_t)8 * compact.back().np;
        scratch = std::max<size_t>(scratch, CYC_GRID_SCRATCH);
        // optional sections, most frequently read first
        auto fits = [&](size_t bytes) { return cursor + bytes + scratch <= cap; };
        for (auto& H : compact) if (H.hasChild && fits((size_t)16 * H.np)) { H.oChild = (int)cursor; cursor += (size_t)16 * H.np; }
        for (size_t i = 0; i + 1 < compact.size(); i++) { auto& H = compact[i]; if (fits((size_t)2 * H.np)) { H.oParent = (int)cursor; cursor += (size_t)2 * H.np; } }
        for (auto& H : compact) if (H.oDiag < 0 && fits((size_t)4 * H.np)) { H.oDiag = (int)cursor; cursor += (size_t)4 * H.np; }
        scratchOff = (int)cursor;
        for (auto& H : compact) { H.xOff = (int)cursor; cursor += (size_t)4 * H.np; H.bOff = (int)cursor; cursor += (size_t)4 * H.np; }
        cgOff = (int)cursor; cursor += (size_t)8 * compact.back().np;
        cycleSmem = std::max<size_t>(cursor, (size_t)scratchOff + CYC_GRID_SCRATCH);
        if (cycleSmem > cap) return;
        std::vector<uint8_t>& ops = hostOps;    // members: the asynchronous copies below read them after this returns
        std::vector<uint8_t>& ops2 = hostOps2;
        ops.clear(); ops2.clear();
        emit_cycle(ops, 0, n, true);
        cycleOps = (int)ops.size();
        cycleProg.alloc(ops.size(), w->stream);
        FB_CUDA(cudaMemcpyAsync(cycleProg.p, ops.data(), ops.size(), cudaMemcpyHostToDevice, w->stream));
        emit_cycle(ops2, 0, n, false);
        cycleOps2 = (int)ops2.size();
        cycleProg2.alloc(ops2.size() + 1, w->stream);
        if (cycleOps2) FB_CUDA(cudaMemcpyAsync(cycleProg2.p, ops2.data(), ops2.size(), cudaMemcpyHostToDevice, w->stream));
        cycleBarrier.alloc(1, w->stream);
        FB_CUDA(cudaFuncSetAttribute(mg_cycle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cycleSmem));
        FB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, mg_cycle_kernel, BOT_THREADS, cycleSmem));
        if (tiles) cluster_prepare(n);
        if (!coop || perSm < 1) return;
        cycleGrid = compactFirst == 0 ? 1 : std::min(sms, (int)cycleGridMax);
        cycleReady = true;
        if (tiles) hybrid_prepare(n);
    }
    // ---- mg_cluster_kernel: which cluster size the part gives us (16 is the non-portable maximum)
    static int cluster_size(size_t smem) {
        static int cached = -1;
        if (cached >= 0) return cached;
        cached = 0;
        if (const char* e = getenv("FLIPB200_CLUSTER")) { const int v = atoi(e); if (v == 0) return cached; }
        if (cudaFuncSetAttribute(mg_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); }
        if (cudaFuncSetAttribute(mg_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return cached; }
        int want = 16;
        if (const char* e = getenv("FLIPB200_CLUSTER")) want = atoi(e);
        for (int c = want; c >= 2; c >>= 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(c); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, mg_cluster_kernel, &cfg) == cudaSuccess && nc >= 1) { cached = c; break; }
            cudaGetLastError();
        }
        return cached;
    }
    size_t cluster_cap() const {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, mg_cluster_kernel) != cudaSuccess) { cudaGetLastError(); return 0; }
        return (size_t)optin - fa.sharedSizeBytes - 256;
    }
    // step 1 (before the host wait of prepare_cycle): rank the leaves of every candidate level
    void cluster_rank_leaves() {
        clusterReady = false;
        const int nl = (int)levels.size();
        clHost.clear();
        clHost.resize(nl);
        clSize = cluster_size(cluster_cap());
        if (clSize < 2) return;
        clCandFirst = nl;
        clCounts.alloc((size_t)2 * nl, w->stream);
        clCounts.zero();
        for (int l = nl - 1; l >= 0 && nl - l <= CL_MAX_LEVELS && levels[l]->n <= 1024; l--) {
            clCandFirst = l;
            clHost[l].assign.alloc(levels[l]->n, w->stream);
            FB_LAUNCH(w, "mg_cluster_assign", (size_t)levels[l]->n * 132)
                cl_assign_kernel<<<1, 1024, 0, w->stream>>>(levels[l]->info.p, levels[l]->n, clSize - 1, clHost[l].assign.p, clCounts.p + 2 * l);
        }
        check_launch("cl_assign");
        if (clCandFirst < nl) d2h_words(w, w->hostScratch + 256, clCounts.p, (size_t)8 * nl);
    }
    void emit_cl(std::vector<uint8_t>& ops, int li, int nb, int n, bool skipFirst) {
        auto put = [&](int code) { ops.push_back((uint8_t)(code | (li << 3))); };
        if (li == nb - 1) { if (skipFirst) put(OP_COARSE); return; }
        if (skipFirst) { put(OP_ZERO_RED); put(OP_BLACK); }
        for (int i = (skipFirst ? 1 : 0); i < n; i++) { put(OP_RED); put(OP_BLACK); }
        put(OP_RESID_RESTRICT);
        emit_cl(ops, li + 1, nb, n, true);
        emit_cl(ops, li + 1, nb, n, false);
        put(OP_PROLONG);
        for (int i = 0; i < n; i++) { put(OP_BLACK); put(OP_RED); }
    }
    void emit_hybrid(std::vector<std::vector<uint8_t>>& segs, int level, int n, bool skipFirst) {
        auto put = [&](int code) { segs.back().push_back((uint8_t)(code | (level << 3))); };
        if (skipFirst) { put(OP_ZERO_RED); put(OP_BLACK); }
        for (int i = (skipFirst ? 1 : 0); i < n; i++) { put(OP_RED); put(OP_BLACK); }
        put(OP_RESID_RESTRICT);
        if (level + 1 == clFirst) segs.emplace_back();   // both visits of level clFirst run in the cluster kernel, between two segments
        else { emit_hybrid(segs, level + 1, n, true); emit_hybrid(segs, level + 1, n, false); }
        put(OP_PROLONG);
        for (int i = 0; i < n; i++) { put(OP_BLACK); put(OP_RED); }
    }
    void hybrid_prepare(int n) {
        hybridReady = false;
        if (!cycleReady || !clusterReady || clFirst <= 0 || clFirst > compactFirst) return;
        hybridHost.clear(); hybridSeg.clear(); hybridSeg2.clear();
        for (int v = 0; v < 2; v++) {
            std::vector<std::vector<uint8_t>> segs(1);
            emit_hybrid(segs, 0, n, v == 0);
            auto& dst = v == 0 ? hybridSeg : hybridSeg2;
            for (auto& sg : segs) { dst.emplace_back((int)hybridHost.size(), (int)sg.size()); hybridHost.insert(hybridHost.end(), sg.begin(), sg.end()); }
        }
        hybridProg.alloc(hybridHost.size() + 1, w->stream);
        FB_CUDA(cudaMemcpyAsync(hybridProg.p, hybridHost.data(), hybridHost.size(), cudaMemcpyHostToDevice, w->stream));
        hybridReady = true;
    }
    void launch_hybrid(float* x, const float* b, bool skipFirst) {
        const auto& seg = skipFirst ? hybridSeg : hybridSeg2;
        for (size_t sgi = 0; sgi < seg.size(); sgi++) {
            if (seg[sgi].second) launch_cycle(x, b, false, hybridProg.p + seg[sgi].first, seg[sgi].second);
            if (sgi + 1 < seg.size()) launch_cluster(levels[clFirst]->x.p, levels[clFirst]->b.p, 2);
        }
    }
    // step 2 (after the host wait): how many levels fit, shared-memory map, op lists
    void cluster_prepare(int n) {
        const int nl = (int)levels.size();
        if (clSize < 2 || clCandFirst >= nl || compact.empty()) return;
        const int* cnt = reinterpret_cast<const int*>(w->hostScratch + 256);
        for (int l = clCandFirst; l < nl; l++) { clHost[l].nN = cnt[2 * l]; clHost[l].nC = cnt[2 * l + 1]; }
        const size_t cap = cluster_cap();
        const int ranks = clSize - 1;
        const int npc = compact.back().np;
        const size_t cgBytes = (size_t)48 * npc;
        const size_t sresBytes = (size_t)CL_GROUPS * LEAF * 4 + CL_XB_BYTES + CL_MAX_OPS;
        if (cgBytes + CL_MAX_OPS > cap) return;
        auto need = [&](int first) {
            size_t b = sresBytes;
            for (int l = first; l < nl; l++) {
                const ClusterHost& H = clHost[l];
                const size_t per = (H.nN + ranks - 1) / ranks, perC = (H.nC + ranks - 1) / ranks;
                b += (per + perC) * (CL_XB_BYTES + sizeof(ClSlot)) + per * CL_COEF_BYTES;
            }
            return b;
        };
        if (need(nl - 1) > cap) return;
        clFirst = nl - 1;
        while (clFirst > clCandFirst && need(clFirst - 1) <= cap) clFirst--;
        if (const char* e = getenv("FLIPB200_CLUSTER_FIRST")) clFirst = std::max(clFirst, std::min(nl - 1, atoi(e)));
        size_t cursor = 0;
        for (int l = clFirst; l < nl; l++) {
            ClusterHost& H = clHost[l];
            H.coefPer = (H.nN + ranks - 1) / ranks;
            H.xbPer = H.coefPer + (H.nC + ranks - 1) / ranks;
            H.xbOff = (int)cursor; cursor += (size_t)H.xbPer * CL_XB_BYTES;
        }
        clXbBytes = (int)cursor;
        for (int l = clFirst; l < nl; l++) { ClusterHost& H = clHost[l]; H.coefOff = (int)cursor; cursor += (size_t)H.coefPer * CL_COEF_BYTES; }
        for (int l = clFirst; l < nl; l++) { ClusterHost& H = clHost[l]; H.metaOff = (int)cursor; cursor += (size_t)H.xbPer * sizeof(ClSlot); }
        clSresOff = (int)cursor; cursor += (size_t)CL_GROUPS * LEAF * 4;
        clDumpOff = (int)cursor; cursor += CL_XB_BYTES;
        clCgOff = 0;
        cursor = std::max(cursor, (cgBytes + 15) & ~(size_t)15);
        clProgOff = (int)cursor; cursor += CL_MAX_OPS;
        clSmem = cursor;
        if (clSmem > cap) return;
        for (int v = 0; v < 3; v++) {
            std::vector<uint8_t>& ops = clHostOps[v];
            ops.clear();
            if (v != 1) emit_cl(ops, 0, nl - clFirst, n, true);
            if (v != 0) emit_cl(ops, 0, nl - clFirst, n, false);
            clOps[v] = (int)ops.size();
            if (clOps[v] > CL_MAX_OPS) return;
            clProg[v].alloc(ops.size() + 1, w->stream);
            if (!ops.empty()) FB_CUDA(cudaMemcpyAsync(clProg[v].p, ops.data(), ops.size(), cudaMemcpyHostToDevice, w->stream));
        }
        for (int l = clFirst; l < nl; l++) {   // the per-slot records (neighbours, parent, masks), once per solve
            ClusterHost& H = clHost[l];
            H.meta.alloc((size_t)clSize * H.xbPer, w->stream);
            H.meta.fill_bytes(0xff);
            const bool hasC = l + 1 < nl;
            FB_LAUNCH(w, "mg_cluster_meta", (size_t)levels[l]->n * 400)
                cl_meta_kernel<<<(levels[l]->n + 127) / 128, 128, 0, w->stream>>>(view_of(*levels[l]), H.assign.p, levels[l]->n, H.xbPer, H.meta.p,
                                                                                   view_of(*levels[hasC ? l + 1 : l]), hasC ? clHost[l + 1].assign.p : nullptr);
        }
        check_launch("cl_meta");
        FB_CUDA(cudaFuncSetAttribute(mg_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster_cap()));
        clusterReady = true;
    }
    // variant 0: one visit of level clFirst from a zero guess, 1: one visit from the iterate in x, 2: both in a row
    void launch_cluster(float* x, const float* b, int variant) {
        if (clOps[variant] == 0) return;
        const int nl = (int)levels.size();
        ClusterParams P;
        memset(&P, 0, sizeof(P));
        for (int l = clFirst; l < nl; l++) {
            ClLevel& D = P.lv[l - clFirst];
            const ClusterHost& H = clHost[l];
            D.v = view_of(*levels[l]);
            D.assign = H.assign.p; D.meta = H.meta.p;
            D.n = levels[l]->n; D.xbPer = H.xbPer; D.coefPer = H.coefPer; D.nN = H.nN; D.nC = H.nC;
            D.xbOff = H.xbOff; D.coefOff = H.coefOff; D.metaOff = H.metaOff;
        }
        P.nLevels = nl - clFirst; P.nOps = clOps[variant]; P.prog = clProg[variant].p;
        P.topX = x; P.topB = b; P.loadX = variant == 1 ? 1 : 0;
        P.xbBytes = clXbBytes;
        P.w = 1.2f; P.oneMinusW = 1.0f - 1.2f; P.prolongAlpha = 1.0f;
        const CompactHost& H = compact.back();
        P.cg = CompactDev{H.n, H.np, H.nRed, H.hasChild ? 1 : 0, 0, 0, 0, 0, 0, 0, 0, 0, H.blob.p, H.voxelOfRow.p};
        P.cgOff = clCgOff; P.sresOff = clSresOff; P.progOff = clProgOff; P.dumpOff = clDumpOff;
        P.cgCompat = getenv("FLIPB200_CG_COMPAT") ? atoi(getenv("FLIPB200_CG_COMPAT")) : 0;
        static const char* tracePathEnv = getenv("FLIPB200_TRACE_CLUSTER");
        static int traced = 0;
        DBuf<unsigned long long> trace;
        const bool doTrace = tracePathEnv && variant == 2 && traced++ == 8;
        if (doTrace) { trace.alloc(5 * P.nOps + 4, w->stream); trace.zero(); P.trace = trace.p; }
        static const int dbg = getenv("FLIPB200_CLUSTER_DBG") ? atoi(getenv("FLIPB200_CLUSTER_DBG")) : 0;
        P.dbg = dbg;
        uint64_t bytes = 0;
        for (int l = clFirst; l < nl; l++) bytes += (((uint64_t)levels[l]->numDof * 121) << (l - clFirst)) * (variant == 2 ? 2 : 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clSize); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = clSmem; cfg.stream = w->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = clSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        FB_LAUNCH(w, "mg_cluster", bytes) FB_CUDA(cudaLaunchKernelEx(&cfg, mg_cluster_kernel, P));
        check_launch("mg_cluster");
        if (doTrace) {
            std::vector<unsigned long long> t(5 * P.nOps + 4);
            FB_CUDA(cudaMemcpyAsync(t.data(), trace.p, t.size() * 8, cudaMemcpyDeviceToHost, w->stream));
            sync(w);
            if (FILE* f = fopen(tracePathEnv, "w")) {
                fprintf(f, "# staging %llu ns, ops %llu ns\n", t[2 * P.nOps + 3] - t[2 * P.nOps + 2], t[P.nOps] - t[0]);
                fprintf(f, "k,op,level,leaves,dofs,ns,work_ns,cyc_dispatch,cyc_work,cyc_barrier\n");
                for (int k = 0; k < P.nOps; k++) {
                    const int l = clFirst + (clHostOps[variant][k] >> 3);
                    fprintf(f, "%d,%d,%d,%d,%d,%llu,%lld,%llu,%llu,%llu\n", k, clHostOps[variant][k] & 7, l, levels[l]->n, levels[l]->numDof, t[k + 1] - t[k],
                            (long long)(t[P.nOps + 1 + k] - t[k]), t[2 * P.nOps + 4 + 3 * k], t[2 * P.nOps + 4 + 3 * k + 1], t[2 * P.nOps + 4 + 3 * k + 2]);
                }
                fclose(f);
            }
        }
    }
    void launch_cycle(float* x, const float* b, bool second = false, const uint8_t* segProg = nullptr, int segOps = 0) {
        const int nl = (int)levels.size();
        if (!segProg && second && cycleOps2 == 0) return;
        CycleParams P;
        memset(&P, 0, sizeof(P));
        for (int i = 0; i < nl; i++) {
            Level& L = *levels[i];
            P.lv[i].v = view_of(L);
            P.lv[i].x = i == 0 ? x : L.x.p;
            P.lv[i].b = i == 0 ? const_cast<float*>(b) : L.b.p;
            P.lv[i].n = L.n;
            P.lv[i].xoff = P.lv[i].boff = -1;
            P.lv[i].bReadOnly = i == 0 ? 1 : 0;  // level 0 iterates on the caller's residual, never written in the launch
            P.lv[i].hasParent = L.hasParent ? 1 : 0;
            if (i >= compactFirst) {
                const CompactHost& H = compact[i - compactFirst];
                P.cl[i] = CompactDev{H.n, H.np, H.nRed, H.hasChild ? 1 : 0, H.oInv, H.oMinus, H.oCols, H.oDiag, H.oParent, H.oChild,
                                     H.xOff, H.bOff, H.blob.p, H.voxelOfRow.p};
            }
        }
        P.prog = second ? cycleProg2.p : cycleProg.p; P.nOps = second ? cycleOps2 : cycleOps; P.nLevels = nl; P.compactFirst = compactFirst;
        if (segProg) { P.prog = segProg; P.nOps = segOps; P.gridOnly = 1; }
        static const int cacheInfo = getenv("FLIPB200_MG_INFO_CACHE") ? atoi(getenv("FLIPB200_MG_INFO_CACHE")) : 1;
        P.cacheInfo = cacheInfo;
        const int l1Loads = getenv("FLIPB200_MG_L1") ? atoi(getenv("FLIPB200_MG_L1")) : 1;   // read per launch: the tests compare both
        P.l1Loads = l1Loads;
        P.scratchOff = scratchOff; P.cgOff = cgOff;
        P.w = 1.2f; P.oneMinusW = 1.0f - 1.2f; P.prolongAlpha = 1.0f;
        P.barrier = cycleBarrier.p;
        // FLIPB200_TRACE_CYCLE=<file>: per-op device timestamps of the next application, written as CSV (debug aid)
        static const char* tracePathEnv = getenv("FLIPB200_TRACE_CYCLE");
        const char* tracePath = tracePathEnv;
        DBuf<unsigned long long> trace;
        P.trace = nullptr;
        if (second) tracePath = nullptr;
        if (tracePath) { trace.alloc((segProg ? segOps : cycleOps) + 1, w->stream); trace.zero(); P.trace = trace.p; }
        FB_CUDA(cudaMemsetAsync(cycleBarrier.p, 0, sizeof(unsigned), w->stream));
        uint64_t bytes = 0;  // 121 B/DOF per level visit, level l is visited 2^l times
        for (int i = 0; i < nl; i++) bytes += ((uint64_t)levels[i]->numDof * 121) << i;
        if (segProg) {   // a segment's share: the colour passes it holds, 6 B/DOF each (12 B per sweep), + 12 / 4.5 / 8.5 B for the transfers
            bytes = 0;
            for (int k = 0; k < segOps; k++) {
                const int code = hybridHost[(segProg - hybridProg.p) + k] & 7, l = hybridHost[(segProg - hybridProg.p) + k] >> 3;
                bytes += (uint64_t)levels[l]->numDof * (code == OP_RESID_RESTRICT ? 16 : (code == OP_PROLONG ? 9 : 6));
            }
        }
        void* args[] = {(void*)&P};
        FB_LAUNCH(w, segProg ? "mg_cycle_upper" : "mg_cycle", bytes)
            FB_CUDA(cudaLaunchCooperativeKernel((const void*)mg_cycle_kernel, dim3(cycleGrid), dim3(BOT_THREADS), args, cycleSmem, w->stream));
        check_launch("mg_cycle");
        if (tracePath) {
            const int nTr = segProg ? segOps : cycleOps;
            std::vector<unsigned long long> t(nTr + 1);
            std::vector<uint8_t> ops(nTr);
            FB_CUDA(cudaMemcpyAsync(t.data(), trace.p, t.size() * 8, cudaMemcpyDeviceToHost, w->stream));
            FB_CUDA(cudaMemcpyAsync(ops.data(), P.prog, ops.size(), cudaMemcpyDeviceToHost, w->stream));
            sync(w);
            if (FILE* f = fopen(tracePath, segProg ? "a" : "w")) {   // the segments of the hybrid path append (one block per launch)
                fprintf(f, "k,op,level,leaves,dofs,compact,ns\n");
                for (int k = 0; k < nTr; k++)
                    fprintf(f, "%d,%d,%d,%d,%d,%d,%llu\n", k, ops[k] & 7, ops[k] >> 3, levels[ops[k] >> 3]->n, levels[ops[k] >> 3]->numDof,
                            (ops[k] >> 3) >= compactFirst ? 1 : 0, t[k + 1] - t[k]);
                fclose(f);
            }
        }
    }
    void launch_bottom(float* x, const float* b, int n, bool skipFirst, bool precond, int postSmooth) {
        const int nl = (int)levels.size();
        const int nBottom = nl - bottomFirst;
        BottomParams P;
        std::vector<uint8_t> ops;
        emit(ops, 0, nBottom, n, skipFirst, precond, postSmooth, bottomFirst == 0);
        // a program that needs several launches keeps its state in global memory between them
        const bool smem = bottomInSmem && ops.size() <= (size_t)BOT_MAX_OPS;
        int cursor = 5 * ndofPad;
        for (int i = 0; i < nBottom; i++) {
            Level& L = *levels[bottomFirst + i];
            P.lv[i].v = view_of(L);
            P.lv[i].x = i == 0 ? x : L.x.p;
            P.lv[i].b = i == 0 ? const_cast<float*>(b) : L.b.p;
            P.lv[i].n = L.n;
            P.lv[i].xoff = P.lv[i].boff = -1;
            if (smem) {
                P.lv[i].xoff = cursor; cursor += L.n * LEAF;
                if (i > 0) { P.lv[i].boff = cursor; cursor += L.n * LEAF; }
            }
        }
        P.nLevels = nBottom;
        P.loadX = (ops.empty() || (ops[0] & 7) != OP_ZERO_RED) && (ops.empty() || (ops[0] & 7) != OP_COARSE) ? 1 : 0;
        Level& C = *levels.back();
        P.ell = CoarseELL{ndof, ndofPad, C.n * LEAF, 0, ellCols.p, ellVals.p, rowOfVoxel.p};
        P.w = precond ? 1.2f : 1.0f;
        P.oneMinusW = 1.0f - P.w;
        P.prolongAlpha = precond ? 1.0f : 0.5f;
        // long programs (the pure-multigrid fallback smooths 10n times at the coarsest level) go out in pieces
        uint64_t dofs = 0;
        for (int i = 0; i < nBottom; i++) dofs += (uint64_t)levels[bottomFirst + i]->numDof << i;
        for (size_t pos = 0; pos < ops.size(); pos += BOT_MAX_OPS) {
            P.nOps = (int)std::min<size_t>(BOT_MAX_OPS, ops.size() - pos);
            std::copy(ops.begin() + pos, ops.begin() + pos + P.nOps, P.op);
            FB_LAUNCH(w, "mg_bottom", dofs * 121) mg_bottom_kernel<<<1, BOT_THREADS, bottomSmem, w->stream>>>(P);
        }
        check_launch("mg_bottom");
    }
    void rbgs_pass(Level& L, float* x, const float* b, int colour, float wSor) {
        // one colour: read x (own + halo), b, invdiag, write half of x  -> ~14 B/DOF + coefficients
        if (tiles) { FB_LAUNCH(w, "mg_rbgs_rows", (uint64_t)L.numDof * 14) rbgs_rows_kernel<<<(L.n + 3) / 4, 256, 0, w->stream>>>(view_of(L), x, b, colour, wSor, 1.0f - wSor); }
        else { FB_LAUNCH(w, "mg_rbgs", (uint64_t)L.numDof * 14) rbgs_kernel<<<L.n, 256, 0, w->stream>>>(view_of(L), x, b, colour, wSor, 1.0f - wSor); }
    }
    void rbgs(Level& L, float* x, const float* b, bool redFirst, float wSor) {
        rbgs_pass(L, x, b, redFirst ? 0 : 1, wSor);
        rbgs_pass(L, x, b, redFirst ? 1 : 0, wSor);
        check_launch("rbgs");
    }
    void residual_restrict(Level& L, Level& P, const float* x, const float* b, int leafLo = 0, int leafHi = -1) {
        if (leafHi < 0) leafHi = L.n;
        if (leafHi <= leafLo) return;
        FB_LAUNCH(w, "mg_residual_restrict", (uint64_t)L.numDof * 8 + (uint64_t)P.numDof * 4)
            residual_restrict_kernel<<<leafHi - leafLo, 512, 0, w->stream>>>(view_of(L), view_of(P), x, b, P.b.p, leafLo);
        check_launch("residual_restrict");
    }
    void prolong(Level& L, Level& P, float* x, float alpha) {
        FB_LAUNCH(w, "mg_prolong", (uint64_t)L.numDof * 8 + (uint64_t)P.numDof * 4) prolong_kernel<<<L.n, 512, 0, w->stream>>>(view_of(L), view_of(P), x, P.x.p, alpha);
        check_launch("prolong");
    }
    // muCyclePreconditioner<2, skip_first> with the RBGS smoother (uaamg.cpp:1993-2126)
    // slab decomposition, level 0. Whole ghost LEAVES are exchanged, so eight colour passes fit between two exchanges:
    // what pass k computes in the ghost layer is wrong only within k voxels of its far face, and the owned DOFs read
    // just the nearest ghost plane. Per application: r in, the iterate after pre-smoothing, the coarse right-hand side.
    // The eight colour passes between two exchanges as ONE launch of the cycle kernel (a program of level-0 ops only; grid
    // barriers between the passes instead of launch boundaries, leaf records in shared memory, iterate read through L1).
    DBuf<uint8_t> ddProg;
    DBuf<unsigned> ddBarrier;
    std::vector<uint8_t> ddHost;
    int ddGrid = 0, ddDown = 0, ddUp = 0;
    bool ddPassesReady = false;
    void dd_passes_prepare(int n) {
        ddPassesReady = false;
        if (getenv("FLIPB200_DD_PASS_KERNELS")) return;
        int dev = 0, sms = 0, perSm = 0, coop = 0;
        FB_CUDA(cudaGetDevice(&dev));
        FB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        FB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, mg_cycle_kernel, BOT_THREADS, CYC_GRID_SCRATCH));
        if (!coop || perSm < 1) return;
        ddGrid = std::min(sms, (int)cycleGridMax);
        ddHost.clear();
        ddHost.push_back(OP_ZERO_RED); ddHost.push_back(OP_BLACK);                              // down: the sweep from a zero guess,
        for (int i = 1; i < n; i++) { ddHost.push_back(OP_RED); ddHost.push_back(OP_BLACK); }   // then red-first sweeps
        ddDown = (int)ddHost.size();
        for (int i = 0; i < n; i++) { ddHost.push_back(OP_BLACK); ddHost.push_back(OP_RED); }   // up: black-first sweeps
        ddUp = (int)ddHost.size() - ddDown;
        ddProg.alloc(ddHost.size(), w->stream);
        FB_CUDA(cudaMemcpyAsync(ddProg.p, ddHost.data(), ddHost.size(), cudaMemcpyHostToDevice, w->stream));
        ddBarrier.alloc(1, w->stream);
        ddPassesReady = true;
    }
    void launch_dd_passes(float* x, const float* b, bool down) {
        Level& L = *levels[0];
        CycleParams P;
        memset(&P, 0, sizeof(P));
        P.lv[0].v = view_of(L);
        P.lv[0].x = x; P.lv[0].b = const_cast<float*>(b); P.lv[0].n = L.n;
        P.lv[0].xoff = P.lv[0].boff = -1; P.lv[0].bReadOnly = 1;
        P.prog = down ? ddProg.p : ddProg.p + ddDown; P.nOps = down ? ddDown : ddUp;
        P.nLevels = 1; P.compactFirst = 1; P.gridOnly = 1;
        P.w = 1.2f; P.oneMinusW = 1.0f - 1.2f; P.prolongAlpha = 1.0f;
        P.barrier = ddBarrier.p;
        P.cacheInfo = 1; P.l1Loads = 1;
        FB_CUDA(cudaMemsetAsync(ddBarrier.p, 0, sizeof(unsigned), w->stream));
        void* args[] = {(void*)&P};
        FB_LAUNCH(w, "mg_dd_passes", (uint64_t)L.numDof * 6 * P.nOps)
            FB_CUDA(cudaLaunchCooperativeKernel((const void*)mg_cycle_kernel, dim3(ddGrid), dim3(BOT_THREADS), args, CYC_GRID_SCRATCH, w->stream));
        check_launch("mg_dd_passes");
    }
    void dd_cycle0(float* x, const float* b, int n) {
        Level& L = *levels[0];
        Solver& G = *coarse;
        Level& P = *G.levels[0];
        const float wS = 1.2f;
        dd_refresh(w, {DDArray{const_cast<float*>(b), LEAF * 4}}, 1);
        if (ddPassesReady && n == 4) launch_dd_passes(x, b, true);
        else {
            FB_LAUNCH(w, "mg_zero_red", (uint64_t)L.numDof * 10) zero_red_kernel<<<L.n, 512, 0, w->stream>>>(view_of(L), x, b, wS);
            rbgs_pass(L, x, b, 1, wS);
            for (int i = 1; i < n; i++) rbgs(L, x, b, true, wS);
        }
        dd_refresh(w, {DDArray{x, LEAF * 4}}, 1);
        P.b.zero();
        residual_restrict(L, P, x, b, L.ownLo, L.ownHi);
        { FB_PHASE(w, "ppe allreduce coarse rhs"); comm_allreduce(w, P.b.p, (size_t)P.n * LEAF, CT_F32, false); }   // every coarse cell has exactly one non-zero contributor
        { FB_PHASE(w, "ppe coarse cycles (2 visits)");
          G.mu_cycle_precond(P.x.p, P.b.p, 0, n, true);
          G.mu_cycle_precond(P.x.p, P.b.p, 0, n, false); }
        prolong(L, P, x, 1.0f);
        if (ddPassesReady && n == 4) launch_dd_passes(x, b, false);
        else for (int i = 0; i < n; i++) rbgs(L, x, b, false, wS);
    }
    void mu_cycle_precond(float* x, const float* b, int level, int n, bool skipFirst) {
        if (level == 0 && dd) { dd_cycle0(x, b, n); return; }
        const bool cl = tiles && clusterReady && n == 4;
        if (cl && level == 0 && hybridReady && !getenv("FLIPB200_NO_HYBRID")) { launch_hybrid(x, b, skipFirst); return; }
        if (cl && level == clFirst) { launch_cluster(x, b, skipFirst ? 0 : 1); return; }
        if (!cl && level == 0 && n == 4 && cycleReady && (skipFirst || coarseOnly)) { launch_cycle(x, b, !skipFirst); return; }
        if (!(cl && level < clFirst) && level >= bottomFirst) { launch_bottom(x, b, n, skipFirst, true, 0); return; }
        Level& L = *levels[level];
        const float wS = 1.2f;
        if (skipFirst) {
            // setGridToResultAfterFirstRBGS == the red pass of a sweep from a zero guess, then the black pass
            FB_LAUNCH(w, "mg_zero_red", (uint64_t)L.numDof * 10) zero_red_kernel<<<L.n, 512, 0, w->stream>>>(view_of(L), x, b, wS);
            rbgs_pass(L, x, b, 1, wS);
        }
        for (int i = (skipFirst ? 1 : 0); i < n; i++) rbgs(L, x, b, true, wS);
        Level& P = *levels[level + 1];
        residual_restrict(L, P, x, b);
        if (cl && level + 1 == clFirst) launch_cluster(P.x.p, P.b.p, 2);   // both visits of the cluster's top level in one launch
        else {
            mu_cycle_precond(P.x.p, P.b.p, level + 1, n, true);
            mu_cycle_precond(P.x.p, P.b.p, level + 1, n, false);
        }
        prolong(L, P, x, 1.0f);
        for (int i = 0; i < n; i++) rbgs(L, x, b, false, wS);
    }
    // muCycleIterative<2> with RBGS, w = 1 (uaamg.cpp:2127-2288)
    void mu_cycle_iter(float* x, const float* b, int level, int n, int postSmooth) {
        if (level >= bottomFirst) { launch_bottom(x, b, n, false, false, postSmooth); return; }
        Level& L = *levels[level];
        const float wS = 1.0f;
        for (int i = 0; i < n; i++) rbgs(L, x, b, true, wS);
        Level& P = *levels[level + 1];
        residual_restrict(L, P, x, b);
        for (int mu = 0; mu < 2; mu++) mu_cycle_iter(P.x.p, P.b.p, level + 1, n, 0);
        prolong(L, P, x, 0.5f);
        for (int i = 0; i < n; i++) rbgs(L, x, b, false, wS);
        for (int i = 0; i < postSmooth && level == 0; i++) rbgs(L, x, b, false, wS);
    }
    // level-0 vector kernels
    int defer() const { return dd ? FIN_DEFER : 0; }
    // slab decomposition: s[6] holds this rank's partial; all-reduce it, then apply the scalar update
    void finish(int fin, bool isMax) {
        if (!dd) return;
        comm_allreduce(w, scal.p + 6, 1, CT_F32, isMax);
        FB_LAUNCH(w, "pcg_scalar", 32) scalar_fin_kernel<<<1, 1, 0, w->stream>>>(scal.p, fin);
        check_launch("scalar_fin");
    }
    void residual0(Level& L0, float* out, const float* x, const float* b) {
        FB_LAUNCH(w, "pcg_residual", (uint64_t)L0.numDof * 16) apply_kernel<MODE_RESIDUAL><<<std::min(L0.n, RED_GRID), 512, 0, w->stream>>>(view_of(L0), x, b, out, partial.p, counter.p, scal.p, defer());
        check_launch("residual");
        finish(FIN_NU, true);
    }
    void laplacian0(Level& L0, float* out, const float* x) {
        FB_LAUNCH(w, "pcg_laplacian_dot", (uint64_t)L0.numDof * 12) apply_kernel<MODE_LAPLACIAN><<<std::min(L0.n, RED_GRID), 512, 0, w->stream>>>(view_of(L0), x, nullptr, out, partial.p, counter.p, scal.p, defer());
        check_launch("laplacian");
        finish(FIN_SIGMA_ALPHA, false);
    }
    void dot0(Level& L0, const float* a, const float* b, int fin) {
        FB_LAUNCH(w, "pcg_dot", (uint64_t)L0.numDof * 8) dot_kernel<<<std::min(L0.n, RED_GRID), 512, 0, w->stream>>>(view_of(L0), a, b, partial.p, counter.p, scal.p, fin | defer());
        check_launch("dot");
        finish(fin, false);
    }
    void axpy_absmax0(Level& L0, const float* z, float* r) {
        FB_LAUNCH(w, "pcg_axpy_absmax", (uint64_t)L0.numDof * 12) axpy_absmax_kernel<<<std::min(L0.n, RED_GRID), 512, 0, w->stream>>>(view_of(L0), scal.p, z, r, partial.p, counter.p, defer());
        check_launch("axpy_absmax");
        finish(FIN_NU, true);
    }
    float read_scalar(int slot) {
        float h = 0.f;
        read_back(w, &h, scal.p + slot, 4);
        return h;
    }
};
}  // namespace

// AssembleSolvePPE::apply (FF/nosys/SolvePoissonPressureEqn.cpp:23-64)
void solve_ppe(World* w, float dt, float dx, float relTol, int maxIter) {
    ensure_pool(w, {FLIPB200_LIQUID_SDF, FLIPB200_FACE_WEIGHT, FLIPB200_VELOCITY}, false);
    refresh_solid_views(w);
    TopoPtr pool = w->pool;
    SolverStats& st = w->solver;
    st = SolverStats();
    GridF& phi = w->F(FLIPB200_LIQUID_SDF);
    GridV& fw = w->V(FLIPB200_FACE_WEIGHT);
    GridV& vel = w->V(FLIPB200_VELOCITY);
    const int n = pool->n;
    const bool dd = dd_on(w);
    FB_REQUIRE(!dd || n > 0, FLIPB200_ERR_DOMAIN, "slab decomposition: a rank without fluid cannot take part in the solve");
    if (n == 0) return;  // "skip if there is no dof to solve" (FF/FLIP_vdb.cpp:3044-3047)
    int ownLo = 0, ownHi = n;
    if (dd) dd_owned_slots(w, &ownLo, &ownHi);

    auto ph0 = std::make_unique<PhaseTimer>(w, "ppe level-0 matrix");
    Solver S;
    S.w = w;
    S.dt = dt;
    S.dd = dd;
    if (const char* e = getenv("FLIPB200_MG_PATH")) S.tiles = strcmp(e, "cycle") != 0;
    {
        auto Lp = std::make_unique<Level>();
        Level& L = *Lp;
        L.topo = pool; L.n = n; L.dx = dx; L.term = dt / (dx * dx);
        size_t nv = (size_t)n * LEAF;
        L.dof.alloc((size_t)n * 8, w->stream);
        L.diag.alloc(nv, w->stream); L.xe.alloc(nv, w->stream); L.ye.alloc(nv, w->stream); L.ze.alloc(nv, w->stream);
        FB_LAUNCH(w, "mg_build_finest", nv * 36) build_finest_kernel<<<n, 512, 0, w->stream>>>(pool->view(), phi.val.p, phi.bg, phi.mask.p, fw.val[0].p, fw.val[1].p, fw.val[2].p, dt / (dx * dx), L.dof.p, L.diag.p, L.xe.p, L.ye.p, L.ze.p);
        check_launch("build_finest");
        S.finish_level(L);
        if (dd) {
            // reductions and the DOF count cover the owned leaves only; the count is global
            L.ownLo = ownLo; L.ownHi = ownHi;
            DBuf<int> c(1, w->stream);
            const int mine = ownHi > ownLo ? (int)mask_count(w, L.dof.p + (size_t)ownLo * 8, ownHi - ownLo) : 0;
            FB_CUDA(cudaMemcpyAsync(c.p, &mine, 4, cudaMemcpyHostToDevice, w->stream));
            comm_allreduce(w, c.p, 1, CT_I32, false);
            FB_CUDA(cudaMemcpyAsync(&L.numDof, c.p, 4, cudaMemcpyDeviceToHost, w->stream));
            sync(w);
        }
        S.levels.push_back(std::move(Lp));  // level 0 iterates on the PCG vectors; it owns no x/b
    }
    ph0.reset();
    if (dd) { S.dd_build_coarse(); S.dd_passes_prepare(4); }
    else {
        while (S.levels.back()->numDof > MAX_COARSEST) S.coarsen();
        S.build_coarsest();
        if (!getenv("FLIPB200_NO_CYCLE_KERNEL")) S.prepare_cycle(4);
    }
    Level& L0 = *S.levels[0];
    st.levels = dd ? 1 + (int)S.coarse->levels.size() : (int)S.levels.size();
    st.numDof = L0.numDof;
    S.partial.alloc(n, w->stream);
    S.scal.alloc(8, w->stream);
    S.scal.zero();
    S.counter.alloc(1, w->stream);
    S.counter.zero();

    const size_t nv = (size_t)n * LEAF;
    DBuf<float> rhs(nv, w->stream), x(nv, w->stream), r(nv, w->stream), p(nv, w->stream), z(nv, w->stream);
    x.zero(); p.zero(); z.zero(); r.zero();
    TensionArgs T;
    memset(&T, 0, sizeof(T));
    if (w->tensionCoef > 0.f) {   // enable_tension (FF/nosys/SolvePoissonPressureEqn.cpp:45); tension = 2 coef / density (FF/FLIP_vdb.cpp:3052)
        GridF& cv = w->F(FLIPB200_CURVATURE);
        T.on = 1; T.tension = 2 * w->tensionCoef / w->density; T.dtOverDxSqr = dt / (dx * dx);
        T.phi = phi.val.p; T.phiBg = phi.bg;
        if (cv.topo) { T.ct = cv.topo->view(); T.curv = cv.val.p; }
        T.curvBg = cv.bg;
    }
    FB_LAUNCH(w, "mg_rhs", nv * 40) rhs_kernel<<<n, 512, 0, w->stream>>>(pool->view(), L0.dof.p, fw.val[0].p, fw.val[1].p, fw.val[2].p, vel.val[0].p, vel.val[1].p, vel.val[2].p,
                                                                        w->solidVelView[0].p, w->solidVelView[1].p, w->solidVelView[2].p, 1.0f / dx, rhs.p, T);
    check_launch("rhs");

    // solveMultigridPCG (uaamg.cpp:2332-2403)
    FB_PHASE(w, "ppe pcg loop");
    int status = 1, iter = 0;
    S.residual0(L0, r.p, x.p, rhs.p);
    float nu = S.read_scalar(4);
    const float initAbs = nu + 1e-16f;
    float numax = relTol * nu;
    st.history.push_back(nu / initAbs);
    if (nu <= numax) status = 0;
    else {
        S.mu_cycle_precond(p.p, r.p, 0, 4, true);
        S.dot0(L0, p.p, r.p, FIN_RHO_INIT);
        float nuOld = nu;
        const LevelView v0 = view_of(L0);
        for (; iter < maxIter; iter++) {
            if (dd) dd_refresh(w, {DDArray{p.p, LEAF * 4}}, 1);   // A p reads the nearest ghost plane of p
            S.laplacian0(L0, z.p, p.p);  // z = A p, sigma = p.z, alpha = rho / sigma
            S.axpy_absmax0(L0, z.p, r.p);
            nuOld = nu;
            nu = S.read_scalar(4);
            st.history.push_back(nu / initAbs);
            if (nu <= numax) {
                FB_LAUNCH(w, "pcg_update", (uint64_t)L0.numDof * 12) update_kernel<<<n, 512, 0, w->stream>>>(v0, S.scal.p, x.p, p.p, z.p, 1);
                status = 0;
                break;
            }
            if (nu > nuOld && iter > 3) { status = 1; break; }
            S.mu_cycle_precond(z.p, r.p, 0, 4, true);
            S.dot0(L0, z.p, r.p, FIN_RHO_BETA);  // rho_new = z.r, beta = rho_new / rho, rho = rho_new
            FB_LAUNCH(w, "pcg_update", (uint64_t)L0.numDof * 20) update_kernel<<<n, 512, 0, w->stream>>>(v0, S.scal.p, x.p, p.p, z.p, 0);
        }
        check_launch("pcg");
    }
    st.iterations = iter;
    st.status = status;
    if (status != 0 && !dd) {   // (not available under slab decomposition: the status is reported instead)
        // MGPCG failed: warm start from the previous pressure + pure multigrid (FF/FLIP_vdb.cpp:3089-3097)
        GridF& oldP = w->F(FLIPB200_PRESSURE);
        TopoView ot = oldP.topo ? oldP.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
        FB_LAUNCH(w, "pcg_warm_start", nv * 8) warm_start_kernel<<<n, 512, 0, w->stream>>>(pool->view(), L0.dof.p, ot, oldP.val.p, oldP.bg, x.p);
        S.residual0(L0, r.p, x.p, rhs.p);
        nu = S.read_scalar(4);
        numax = relTol * nu;
        if (!(nu <= numax)) {
            for (int it2 = 0; it2 < 100; it2++) {
                S.mu_cycle_iter(x.p, rhs.p, 0, 8, 8);
                S.residual0(L0, r.p, x.p, rhs.p);
                nu = S.read_scalar(4);
                if (nu <= numax) break;
            }
        }
    }
    st.relResidual = nu / initAbs;

    // outputs: Pressure = new grid on the DOF mask, Divergence = RHS grid (FF/FLIP_vdb.cpp:3063-3067,3087,3099)
    GridF np, nd;
    np.topo = pool; np.bg = 0.f; np.val = std::move(x);
    np.mask.alloc((size_t)n * 8, w->stream);
    np.alloc.alloc(n, w->stream); np.alloc.zero();
    FB_CUDA(cudaMemcpyAsync(np.mask.p, L0.dof.p, (size_t)n * 64, cudaMemcpyDeviceToDevice, w->stream));
    nd.topo = pool; nd.bg = 0.f; nd.val = std::move(rhs);
    nd.mask.alloc((size_t)n * 8, w->stream);
    nd.alloc.alloc(n, w->stream); nd.alloc.zero();
    FB_CUDA(cudaMemcpyAsync(nd.mask.p, L0.dof.p, (size_t)n * 64, cudaMemcpyDeviceToDevice, w->stream));
    w->F(FLIPB200_PRESSURE) = std::move(np);
    w->F(FLIPB200_DIVERGENCE) = std::move(nd);
    if (dd && S.coarse && !S.coarse->levels.empty()) {   // keep the global level-1 arrays for the next solve (ownership only: nothing is freed or copied)
        Level& G1 = *S.coarse->levels[0];
        w->ddCoarse[0] = std::move(G1.diag); w->ddCoarse[1] = std::move(G1.xe); w->ddCoarse[2] = std::move(G1.ye); w->ddCoarse[3] = std::move(G1.ze);
    }
    if (dd) {   // one exchange for both grids
        GridF& gp = w->F(FLIPB200_PRESSURE); GridF& gd = w->F(FLIPB200_DIVERGENCE);
        dd_refresh(w, {DDArray{gp.val.p, LEAF * 4}, DDArray{gp.mask.p, 64}, DDArray{gp.alloc.p, 1}, DDArray{gd.val.p, LEAF * 4}, DDArray{gd.mask.p, 64}, DDArray{gd.alloc.p, 1}}, 2);
    }
    sync(w);
}

}  // namespace fb
