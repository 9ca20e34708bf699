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
#include "levelset.cuh"

namespace fb {
namespace {

constexpr int NB_XM = 4, NB_XP = 22, NB_YM = 10, NB_YP = 16, NB_ZM = 12, NB_ZP = 14;  // nbr27 indices
constexpr int MAX_COARSEST = 4000;  // uaamg.cpp:1965
constexpr int RED_THREADS = 256;

struct Level {
    TopoPtr topo;
    int n = 0;
    float dx = 0.f, term = 0.f;
    int numDof = 0;
    DBuf<uint64_t> dof;
    DBuf<float> diag, invdiag, xe, ye, ze;
    DBuf<uint8_t> flags;   // bit0 diag, bit1 x, bit2 y, bit3 z read as the default
    DBuf<float> x, b, tmp;
};
struct LevelView {
    TopoView t;
    const uint64_t* dof;
    const float *diag, *invdiag, *xe, *ye, *ze;
    const uint8_t* flags;
    float term;
};
LevelView view_of(const Level& L) {
    return LevelView{L.topo->view(), L.dof.p, L.diag.p, L.invdiag.p, L.xe.p, L.ye.p, L.ze.p, L.flags.p, L.term};
}

// ---------------------------------------------------------------- reductions (deterministic)
// per-leaf partials are written by the producing kernel; this folds them in a fixed order.
template <bool IS_MAX>
__global__ void __launch_bounds__(RED_THREADS) fold_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    __shared__ float sm[RED_THREADS];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += RED_THREADS) {
        float v = partial[i];
        if (IS_MAX) a = (isfinite(a) ? (isfinite(v) ? fmaxf(a, v) : v) : a);
        else a = __fadd_rn(a, v);
    }
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            float u = sm[threadIdx.x], v = sm[threadIdx.x + s];
            if (IS_MAX) sm[threadIdx.x] = (isfinite(u) ? (isfinite(v) ? fmaxf(u, v) : v) : u);
            else sm[threadIdx.x] = __fadd_rn(u, v);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}
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
__global__ void __launch_bounds__(512) rhs_kernel(TopoView t, const uint64_t* __restrict__ dof,
                                                  const float* __restrict__ fw0, const float* __restrict__ fw1, const float* __restrict__ fw2,
                                                  const float* __restrict__ v0, const float* __restrict__ v1, const float* __restrict__ v2,
                                                  const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                                                  float invdx, float* __restrict__ rhs) {
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    float r = 0.f;
    if (mask_get(dof, leaf, off)) {
        int3 o = t.origin[leaf];
        int g[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
        float weightSum = 0.f;
        bool nonZero = false;
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
        }
        if (!nonZero || (double)weightSum < 0.1) r = 0.f;
    }
    rhs[i] = r;
}

// ---------------------------------------------------------------- stencil kernels
struct Nbr { int xm, xp, ym, yp, zm, zp; };
__device__ __forceinline__ Nbr load_nbr(const TopoView& t, int leaf) {
    const int* nb = t.nbr27 + (size_t)leaf * 27;
    return Nbr{nb[NB_XM], nb[NB_XP], nb[NB_YM], nb[NB_YP], nb[NB_ZM], nb[NB_ZP]};
}
// off-diagonal sum with the reference's association (uaamg.cpp:1044-1049):
// ((x+ c_x+ + x- c_x-) + (y+ c_y+ + y- c_y-)) + (z+ c_z+ + z- c_z-)
template <bool CONST_COEF>
__device__ __forceinline__ float offdiag(const LevelView& L, const float* __restrict__ x, int leaf, int off, const Nbr& nb) {
    const int X = off >> 6, Y = (off >> 3) & 7, Z = off & 7;
    const size_t base = (size_t)leaf * LEAF;
    const float def = -L.term;
    float xp, xm, yp, ym, zp, zm, cxp, cxm, cyp, cym, czp, czm;
    if (X < 7) { xp = x[base + off + 64]; cxp = CONST_COEF ? def : L.xe[base + off + 64]; }
    else { xp = nb.xp >= 0 ? x[(size_t)nb.xp * LEAF + off - 448] : 0.f; cxp = (CONST_COEF || nb.xp < 0) ? def : L.xe[(size_t)nb.xp * LEAF + off - 448]; }
    xm = X > 0 ? x[base + off - 64] : (nb.xm >= 0 ? x[(size_t)nb.xm * LEAF + off + 448] : 0.f);
    cxm = CONST_COEF ? def : L.xe[base + off];
    if (Y < 7) { yp = x[base + off + 8]; cyp = CONST_COEF ? def : L.ye[base + off + 8]; }
    else { yp = nb.yp >= 0 ? x[(size_t)nb.yp * LEAF + off - 56] : 0.f; cyp = (CONST_COEF || nb.yp < 0) ? def : L.ye[(size_t)nb.yp * LEAF + off - 56]; }
    ym = Y > 0 ? x[base + off - 8] : (nb.ym >= 0 ? x[(size_t)nb.ym * LEAF + off + 56] : 0.f);
    cym = CONST_COEF ? def : L.ye[base + off];
    if (Z < 7) { zp = x[base + off + 1]; czp = CONST_COEF ? def : L.ze[base + off + 1]; }
    else { zp = nb.zp >= 0 ? x[(size_t)nb.zp * LEAF + off - 7] : 0.f; czp = (CONST_COEF || nb.zp < 0) ? def : L.ze[(size_t)nb.zp * LEAF + off - 7]; }
    zm = Z > 0 ? x[base + off - 1] : (nb.zm >= 0 ? x[(size_t)nb.zm * LEAF + off + 7] : 0.f);
    czm = CONST_COEF ? def : L.ze[base + off];
    float fx = __fadd_rn(__fmul_rn(xp, cxp), __fmul_rn(xm, cxm));
    float fy = __fadd_rn(__fmul_rn(yp, cyp), __fmul_rn(ym, cym));
    float fz = __fadd_rn(__fmul_rn(zp, czp), __fmul_rn(zm, czm));
    return __fadd_rn(__fadd_rn(fx, fy), fz);
}
// a leaf takes the constant path when its own and its upper neighbours' face leaves read as default
__device__ __forceinline__ bool leaf_const_faces(const LevelView& L, int leaf, const Nbr& nb) {
    uint8_t f = L.flags[leaf];
    if ((f & 14) != 14) return false;
    if (nb.xp >= 0 && !(L.flags[nb.xp] & 2)) return false;
    if (nb.yp >= 0 && !(L.flags[nb.yp] & 4)) return false;
    if (nb.zp >= 0 && !(L.flags[nb.zp] & 8)) return false;
    return true;
}

enum { MODE_LAPLACIAN = 0, MODE_RESIDUAL = 1 };
// y = A x  or  y = b - A x  (uaamg.cpp:1085-1106); optional per-leaf partial of x.y (Laplacian) or
// |y|_inf (Residual) for the fused reductions
template <int MODE>
__global__ void __launch_bounds__(512) apply_kernel(LevelView L, const float* __restrict__ x, const float* __restrict__ b,
                                                    float* __restrict__ y, float* __restrict__ partial) {
    __shared__ float sm16[16];
    const int leaf = blockIdx.x, off = threadIdx.x;
    const size_t i = (size_t)leaf * LEAF + off;
    const uint64_t* m = L.dof + (size_t)leaf * 8;
    uint64_t any = m[0] | m[1] | m[2] | m[3] | m[4] | m[5] | m[6] | m[7];
    if (!any) {  // the reference skips empty rows; vectors stay zero there
        if (partial && threadIdx.x == 0) partial[leaf] = 0.f;
        return;
    }
    const bool on = (m[off >> 6] >> (off & 63)) & 1ull;
    float out = 0.f, red = 0.f;
    if (on) {
        Nbr nb = load_nbr(L.t, leaf);
        float od = leaf_const_faces(L, leaf, nb) ? offdiag<true>(L, x, leaf, off, nb) : offdiag<false>(L, x, leaf, off, nb);
        float xi = x[i];
        float ax = __fmaf_rn(xi, L.diag[i], od);
        if (MODE == MODE_RESIDUAL) { out = __fsub_rn(b[i], ax); red = out; }
        else { out = ax; red = __fmul_rn(xi, ax); }
    }
    y[i] = out;
    if (partial) {
        float r = MODE == MODE_RESIDUAL ? block_absmax_512(red, sm16) : block_sum_512(red, sm16);
        if (threadIdx.x == 0) partial[leaf] = r;
    }
}
// one colour of red-black SOR, in place (uaamg.cpp:1109-1150): x <- fma(x, 1-w, ((b - off) * invdiag) * w)
// 256 threads: each owns one voxel of the colour. colour 0 = red = (x+y+z) even.
__global__ void __launch_bounds__(256) rbgs_kernel(LevelView L, float* __restrict__ x, const float* __restrict__ b, int colour,
                                                   float w, float oneMinusW) {
    const int leaf = blockIdx.x;
    const uint64_t* m = L.dof + (size_t)leaf * 8;
    uint64_t any = m[0] | m[1] | m[2] | m[3] | m[4] | m[5] | m[6] | m[7];
    if (!any) return;
    const int t = threadIdx.x;
    const int X = t >> 5, Y = (t >> 2) & 7;
    const int Z = ((t & 3) << 1) | ((X + Y + colour) & 1);
    const int off = (X << 6) | (Y << 3) | Z;
    if (!((m[X] >> (off & 63)) & 1ull)) return;
    Nbr nb = load_nbr(L.t, leaf);
    float od = leaf_const_faces(L, leaf, nb) ? offdiag<true>(L, x, leaf, off, nb) : offdiag<false>(L, x, leaf, off, nb);
    const size_t i = (size_t)leaf * LEAF + off;
    float tt = __fmul_rn(__fmul_rn(__fsub_rn(b[i], od), L.invdiag[i]), w);
    x[i] = __fmaf_rn(x[i], oneMinusW, tt);
}
// restriction (uaamg.cpp:1773-1833): coarse = 1/8 sum of the active fine 2^3
__global__ void __launch_bounds__(512) restrict_kernel(LevelView F, LevelView C, const float* __restrict__ fine, float* __restrict__ coarse) {
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    if (!mask_get(C.dof, leaf, off)) return;
    int3 o = C.t.origin[leaf];
    int fx = 2 * (o.x + (off >> 6)), fy = 2 * (o.y + ((off >> 3) & 7)), fz = 2 * (o.z + (off & 7));
    int fl = topo_find(F.t, fx, fy, fz);
    if (fl < 0) return;
    int fb = voxel_off(fx, fy, fz);
    float sum = 0.f;
#pragma unroll
    for (int ii = 0; ii < 2; ii++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++)
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
                int fo = fb + 64 * ii + 8 * jj + kk;
                if (mask_get(F.dof, fl, fo)) sum = __fadd_rn(sum, fine[(size_t)fl * LEAF + fo]);
            }
    coarse[i] = __fmul_rn(sum, 0.125f);
}
// prolongation<inplace_add> (uaamg.cpp:1835-1903), gathered per fine voxel: fine += alpha * coarse(parent)
__global__ void __launch_bounds__(512) prolong_kernel(LevelView F, LevelView C, float* __restrict__ fine, const float* __restrict__ coarse, float alpha) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!mask_get(F.dof, leaf, off)) return;
    int3 o = F.t.origin[leaf];
    int gx = o.x + (off >> 6), gy = o.y + ((off >> 3) & 7), gz = o.z + (off & 7);
    int cx = gx >> 1, cy = gy >> 1, cz = gz >> 1;
    int cl = topo_find(C.t, cx, cy, cz);
    if (cl < 0) return;
    int co = voxel_off(cx, cy, cz);
    if (!mask_get(C.dof, cl, co)) return;
    size_t i = (size_t)leaf * LEAF + off;
    fine[i] = __fadd_rn(fine[i], __fmul_rn(alpha, coarse[(size_t)cl * LEAF + co]));
}

// ---------------------------------------------------------------- level-0 vector kernels
// scalars live on the device: s[0]=rho s[1]=sigma s[2]=alpha s[3]=beta s[4]=nu s[5]=rho_new
__global__ void alpha_kernel(float* s) { s[2] = __fdiv_rn(s[0], s[1]); }
__global__ void beta_kernel(float* s) { s[3] = __fdiv_rn(s[5], s[0]); s[0] = s[5]; }
// r -= alpha z ; per-leaf |r|_inf (levelAlphaXPlusY + levelAbsMax, uaamg.cpp:2447-2479)
__global__ void __launch_bounds__(512) axpy_absmax_kernel(const uint64_t* __restrict__ dof, const float* __restrict__ s,
                                                          const float* __restrict__ z, float* __restrict__ r, float* __restrict__ partial) {
    __shared__ float sm16[16];
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    float v = 0.f;
    if (mask_get(dof, leaf, off)) { v = __fadd_rn(r[i], __fmul_rn(-s[2], z[i])); r[i] = v; }
    float m = block_absmax_512(v, sm16);
    if (threadIdx.x == 0) partial[leaf] = m;
}
__global__ void __launch_bounds__(512) dot_kernel(const uint64_t* __restrict__ dof, const float* __restrict__ a,
                                                  const float* __restrict__ b, float* __restrict__ partial) {
    __shared__ float sm16[16];
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t i = (size_t)leaf * LEAF + off;
    float v = mask_get(dof, leaf, off) ? __fmul_rn(a[i], b[i]) : 0.f;
    float m = block_sum_512(v, sm16);
    if (threadIdx.x == 0) partial[leaf] = m;
}
// x += alpha p ; p = z + beta p   (uaamg.cpp:2396-2397); final=1: only the x update (:2375)
__global__ void __launch_bounds__(512) update_kernel(const uint64_t* __restrict__ dof, const float* __restrict__ s,
                                                     float* __restrict__ x, float* __restrict__ p, const float* __restrict__ z, int final) {
    int leaf = blockIdx.x, off = threadIdx.x;
    if (!mask_get(dof, leaf, off)) return;
    size_t i = (size_t)leaf * LEAF + off;
    float pv = p[i];
    x[i] = __fadd_rn(x[i], __fmul_rn(s[2], pv));
    if (!final) p[i] = __fadd_rn(z[i], __fmul_rn(s[3], pv));
}

// ---------------------------------------------------------------- coarsest level
// Compact ELL form of the coarsest matrix (getTriplets, uaamg.cpp:278-355) and a single-CTA
// Jacobi-preconditioned CG, <= 10 iterations, tolerance float epsilon, zero initial guess
// (Eigen::ConjugateGradient defaults, uaamg.cpp:2291-2303,2019-2023; Eigen itself is not in
// the reference tree -> this follows Eigen's published algorithm, parity unpinned).
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
constexpr int CG_THREADS = 1024;
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
__global__ void __launch_bounds__(CG_THREADS) coarse_cg_kernel(int ndof, int ndofPad, const int* __restrict__ cols,
                                                               const float* __restrict__ vals, const int* __restrict__ rowOfVoxel,
                                                               int nVoxels, const float* __restrict__ rhsGrid, float* __restrict__ lhsGrid) {
    extern __shared__ float sm[];
    float* X = sm; float* R = X + ndofPad; float* P = R + ndofPad; float* T = P + ndofPad; float* DI = T + ndofPad;
    __shared__ float sm33[33];
    const int tid = threadIdx.x;
    for (int v = tid; v < nVoxels; v += CG_THREADS) { int r = rowOfVoxel[v]; if (r >= 0) R[r] = rhsGrid[v]; }
    for (int r = tid; r < ndof; r += CG_THREADS) { X[r] = 0.f; float d = vals[r]; DI[r] = d != 0.f ? __fdiv_rn(1.0f, d) : 1.0f; }
    __syncthreads();
    float acc = 0.f;
    for (int r = tid; r < ndof; r += CG_THREADS) acc = __fadd_rn(acc, __fmul_rn(R[r], R[r]));
    float rhsNorm2 = cg_block_sum(acc, sm33);
    if (rhsNorm2 != 0.f) {
        const float tol = 1.1920929e-07f;
        float threshold = fmaxf(__fmul_rn(__fmul_rn(tol, tol), rhsNorm2), 1.17549435e-38f);
        float residualNorm2 = rhsNorm2;
        if (!(residualNorm2 < threshold)) {
            acc = 0.f;
            for (int r = tid; r < ndof; r += CG_THREADS) { float pv = __fmul_rn(DI[r], R[r]); P[r] = pv; acc = __fadd_rn(acc, __fmul_rn(R[r], pv)); }
            float absNew = cg_block_sum(acc, sm33);
            for (int it = 0; it < 10; it++) {
                acc = 0.f;
                for (int r = tid; r < ndof; r += CG_THREADS) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < 7; k++) { int c = cols[k * ndofPad + r]; if (c >= 0) s = __fadd_rn(s, __fmul_rn(vals[k * ndofPad + r], P[c])); }
                    T[r] = s;
                    acc = __fadd_rn(acc, __fmul_rn(P[r], s));
                }
                float pt = cg_block_sum(acc, sm33);
                float alpha = __fdiv_rn(absNew, pt);
                acc = 0.f;
                for (int r = tid; r < ndof; r += CG_THREADS) {
                    X[r] = __fadd_rn(X[r], __fmul_rn(alpha, P[r]));
                    float rv = __fsub_rn(R[r], __fmul_rn(alpha, T[r]));
                    R[r] = rv;
                    acc = __fadd_rn(acc, __fmul_rn(rv, rv));
                }
                residualNorm2 = cg_block_sum(acc, sm33);
                if (residualNorm2 < threshold) break;
                acc = 0.f;
                for (int r = tid; r < ndof; r += CG_THREADS) { float zv = __fmul_rn(DI[r], R[r]); T[r] = zv; acc = __fadd_rn(acc, __fmul_rn(R[r], zv)); }
                float absOld = absNew;
                absNew = cg_block_sum(acc, sm33);
                float beta = __fdiv_rn(absNew, absOld);
                for (int r = tid; r < ndof; r += CG_THREADS) P[r] = __fadd_rn(T[r], __fmul_rn(beta, P[r]));
                __syncthreads();
            }
        }
    }
    __syncthreads();
    for (int v = tid; v < nVoxels; v += CG_THREADS) { int r = rowOfVoxel[v]; if (r >= 0) lhsGrid[v] = X[r]; }
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

// ---------------------------------------------------------------- host-side solver
struct Solver {
    World* w;
    std::vector<std::unique_ptr<Level>> levels;
    // coarsest ELL
    int ndof = 0, ndofPad = 0;
    DBuf<int> ellCols, rowOfVoxel;
    DBuf<float> ellVals;
    DBuf<float> partial, scal;  // [max leaves], [8]

    void alloc_vectors(Level& L) {
        size_t n = (size_t)L.n * LEAF;
        L.x.alloc(n, w->stream); L.b.alloc(n, w->stream); L.tmp.alloc(n, w->stream);
        L.x.zero(); L.b.zero(); L.tmp.zero();
    }
    void finish_level(Level& L) {
        L.invdiag.alloc((size_t)L.n * LEAF, w->stream);
        L.flags.alloc(L.n, w->stream);
        FB_LAUNCH(w, "mg_trim", (size_t)L.n * LEAF * 24) trim_kernel<<<L.n, 512, 0, w->stream>>>(L.dof.p, L.diag.p, L.xe.p, L.ye.p, L.ze.p, L.invdiag.p, L.flags.p, L.term);
        check_launch("trim");
        L.numDof = (int)mask_count(w, L.dof.p, L.n);
    }
    void coarsen() {
        const Level& F = *levels.back();
        DBuf<int3> cand(F.n, w->stream);
        DBuf<uint32_t> counter(1, w->stream);
        counter.zero();
        FB_LAUNCH(w, "mg_coarse_leaves", (size_t)F.n * 76) dof_leaf_origins_kernel<<<(F.n + 127) / 128, 128, 0, w->stream>>>(F.topo->view(), F.dof.p, cand.p, counter.p);
        check_launch("dof_leaf_origins");
        uint32_t cnt = 0;
        FB_CUDA(cudaMemcpyAsync(&cnt, counter.p, 4, cudaMemcpyDeviceToHost, w->stream));
        sync(w);
        auto Lp = std::make_unique<Level>();
        Level& L = *Lp;
        L.topo = topo_from_origins_dev(w, cand.p, (int)cnt, false);
        L.n = L.topo->n;
        L.dx = 2.0f * F.dx;
        L.term = dt / (L.dx * L.dx);
        size_t n = (size_t)L.n * LEAF;
        L.dof.alloc((size_t)L.n * 8, w->stream);
        L.diag.alloc(n, w->stream); L.xe.alloc(n, w->stream); L.ye.alloc(n, w->stream); L.ze.alloc(n, w->stream);
        FB_LAUNCH(w, "mg_coarsen", (size_t)L.n * LEAF * 16 + (size_t)F.n * LEAF * 16) coarsen_kernel<<<L.n, 512, 0, w->stream>>>(view_of(F), L.topo->view(), L.term, L.dof.p, L.diag.p, L.xe.p, L.ye.p, L.ze.p);
        check_launch("coarsen");
        finish_level(L);
        alloc_vectors(L);
        levels.push_back(std::move(Lp));
    }
    float dt = 0.f;

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
        size_t smemBytes = (size_t)5 * ndofPad * sizeof(float);
        FB_CUDA(cudaFuncSetAttribute(coarse_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smemBytes, 1024)));
    }
    void coarsest_solve(float* lhs, const float* rhs) {
        Level& L = *levels.back();
        size_t smemBytes = (size_t)5 * ndofPad * sizeof(float);
        FB_LAUNCH(w, "mg_coarse_cg", (size_t)ndof * 7 * 8 * 10) coarse_cg_kernel<<<1, CG_THREADS, smemBytes, w->stream>>>(ndof, ndofPad, ellCols.p, ellVals.p, rowOfVoxel.p, L.n * LEAF, rhs, lhs);
        check_launch("coarse_cg");
    }
    void rbgs(Level& L, float* x, const float* b, bool redFirst, float wSor) {
        LevelView v = view_of(L);
        float omw = 1.0f - wSor;
        // one colour: read x (own + halo), b, invdiag, write half of x  -> ~14 B/DOF + coefficients
        uint64_t bytes = (uint64_t)L.numDof * 14;
        FB_LAUNCH(w, "mg_rbgs", bytes) rbgs_kernel<<<L.n, 256, 0, w->stream>>>(v, x, b, redFirst ? 0 : 1, wSor, omw);
        FB_LAUNCH(w, "mg_rbgs", bytes) rbgs_kernel<<<L.n, 256, 0, w->stream>>>(v, x, b, redFirst ? 1 : 0, wSor, omw);
        check_launch("rbgs");
    }
    void residual(Level& L, float* out, const float* x, const float* b, float* partialOut) {
        FB_LAUNCH(w, "mg_residual", (uint64_t)L.numDof * 16) apply_kernel<MODE_RESIDUAL><<<L.n, 512, 0, w->stream>>>(view_of(L), x, b, out, partialOut);
        check_launch("residual");
    }
    void laplacian(Level& L, float* out, const float* x, float* partialOut) {
        FB_LAUNCH(w, "mg_laplacian", (uint64_t)L.numDof * 12) apply_kernel<MODE_LAPLACIAN><<<L.n, 512, 0, w->stream>>>(view_of(L), x, nullptr, out, partialOut);
        check_launch("laplacian");
    }
    // muCyclePreconditioner<2, skip_first> with the RBGS smoother (uaamg.cpp:1993-2126)
    void mu_cycle_precond(float* x, const float* b, int level, int n, bool skipFirst) {
        const int nlevel = (int)levels.size();
        Level& L = *levels[level];
        if (level == nlevel - 1) { coarsest_solve(x, b); return; }
        const float wS = 1.2f;
        if (skipFirst) {
            // setGridToResultAfterFirstRBGS == a red-first sweep from a zero guess (oracle/poisson.cpp)
            FB_CUDA(cudaMemsetAsync(x, 0, (size_t)L.n * LEAF * 4, w->stream));
            rbgs(L, x, b, true, wS);
        }
        for (int i = (skipFirst ? 1 : 0); i < n; i++) rbgs(L, x, b, true, wS);
        residual(L, L.tmp.p, x, b, nullptr);
        Level& P = *levels[level + 1];
        FB_LAUNCH(w, "mg_restrict", (uint64_t)L.numDof * 4 + (uint64_t)P.numDof * 4) restrict_kernel<<<P.n, 512, 0, w->stream>>>(view_of(L), view_of(P), L.tmp.p, P.b.p);
        check_launch("restrict");
        mu_cycle_precond(P.x.p, P.b.p, level + 1, n, true);
        mu_cycle_precond(P.x.p, P.b.p, level + 1, n, false);
        FB_LAUNCH(w, "mg_prolong", (uint64_t)L.numDof * 8 + (uint64_t)P.numDof * 4) prolong_kernel<<<L.n, 512, 0, w->stream>>>(view_of(L), view_of(P), x, P.x.p, 1.0f);
        check_launch("prolong");
        for (int i = 0; i < n; i++) rbgs(L, x, b, false, wS);
    }
    // muCycleIterative<2> with RBGS, w = 1 (uaamg.cpp:2127-2288)
    void mu_cycle_iter(float* x, const float* b, int level, int n, int postSmooth) {
        const int nlevel = (int)levels.size();
        Level& L = *levels[level];
        const float wS = 1.0f;
        if (level == nlevel - 1) { for (int i = 0; i < 10 * n; i++) rbgs(L, x, b, true, wS); return; }
        for (int i = 0; i < n; i++) rbgs(L, x, b, true, wS);
        residual(L, L.tmp.p, x, b, nullptr);
        Level& P = *levels[level + 1];
        FB_LAUNCH(w, "mg_restrict", (uint64_t)L.numDof * 4 + (uint64_t)P.numDof * 4) restrict_kernel<<<P.n, 512, 0, w->stream>>>(view_of(L), view_of(P), L.tmp.p, P.b.p);
        check_launch("restrict");
        for (int mu = 0; mu < 2; mu++) mu_cycle_iter(P.x.p, P.b.p, level + 1, n, 0);
        FB_LAUNCH(w, "mg_prolong", (uint64_t)L.numDof * 8 + (uint64_t)P.numDof * 4) prolong_kernel<<<L.n, 512, 0, w->stream>>>(view_of(L), view_of(P), x, P.x.p, 0.5f);
        check_launch("prolong");
        for (int i = 0; i < n; i++) rbgs(L, x, b, false, wS);
        for (int i = 0; i < postSmooth && level == 0; i++) rbgs(L, x, b, false, wS);
    }
    float fold(bool isMax, int n, int slot) {
        if (isMax) fold_kernel<true><<<1, RED_THREADS, 0, w->stream>>>(partial.p, n, scal.p + slot);
        else fold_kernel<false><<<1, RED_THREADS, 0, w->stream>>>(partial.p, n, scal.p + slot);
        w->launches++;
        check_launch("fold");
        return 0.f;
    }
    float read_scalar(int slot) {
        float h = 0.f;
        FB_CUDA(cudaMemcpyAsync(&h, scal.p + slot, 4, cudaMemcpyDeviceToHost, w->stream));
        sync(w);
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
    if (n == 0) return;  // "skip if there is no dof to solve" (FF/FLIP_vdb.cpp:3044-3047)

    Solver S;
    S.w = w;
    S.dt = dt;
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
        S.alloc_vectors(L);
        S.levels.push_back(std::move(Lp));
    }
    while (S.levels.back()->numDof > MAX_COARSEST) S.coarsen();
    S.build_coarsest();
    Level& L0 = *S.levels[0];
    st.levels = (int)S.levels.size();
    st.numDof = L0.numDof;
    int maxLeaves = 0;
    for (auto& L : S.levels) maxLeaves = std::max(maxLeaves, L->n);
    S.partial.alloc(maxLeaves, w->stream);
    S.scal.alloc(8, w->stream);
    S.scal.zero();

    const size_t nv = (size_t)n * LEAF;
    DBuf<float> rhs(nv, w->stream), x(nv, w->stream), r(nv, w->stream), p(nv, w->stream), z(nv, w->stream);
    x.zero(); p.zero(); z.zero();
    FB_LAUNCH(w, "mg_rhs", nv * 40) rhs_kernel<<<n, 512, 0, w->stream>>>(pool->view(), L0.dof.p, fw.val[0].p, fw.val[1].p, fw.val[2].p, vel.val[0].p, vel.val[1].p, vel.val[2].p,
                                                                        w->solidVelView[0].p, w->solidVelView[1].p, w->solidVelView[2].p, 1.0f / dx, rhs.p);
    check_launch("rhs");

    // solveMultigridPCG (uaamg.cpp:2332-2403)
    int status = 1, iter = 0;
    S.residual(L0, r.p, x.p, rhs.p, S.partial.p);
    S.fold(true, n, 4);
    float nu = S.read_scalar(4);
    const float initAbs = nu + 1e-16f;
    float numax = relTol * nu;
    st.history.push_back(nu / initAbs);
    if (nu <= numax) status = 0;
    else {
        S.mu_cycle_precond(p.p, r.p, 0, 4, true);
        FB_LAUNCH(w, "pcg_dot", (uint64_t)L0.numDof * 8) dot_kernel<<<n, 512, 0, w->stream>>>(L0.dof.p, p.p, r.p, S.partial.p);
        S.fold(false, n, 0);  // rho
        float nuOld = nu;
        for (; iter < maxIter; iter++) {
            S.laplacian(L0, z.p, p.p, S.partial.p);
            S.fold(false, n, 1);  // sigma
            alpha_kernel<<<1, 1, 0, w->stream>>>(S.scal.p);
            w->launches++;
            FB_LAUNCH(w, "pcg_axpy_absmax", (uint64_t)L0.numDof * 12) axpy_absmax_kernel<<<n, 512, 0, w->stream>>>(L0.dof.p, S.scal.p, z.p, r.p, S.partial.p);
            S.fold(true, n, 4);
            nuOld = nu;
            nu = S.read_scalar(4);
            st.history.push_back(nu / initAbs);
            if (nu <= numax) {
                FB_LAUNCH(w, "pcg_update", (uint64_t)L0.numDof * 12) update_kernel<<<n, 512, 0, w->stream>>>(L0.dof.p, S.scal.p, x.p, p.p, z.p, 1);
                status = 0;
                break;
            }
            if (nu > nuOld && iter > 3) { status = 1; break; }
            S.mu_cycle_precond(z.p, r.p, 0, 4, true);
            FB_LAUNCH(w, "pcg_dot", (uint64_t)L0.numDof * 8) dot_kernel<<<n, 512, 0, w->stream>>>(L0.dof.p, z.p, r.p, S.partial.p);
            S.fold(false, n, 5);  // rho_new
            beta_kernel<<<1, 1, 0, w->stream>>>(S.scal.p);
            w->launches++;
            FB_LAUNCH(w, "pcg_update", (uint64_t)L0.numDof * 20) update_kernel<<<n, 512, 0, w->stream>>>(L0.dof.p, S.scal.p, x.p, p.p, z.p, 0);
        }
        check_launch("pcg");
    }
    st.iterations = iter;
    st.status = status;
    if (status != 0) {
        // MGPCG failed: warm start from the previous pressure + pure multigrid (FF/FLIP_vdb.cpp:3089-3097)
        GridF& oldP = w->F(FLIPB200_PRESSURE);
        TopoView ot = oldP.topo ? oldP.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
        FB_LAUNCH(w, "pcg_warm_start", nv * 8) warm_start_kernel<<<n, 512, 0, w->stream>>>(pool->view(), L0.dof.p, ot, oldP.val.p, oldP.bg, x.p);
        S.residual(L0, r.p, x.p, rhs.p, S.partial.p);
        S.fold(true, n, 4);
        nu = S.read_scalar(4);
        numax = relTol * nu;
        if (!(nu <= numax)) {
            for (int it2 = 0; it2 < 100; it2++) {
                S.mu_cycle_iter(x.p, rhs.p, 0, 8, 8);
                S.residual(L0, r.p, x.p, rhs.p, S.partial.p);
                S.fold(true, n, 4);
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
    sync(w);
}

}  // namespace fb
