// libflipb200 -- FluidReseed (FF/nosys/FLIP_Reseed.cpp:8-16 -> FLIP_vdb::reseed_fluid, FF/FLIP_vdb.cpp:2047-2220).
//
// Per particle leaf, voxel by voxel in offset order: the voxel keeps its particles; if the liquid SDF at its centre is < dx and it
// holds <= 4 of them, up to 16 jittered candidates are tried while it holds < 8 -- a candidate is taken when the SDF at its
// position is <= -dx and its octant of the voxel is still empty, and gets the StaggeredBoxSampler velocity there. The jitter comes
// from the reference's hash table (FF/FLIP_vdb.h:10-20) read at a running index: every trial consumes three entries whether it
// is taken or not, so where a voxel's draws start depends on every voxel before it in the leaf. That chain is kept:
//   reseed_decide_kernel  one WARP per leaf walks the eligible voxels in order; the 16 trials of a voxel are evaluated by 16
//                         lanes at once (the SDF samples are the expensive part), then resolved in trial order by every lane
//                         identically. Output per voxel: the draw index its trials start at and the 16-bit set of taken trials.
//   (scan of the new per-voxel counts)
//   reseed_write_kernel   one thread per voxel copies the old particles (position decode -> encode once more, as the reference's
//                         write handles do) and appends the taken candidates with their sampled velocity.
// The reference starts each TBB chunk at std::random_device (:2081-2084) and runs on through the chunk's leaves; here a leaf's
// start is a hash of (seed, leaf origin) -- the seeded variant SURVEY 8f-1 asks for; the oracle takes the same starts.
// Both grids share the particles' cell-centred transform (FF/nosys/FLIP_Creator.cpp:37-116): index -> world = ijk * s, world ->
// index = xyz * (1 / s) in double (math/Maps.h ScaleMap), restated literally because it does not round-trip for every ijk.
#include "world.cuh"
#include <climits>
#include <cuda_fp16.h>

namespace fb {
namespace {

__device__ __forceinline__ float rs_frand(unsigned int i) {
    unsigned int value = (i ^ 61u) ^ (i >> 16);
    value *= 9u;
    value ^= value << 4;
    value *= 0x27d4eb2du;
    value ^= value >> 15;
    return __fdiv_rn((float)value, 4294967296.0f);
}
__device__ __forceinline__ float rs_table(unsigned int index) { return __double2float_rn(__dsub_rn((double)rs_frand(index % 21474836u), 0.5)); }
__host__ __device__ __forceinline__ unsigned int rs_leaf_start(uint32_t seed, int ox, int oy, int oz) {
    uint32_t h = seed ^ ((uint32_t)ox * 73856093u) ^ ((uint32_t)oy * 19349663u) ^ ((uint32_t)oz * 83492791u);
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h % 21474836u;
}
// openvdb BoxSampler with double weights (tools/Interpolation.h:712-737): a + float(double(b - a) * w)
__device__ __forceinline__ float rs_ip(float a, float b, double w) { return __fadd_rn(a, __double2float_rn(__dmul_rn((double)__fsub_rn(b, a), w))); }
__device__ __forceinline__ float rs_box(const TopoView& t, const float* __restrict__ val, float bg, double x, double y, double z) {
    const int bx = (int)floor(x), by = (int)floor(y), bz = (int)floor(z);
    const double u = __dsub_rn(x, (double)bx), v = __dsub_rn(y, (double)by), w = __dsub_rn(z, (double)bz);
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = grid_get(t, val, bg, bx + (i >> 2), by + ((i >> 1) & 1), bz + (i & 1));
    return rs_ip(rs_ip(rs_ip(d[0], d[1], w), rs_ip(d[2], d[3], w), v), rs_ip(rs_ip(d[4], d[5], w), rs_ip(d[6], d[7], w), v), u);
}

struct ReseedParams {
    TopoView pt;                 // the particle store's topology
    const uint32_t* voxelStart;
    const uint32_t *w0, *w1, *w2;
    TopoView st; const float* sdf; float sdfBg;
    TopoView vt; const float* vel[3]; float velBg[3];
    float dx; double s, inv;
    uint32_t seed;
    int xLo, xHi;                // slab decomposition: only leaves with xLo <= leaf x < xHi (owned + one ghost layer) are topped up
    uint32_t* tstart;            // [n*512] draw index of the voxel's first trial
    uint16_t* accept;            // [n*512] taken trials
    uint32_t* newCount;          // [n*512 + 1]
    const uint32_t* newStart;    // exclusive prefix of newCount
    uint32_t *o0, *o1, *o2;
};

inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }
constexpr int RS_WARPS = 4;
__global__ void __launch_bounds__(RS_WARPS * 32) reseed_decide_kernel(ReseedParams p) {
    const int lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    if (leaf >= p.pt.n) return;
    const int3 o = p.pt.origin[leaf];
    const uint32_t* vs = p.voxelStart + (size_t)leaf * LEAF;
    // phase 1: which voxels try at all (independent of the draws): <= 4 particles and SDF at the centre < dx
    __shared__ uint32_t sElig[RS_WARPS][16];
    uint32_t* elig = sElig[threadIdx.x >> 5];
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
        const int off = it * 32 + lane;
        const uint32_t cnt = vs[off + 1] - vs[off];
        bool e = false;
        if (cnt <= 4u && (o.x >> 3) >= p.xLo && (o.x >> 3) < p.xHi) {
            const double wx = __dmul_rn((double)(o.x + (off >> 6)), p.s), wy = __dmul_rn((double)(o.y + ((off >> 3) & 7)), p.s), wz = __dmul_rn((double)(o.z + (off & 7)), p.s);
            e = rs_box(p.st, p.sdf, p.sdfBg, __dmul_rn(wx, p.inv), __dmul_rn(wy, p.inv), __dmul_rn(wz, p.inv)) < p.dx;
        }
        const uint32_t eb = __ballot_sync(0xffffffffu, e);
        if (lane == 0) elig[it] = eb;
        p.newCount[(size_t)leaf * LEAF + off] = cnt;
        p.accept[(size_t)leaf * LEAF + off] = 0;
    }
    __syncwarp();
    // phase 2: the eligible voxels in offset order
    unsigned int index = rs_leaf_start(p.seed, o.x, o.y, o.z);
    const float negDx = -p.dx;
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
        uint32_t m = elig[it];
        while (m) {
            const int off = it * 32 + (__ffs(m) - 1);
            m &= m - 1;
            const uint32_t b = vs[off], cnt = vs[off + 1] - b;
            unsigned occBit = 0;
            if ((uint32_t)lane < cnt) {   // cnt <= 4
                const uint32_t a0 = __ldg(&p.w0[b + lane]), a1 = __ldg(&p.w1[b + lane]);
                const float px = fx_decode(a0 & 0xffffu), py = fx_decode(a0 >> 16), pz = fx_decode(a1 & 0xffffu);
                occBit = 1u << (((pz > 0.f) << 2) | ((py > 0.f) << 1) | (px > 0.f ? 1 : 0));
            }
            unsigned occ = __reduce_or_sync(0xffffffffu, occBit);
            const double wx = __dmul_rn((double)(o.x + (off >> 6)), p.s), wy = __dmul_rn((double)(o.y + ((off >> 3) & 7)), p.s), wz = __dmul_rn((double)(o.z + (off & 7)), p.s);
            bool pass = false;
            unsigned sv = 0;
            if (lane < 16) {
                const unsigned int at = index + 3u * (unsigned)lane;
                const float jx = rs_table(at), jy = rs_table(at + 1u), jz = rs_table(at + 2u);
                const double qx = __dadd_rn(__dmul_rn((double)jx, p.s), wx), qy = __dadd_rn(__dmul_rn((double)jy, p.s), wy), qz = __dadd_rn(__dmul_rn((double)jz, p.s), wz);
                const float phi2 = rs_box(p.st, p.sdf, p.sdfBg, __dmul_rn(qx, p.inv), __dmul_rn(qy, p.inv), __dmul_rn(qz, p.inv));
                pass = !(phi2 > negDx);
                sv = ((jz > 0.f) << 2) | ((jy > 0.f) << 1) | (jx > 0.f ? 1 : 0);
            }
            const uint32_t passMask = __ballot_sync(0xffffffffu, pass);
            const uint32_t s0 = __ballot_sync(0xffffffffu, sv & 1u), s1 = __ballot_sync(0xffffffffu, sv & 2u), s2 = __ballot_sync(0xffffffffu, sv & 4u);
            uint32_t here = cnt, used = 0, acc = 0;
            for (int t = 0; t < 16 && here < 8u; t++) {
                used++;
                if ((passMask >> t) & 1u) {
                    const unsigned oc = ((s0 >> t) & 1u) | (((s1 >> t) & 1u) << 1) | (((s2 >> t) & 1u) << 2);
                    if (!((occ >> oc) & 1u)) { occ |= 1u << oc; acc |= 1u << t; here++; }
                }
            }
            if (lane == 0) {
                p.tstart[(size_t)leaf * LEAF + off] = index;
                p.accept[(size_t)leaf * LEAF + off] = (uint16_t)acc;
                p.newCount[(size_t)leaf * LEAF + off] = here;
            }
            index += 3u * used;
        }
    }
}

__global__ void __launch_bounds__(256) reseed_write_kernel(ReseedParams p) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (size_t)p.pt.n * LEAF) return;
    const uint32_t b = p.voxelStart[v], e = p.voxelStart[v + 1];
    uint32_t d = p.newStart[v];
    for (uint32_t i = b; i < e; i++, d++) {   // the write handles re-encode the decoded position; the half velocity round-trips
        const uint32_t a0 = p.w0[i], a1 = p.w1[i];
        p.o0[d] = fx_encode(fx_decode(a0 & 0xffffu)) | (fx_encode(fx_decode(a0 >> 16)) << 16);
        p.o1[d] = fx_encode(fx_decode(a1 & 0xffffu)) | (a1 & 0xffff0000u);
        p.o2[d] = p.w2[i];
    }
    uint32_t acc = p.accept[v];
    if (!acc) return;
    const int leaf = (int)(v >> 9), off = (int)(v & 511);
    const int3 o = p.pt.origin[leaf];
    const double wx = __dmul_rn((double)(o.x + (off >> 6)), p.s), wy = __dmul_rn((double)(o.y + ((off >> 3) & 7)), p.s), wz = __dmul_rn((double)(o.z + (off & 7)), p.s);
    const unsigned int index = p.tstart[v];
    while (acc) {
        const int t = __ffs(acc) - 1;
        acc &= acc - 1;
        const unsigned int at = index + 3u * (unsigned)t;
        const float jx = rs_table(at), jy = rs_table(at + 1u), jz = rs_table(at + 2u);
        const double ix = __dmul_rn(__dadd_rn(__dmul_rn((double)jx, p.s), wx), p.inv), iy = __dmul_rn(__dadd_rn(__dmul_rn((double)jy, p.s), wy), p.inv),
                     iz = __dmul_rn(__dadd_rn(__dmul_rn((double)jz, p.s), wz), p.inv);
        // StaggeredBoxSampler (tools/Interpolation.h:944-953): component c at the point + 0.5 on axis c
        const float vx = rs_box(p.vt, p.vel[0], p.velBg[0], __dadd_rn(ix, 0.5), iy, iz);
        const float vy = rs_box(p.vt, p.vel[1], p.velBg[1], ix, __dadd_rn(iy, 0.5), iz);
        const float vz = rs_box(p.vt, p.vel[2], p.velBg[2], ix, iy, __dadd_rn(iz, 0.5));
        p.o0[d] = fx_encode(jx) | (fx_encode(jy) << 16);
        p.o1[d] = fx_encode(jz) | (h_encode(vx) << 16);
        p.o2[d] = h_encode(vy) | (h_encode(vz) << 16);
        d++;
    }
}

// ---------------------------------------------------------------- ParticleEmitter (FF/nosys/ParticleEmitter.cpp:9-40 ->
// FLIP_vdb::emit_liquid, FF/FLIP_vdb.cpp:2222-2642), the branch without a velocity volume (:2488-2624).
// A leaf box takes part when one of its 9^3 lattice corners samples the shape SDF < 0 (:2270-2312). Per voxel of such a leaf: shape at
// the centre < dx -> up to 16 trials while the voxel holds < 8; a trial is skipped when its octant is taken and taken when the shape
// at the candidate is < -0.1 dx; new particles carry the constant velocity. Other leaves are not visited (no re-encode).
struct EmitParams {
    TopoView pt;
    const uint32_t* voxelStart;
    const uint32_t *w0, *w1, *w2;
    TopoView st; const float* sdf; float sdfBg;
    float dx, thr; double s, inv;
    uint32_t seed;
    const uint8_t* touched;      // [n] 1 = the shape touches this leaf
    uint32_t* tstart; uint16_t* accept; uint32_t* newCount; const uint32_t* newStart;
    uint32_t *o0, *o1, *o2;
    uint32_t velLo, velHi;       // the constant velocity as the store's half codes: w1 high half, w2
};
// one CTA per leaf box of `t`: flag = any lattice corner with shape < 0
__global__ void __launch_bounds__(256) emit_touch_kernel(TopoView t, TopoView st, const float* __restrict__ sdf, float bg, double s, double inv,
                                                         uint8_t* __restrict__ flag) {
    const int3 o = t.origin[blockIdx.x];
    bool hit = false;
    for (int q = threadIdx.x; q < 729; q += 256) {
        const int ii = q / 81, jj = (q / 9) % 9, kk = q % 9;
        const double wx = __dmul_rn((double)(o.x + ii), s), wy = __dmul_rn((double)(o.y + jj), s), wz = __dmul_rn((double)(o.z + kk), s);
        hit |= rs_box(st, sdf, bg, __dmul_rn(wx, inv), __dmul_rn(wy, inv), __dmul_rn(wz, inv)) < 0.f;
    }
    const int any = __syncthreads_or(hit ? 1 : 0);
    if (threadIdx.x == 0) flag[blockIdx.x] = any ? 1 : 0;
}
// selection for the new pool: leaves of the store that hold particles, candidate leaves the shape touches
// slab decomposition: leaves outside [xLo, xHi) (owned + one ghost layer) are not this rank's to fill
__global__ void emit_slab_kernel(const int3* __restrict__ origin, int n, int xLo, int xHi, uint8_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ((origin[i].x >> 3) < xLo || (origin[i].x >> 3) >= xHi)) flag[i] = 0;
}
__global__ void emit_select_kernel(int nP, const uint32_t* __restrict__ voxelStart, int nC, const uint8_t* __restrict__ candFlag, uint32_t* __restrict__ sel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nP) sel[i] = voxelStart[(size_t)(i + 1) * LEAF] > voxelStart[(size_t)i * LEAF] ? 1u : 0u;
    else if (i < nP + nC) sel[i] = candFlag[i - nP];
}
__global__ void emit_gather_kernel(int nP, const int3* __restrict__ po, int nC, const int3* __restrict__ co, const uint32_t* __restrict__ sel,
                                   const uint32_t* __restrict__ pos, int3* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nP + nC || !sel[i]) return;
    out[pos[i]] = i < nP ? po[i] : co[i - nP];
}
// key of every particle on the new pool (thread per old voxel)
__global__ void emit_rekey_kernel(TopoView ot, const uint32_t* __restrict__ voxelStart, TopoView nt, uint32_t* __restrict__ keys) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (size_t)ot.n * LEAF) return;
    const uint32_t b = voxelStart[v], e = voxelStart[v + 1];
    if (b == e) return;
    const int3 o = ot.origin[v >> 9];
    const uint32_t key = (uint32_t)topo_find(nt, o.x, o.y, o.z) * LEAF + (uint32_t)(v & 511);
    for (uint32_t i = b; i < e; i++) keys[i] = key;
}
__global__ void __launch_bounds__(RS_WARPS * 32) emit_decide_kernel(EmitParams p) {
    const int lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    if (leaf >= p.pt.n) return;
    const int3 o = p.pt.origin[leaf];
    const uint32_t* vs = p.voxelStart + (size_t)leaf * LEAF;
    __shared__ uint32_t sElig[RS_WARPS][16];
    uint32_t* elig = sElig[threadIdx.x >> 5];
    const bool touched = p.touched[leaf] != 0;
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
        const int off = it * 32 + lane;
        const uint32_t cnt = vs[off + 1] - vs[off];
        bool e = false;
        if (touched && cnt < 8u) {
            const double wx = __dmul_rn((double)(o.x + (off >> 6)), p.s), wy = __dmul_rn((double)(o.y + ((off >> 3) & 7)), p.s), wz = __dmul_rn((double)(o.z + (off & 7)), p.s);
            e = rs_box(p.st, p.sdf, p.sdfBg, __dmul_rn(wx, p.inv), __dmul_rn(wy, p.inv), __dmul_rn(wz, p.inv)) < p.dx;
        }
        const uint32_t eb = __ballot_sync(0xffffffffu, e);
        if (lane == 0) elig[it] = eb;
        p.newCount[(size_t)leaf * LEAF + off] = cnt;
        p.accept[(size_t)leaf * LEAF + off] = 0;
    }
    __syncwarp();
    if (!touched) return;
    unsigned int index = rs_leaf_start(p.seed, o.x, o.y, o.z);
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
        uint32_t m = elig[it];
        while (m) {
            const int off = it * 32 + (__ffs(m) - 1);
            m &= m - 1;
            const uint32_t b = vs[off], cnt = vs[off + 1] - b;
            unsigned occBit = 0;
            if ((uint32_t)lane < cnt) {   // cnt < 8
                const uint32_t a0 = __ldg(&p.w0[b + lane]), a1 = __ldg(&p.w1[b + lane]);
                const float px = fx_decode(a0 & 0xffffu), py = fx_decode(a0 >> 16), pz = fx_decode(a1 & 0xffffu);
                occBit = 1u << (((pz > 0.f) << 2) | ((py > 0.f) << 1) | (px > 0.f ? 1 : 0));
            }
            unsigned occ = __reduce_or_sync(0xffffffffu, occBit);
            const double wx = __dmul_rn((double)(o.x + (off >> 6)), p.s), wy = __dmul_rn((double)(o.y + ((off >> 3) & 7)), p.s), wz = __dmul_rn((double)(o.z + (off & 7)), p.s);
            bool pass = false;
            unsigned sv = 0;
            if (lane < 16) {
                const unsigned int at = index + 3u * (unsigned)lane;
                const float jx = rs_table(at), jy = rs_table(at + 1u), jz = rs_table(at + 2u);
                const double qx = __dadd_rn(__dmul_rn((double)jx, p.s), wx), qy = __dadd_rn(__dmul_rn((double)jy, p.s), wy), qz = __dadd_rn(__dmul_rn((double)jz, p.s), wz);
                pass = rs_box(p.st, p.sdf, p.sdfBg, __dmul_rn(qx, p.inv), __dmul_rn(qy, p.inv), __dmul_rn(qz, p.inv)) < p.thr;
                sv = ((jz > 0.f) << 2) | ((jy > 0.f) << 1) | (jx > 0.f ? 1 : 0);
            }
            const uint32_t passMask = __ballot_sync(0xffffffffu, pass);
            const uint32_t s0 = __ballot_sync(0xffffffffu, sv & 1u), s1 = __ballot_sync(0xffffffffu, sv & 2u), s2 = __ballot_sync(0xffffffffu, sv & 4u);
            uint32_t here = cnt, used = 0, acc = 0;
            for (int t = 0; t < 16 && here < 8u; t++) {
                used++;
                const unsigned oc = ((s0 >> t) & 1u) | (((s1 >> t) & 1u) << 1) | (((s2 >> t) & 1u) << 2);
                if ((occ >> oc) & 1u) continue;
                if ((passMask >> t) & 1u) { occ |= 1u << oc; acc |= 1u << t; here++; }
            }
            if (lane == 0) {
                p.tstart[(size_t)leaf * LEAF + off] = index;
                p.accept[(size_t)leaf * LEAF + off] = (uint16_t)acc;
                p.newCount[(size_t)leaf * LEAF + off] = here;
            }
            index += 3u * used;
        }
    }
}
__global__ void __launch_bounds__(256) emit_write_kernel(EmitParams p) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (size_t)p.pt.n * LEAF) return;
    const uint32_t b = p.voxelStart[v], e = p.voxelStart[v + 1];
    uint32_t d = p.newStart[v];
    const bool touched = p.touched[v >> 9] != 0;
    for (uint32_t i = b; i < e; i++, d++) {
        const uint32_t a0 = p.w0[i], a1 = p.w1[i];
        p.o0[d] = touched ? (fx_encode(fx_decode(a0 & 0xffffu)) | (fx_encode(fx_decode(a0 >> 16)) << 16)) : a0;
        p.o1[d] = touched ? (fx_encode(fx_decode(a1 & 0xffffu)) | (a1 & 0xffff0000u)) : a1;
        p.o2[d] = p.w2[i];
    }
    uint32_t acc = p.accept[v];
    const unsigned int index = acc ? p.tstart[v] : 0u;
    while (acc) {
        const int t = __ffs(acc) - 1;
        acc &= acc - 1;
        const unsigned int at = index + 3u * (unsigned)t;
        p.o0[d] = fx_encode(rs_table(at)) | (fx_encode(rs_table(at + 1u)) << 16);
        p.o1[d] = fx_encode(rs_table(at + 2u)) | p.velLo;
        p.o2[d] = p.velHi;
        d++;
    }
}

}  // namespace

void emit_liquid(World* w, int shapeGrid, float vx, float vy, float vz, uint32_t seed) {
    FB_REQUIRE(is_float_grid(shapeGrid) && w->F(shapeGrid).topo != nullptr, FLIPB200_ERR_STATE, "ParticleEmitter: the shape SDF grid was not uploaded");
    GridF& g = w->F(shapeGrid);
    // Slab decomposition: every rank is given the same shape; it emits into the leaves of its owned + ghost layers only. The
    // draws of a leaf depend on (seed, leaf origin), its own particles and the shape, so owner and ghost holder agree.
    int xLo = INT_MIN, xHi = INT_MAX;
    if (dd_on(w)) { int lo, hi; dd_owned_coords(w, &lo, &hi); xLo = lo - 1; xHi = hi + 1; }
    if (g.topo->n == 0) return;   // evalLeafBoundingBox fails on an empty shape: the reference returns (:2233-2236)
    const double s = (double)w->dx, inv = 1.0 / s;
    // 1. candidate leaf boxes = the shape's leaves and their 26 neighbours; which of them the shape touches
    TopoPtr cand = topo_from_origins_dev(w, g.topo->origin.p, g.topo->n, /*ring=*/true);
    DBuf<uint8_t> candFlag(cand->n + 1, w->stream);
    FB_LAUNCH(w, "emit_touch", (size_t)cand->n * 729 * 32) emit_touch_kernel<<<cand->n, 256, 0, w->stream>>>(cand->view(), g.topo->view(), g.val.p, g.bg, s, inv, candFlag.p);
    check_launch("emit_touch");
    if (xLo != INT_MIN || xHi != INT_MAX) { emit_slab_kernel<<<nblk(cand->n, 256), 256, 0, w->stream>>>(cand->origin.p, cand->n, xLo, xHi, candFlag.p); check_launch("emit_slab"); }
    // 2. the new pool: leaves that hold particles + touched candidates (+ ring)
    const bool have = w->pts.topo != nullptr && w->pts.topo->n > 0;
    const int nP = have ? w->pts.topo->n : 0, nC = cand->n;
    DBuf<uint32_t> sel(nP + nC + 1, w->stream), pos(nP + nC + 1, w->stream);
    FB_CUDA(cudaMemsetAsync(sel.p + nP + nC, 0, 4, w->stream));
    emit_select_kernel<<<nblk(nP + nC, 256), 256, 0, w->stream>>>(nP, have ? w->pts.voxelStart.p : nullptr, nC, candFlag.p, sel.p);
    check_launch("emit_select");
    uint64_t nSel = 0;
    exclusive_scan_u32(w, sel.p, pos.p, (size_t)nP + nC + 1, &nSel);
    if (nSel == 0) return;
    DBuf<int3> origins(nSel + 1, w->stream);
    emit_gather_kernel<<<nblk(nP + nC, 256), 256, 0, w->stream>>>(nP, have ? w->pts.topo->origin.p : nullptr, nC, cand->origin.p, sel.p, pos.p, origins.p);
    check_launch("emit_gather");
    TopoPtr pool = topo_from_origins_dev(w, origins.p, (int)nSel, /*ring=*/true);
    // 3. the store on the new pool (stable: the order inside every voxel is kept)
    Particles cur;
    const uint64_t n = have ? w->pts.n : 0;
    {
        DBuf<uint32_t> keys(n + 1, w->stream);
        if (n) {
            emit_rekey_kernel<<<nblk((size_t)nP * LEAF, 256), 256, 0, w->stream>>>(w->pts.topo->view(), w->pts.voxelStart.p, pool->view(), keys.p);
            check_launch("emit_rekey");
        }
        DBuf<uint32_t> i0 = std::move(w->pts.w0), i1 = std::move(w->pts.w1), i2 = std::move(w->pts.w2);
        sort_store_by_key(w, pool, keys.p, n, i0.p, i1.p, i2.p, cur);
    }
    // 4. which pool leaves the shape touches, then decide / scan / write as in FluidReseed
    const int nl = pool->n;
    const size_t nv = (size_t)nl * LEAF;
    DBuf<uint8_t> touched(nl + 1, w->stream);
    FB_LAUNCH(w, "emit_touch", (size_t)nl * 729 * 32) emit_touch_kernel<<<nl, 256, 0, w->stream>>>(pool->view(), g.topo->view(), g.val.p, g.bg, s, inv, touched.p);
    check_launch("emit_touch");
    if (xLo != INT_MIN || xHi != INT_MAX) { emit_slab_kernel<<<nblk(nl, 256), 256, 0, w->stream>>>(pool->origin.p, nl, xLo, xHi, touched.p); check_launch("emit_slab"); }
    DBuf<uint32_t> tstart(nv, w->stream), newCount(nv + 1, w->stream), newStart(nv + 1, w->stream);
    DBuf<uint16_t> accept(nv, w->stream);
    FB_CUDA(cudaMemsetAsync(newCount.p + nv, 0, 4, w->stream));
    EmitParams p;
    p.pt = pool->view();
    p.voxelStart = cur.voxelStart.p; p.w0 = cur.w0.p; p.w1 = cur.w1.p; p.w2 = cur.w2.p;
    p.st = g.topo->view(); p.sdf = g.val.p; p.sdfBg = g.bg;
    p.dx = w->dx; p.thr = (float)((double)(-w->dx) * 0.1); p.s = s; p.inv = inv;
    p.seed = seed; p.touched = touched.p;
    p.tstart = tstart.p; p.accept = accept.p; p.newCount = newCount.p; p.newStart = nullptr;
    p.o0 = p.o1 = p.o2 = nullptr;
    p.velLo = (uint32_t)__half_as_ushort(__float2half_rn(vx)) << 16;
    p.velHi = (uint32_t)__half_as_ushort(__float2half_rn(vy)) | ((uint32_t)__half_as_ushort(__float2half_rn(vz)) << 16);
    FB_LAUNCH(w, "emit_decide", nv * 16) emit_decide_kernel<<<(nl + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, 0, w->stream>>>(p);
    check_launch("emit_decide");
    uint64_t total = 0;
    exclusive_scan_u32(w, newCount.p, newStart.p, nv + 1, &total);
    Particles out;
    out.topo = pool;
    out.n = total;
    out.w0.alloc(total + 1, w->stream); out.w1.alloc(total + 1, w->stream); out.w2.alloc(total + 1, w->stream);
    p.newStart = newStart.p;
    p.o0 = out.w0.p; p.o1 = out.w1.p; p.o2 = out.w2.p;
    FB_LAUNCH(w, "emit_write", (n + total) * 12 + nv * 14) emit_write_kernel<<<nblk(nv, 256), 256, 0, w->stream>>>(p);
    check_launch("emit_write");
    out.voxelStart = std::move(newStart);
    w->pts = std::move(out);
    w->pool = pool;
}

// ---------------------------------------------------------------- FLIPApplyBoundary (FF/nosys/Update_Solid_SDF.cpp:9-31 ->
// FLIP_vdb::update_solid_sdf, FF/FLIP_vdb.cpp:1976-2046), one moving solid. Leaves of the new static SDF = its old leaves + the
// leaf under every moving-solid leaf origin + every particle leaf with a box corner inside the moving solid; then every voxel of
// every leaf = min(old value or background, moving solid sampled at the voxel's world position), all active.
// Transforms: particles x = i s; static SDF x = i s + t, t = -0.5 s (vertex centred); the moving solid on either.
namespace {
struct BoundaryParams {
    int nS, nM, nP;
    const int3 *so, *mo, *po;
    const uint8_t *sAlloc, *mAlloc;
    const uint32_t* voxelStart;
    TopoView mt; const float* mval; float mbg;
    double s, inv, t, mtr;       // mtr = the moving solid's translation (0 or t)
};
__device__ __forceinline__ float bd_sample(const BoundaryParams& p, double wx, double wy, double wz) {
    return rs_box(p.mt, p.mval, p.mbg, __dmul_rn(__dsub_rn(wx, p.mtr), p.inv), __dmul_rn(__dsub_rn(wy, p.mtr), p.inv), __dmul_rn(__dsub_rn(wz, p.mtr), p.inv));
}
__global__ void boundary_select_kernel(BoundaryParams p, uint32_t* __restrict__ sel, int3* __restrict__ org) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nS + p.nM + p.nP) return;
    int3 o = make_int3(0, 0, 0);
    uint32_t on = 0;
    if (i < p.nS) { o = p.so[i]; on = p.sAlloc[i] ? 1u : 0u; }
    else if (i < p.nS + p.nM) {
        const int3 m = p.mo[i - p.nS];
        on = p.mAlloc[i - p.nS] ? 1u : 0u;
        const double wx = __dadd_rn(__dmul_rn((double)m.x, p.s), p.mtr), wy = __dadd_rn(__dmul_rn((double)m.y, p.s), p.mtr), wz = __dadd_rn(__dmul_rn((double)m.z, p.s), p.mtr);
        o.x = (int)floor(__dmul_rn(__dsub_rn(wx, p.t), p.inv)) & ~7; o.y = (int)floor(__dmul_rn(__dsub_rn(wy, p.t), p.inv)) & ~7; o.z = (int)floor(__dmul_rn(__dsub_rn(wz, p.t), p.inv)) & ~7;
    } else {
        const int l = i - p.nS - p.nM;
        o = p.po[l];
        if (p.voxelStart[(size_t)(l + 1) * LEAF] > p.voxelStart[(size_t)l * LEAF]) {
            for (int c = 0; c < 8 && !on; c++)
                on = bd_sample(p, __dmul_rn((double)(o.x + ((c >> 2) & 1) * 8), p.s), __dmul_rn((double)(o.y + ((c >> 1) & 1) * 8), p.s), __dmul_rn((double)(o.z + (c & 1) * 8), p.s)) < 0.f ? 1u : 0u;
        }
    }
    sel[i] = on; org[i] = o;
}
__global__ void boundary_gather_kernel(int n, const int3* __restrict__ org, const uint32_t* __restrict__ sel, const uint32_t* __restrict__ pos, int3* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && sel[i]) out[pos[i]] = org[i];
}
__global__ void __launch_bounds__(512) boundary_min_kernel(TopoView nt, TopoView ot, const float* __restrict__ oval, const uint8_t* __restrict__ oalloc, float bg,
                                                           BoundaryParams p, float* __restrict__ val, uint64_t* __restrict__ mask, uint8_t* __restrict__ alloc) {
    const int leaf = blockIdx.x, off = threadIdx.x;
    const int3 o = nt.origin[leaf];
    float cur = bg;
    if (ot.n > 0) {
        const int l = topo_find(ot, o.x, o.y, o.z);
        if (l >= 0 && oalloc[l]) cur = oval[(size_t)l * LEAF + off];
    }
    const double wx = __dadd_rn(__dmul_rn((double)(o.x + (off >> 6)), p.s), p.t), wy = __dadd_rn(__dmul_rn((double)(o.y + ((off >> 3) & 7)), p.s), p.t),
                 wz = __dadd_rn(__dmul_rn((double)(o.z + (off & 7)), p.s), p.t);
    val[(size_t)leaf * LEAF + off] = fminf(cur, bd_sample(p, wx, wy, wz));
    if (off < 8) mask[(size_t)leaf * 8 + off] = ~0ull;
    if (off == 0) alloc[leaf] = 1;
}
}  // namespace

void apply_boundary(World* w, int movingGrid, bool movingVertexCentred) {
    FB_REQUIRE(is_float_grid(movingGrid) && movingGrid != FLIPB200_SOLID_SDF && w->F(movingGrid).topo != nullptr, FLIPB200_ERR_STATE,
               "FLIPApplyBoundary: the moving solid SDF grid was not uploaded");
    // (slab decomposition: every rank is given the same moving solid; its static SDF gains the leaves its own owned + ghost
    // particle leaves call for -- everything a rank's particles can sample. Nothing is exchanged.)
    GridF& mv = w->F(movingGrid);
    GridF& sd = w->F(FLIPB200_SOLID_SDF);
    const bool haveS = sd.topo != nullptr && sd.topo->n > 0, haveP = w->pts.topo != nullptr && w->pts.topo->n > 0;
    BoundaryParams p;
    p.nS = haveS ? sd.topo->n : 0; p.nM = mv.topo->n; p.nP = haveP ? w->pts.topo->n : 0;
    p.so = haveS ? sd.topo->origin.p : nullptr; p.mo = mv.topo->origin.p; p.po = haveP ? w->pts.topo->origin.p : nullptr;
    p.sAlloc = haveS ? sd.alloc.p : nullptr; p.mAlloc = mv.alloc.p;
    p.voxelStart = haveP ? w->pts.voxelStart.p : nullptr;
    p.mt = mv.topo->view(); p.mval = mv.val.p; p.mbg = mv.bg;
    p.s = (double)w->dx; p.inv = 1.0 / p.s; p.t = -0.5 * p.s; p.mtr = movingVertexCentred ? p.t : 0.0;
    const int n = p.nS + p.nM + p.nP;
    if (n == 0) return;
    DBuf<uint32_t> sel(n + 1, w->stream), pos(n + 1, w->stream);
    DBuf<int3> org(n + 1, w->stream);
    FB_CUDA(cudaMemsetAsync(sel.p + n, 0, 4, w->stream));
    FB_LAUNCH(w, "boundary_select", (size_t)n * 300) boundary_select_kernel<<<nblk(n, 128), 128, 0, w->stream>>>(p, sel.p, org.p);
    check_launch("boundary_select");
    uint64_t cnt = 0;
    exclusive_scan_u32(w, sel.p, pos.p, (size_t)n + 1, &cnt);
    if (cnt == 0) return;
    DBuf<int3> origins(cnt + 1, w->stream);
    boundary_gather_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(n, org.p, sel.p, pos.p, origins.p);
    check_launch("boundary_gather");
    TopoPtr nt = topo_from_origins_dev(w, origins.p, (int)cnt, /*ring=*/false);
    GridF g;
    grid_alloc(w, g, nt, sd.bg);
    const TopoView ot = haveS ? sd.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
    FB_LAUNCH(w, "boundary_min", (size_t)nt->n * LEAF * 40) boundary_min_kernel<<<nt->n, 512, 0, w->stream>>>(nt->view(), ot, haveS ? sd.val.p : nullptr, haveS ? sd.alloc.p : nullptr,
                                                                                                       sd.bg, p, g.val.p, g.mask.p, g.alloc.p);
    check_launch("boundary_min");
    w->F(FLIPB200_SOLID_SDF) = std::move(g);
    w->hasSolidSDF = true;
    w->solidViewEpoch = ~0ull;
}

void fluid_reseed(World* w, uint32_t seed) {
    FB_REQUIRE(w->pts.topo != nullptr, FLIPB200_ERR_STATE, "FluidReseed: no particles");
    GridF& sdf = w->F(FLIPB200_LIQUID_SDF);
    GridV& vel = w->V(FLIPB200_VELOCITY);
    FB_REQUIRE(sdf.topo != nullptr && vel.topo != nullptr, FLIPB200_ERR_STATE, "FluidReseed: LiquidSDF / Velocity are not set (run FLIP_P2G or upload them)");
    const TopoPtr topo = w->pts.topo;
    const int n = topo->n;
    if (n == 0) return;
    const size_t nv = (size_t)n * LEAF;
    DBuf<uint32_t> tstart(nv, w->stream), newCount(nv + 1, w->stream), newStart(nv + 1, w->stream);
    DBuf<uint16_t> accept(nv, w->stream);
    FB_CUDA(cudaMemsetAsync(newCount.p + nv, 0, 4, w->stream));
    ReseedParams p;
    p.pt = topo->view();
    p.voxelStart = w->pts.voxelStart.p; p.w0 = w->pts.w0.p; p.w1 = w->pts.w1.p; p.w2 = w->pts.w2.p;
    p.st = sdf.topo->view(); p.sdf = sdf.val.p; p.sdfBg = sdf.bg;
    p.vt = vel.topo->view();
    for (int c = 0; c < 3; c++) { p.vel[c] = vel.val[c].p; p.velBg[c] = vel.bg[c]; }
    p.dx = w->dx; p.s = (double)w->dx; p.inv = 1.0 / (double)w->dx;
    p.seed = seed;
    // Slab decomposition: a leaf's draws start at a hash of (seed, leaf origin) and read the leaf's own particles plus the
    // liquid SDF / velocity within two voxels of it, so the owner of a leaf and the neighbour that holds it as a ghost make the
    // same decisions without an exchange -- as long as the grids' ghost layers are current (FLIP_P2G refreshes two). Leaves
    // beyond the ghost layer (the pool's ring) hold none of their particles here and are left alone.
    p.xLo = INT_MIN; p.xHi = INT_MAX;
    if (dd_on(w)) { int lo, hi; dd_owned_coords(w, &lo, &hi); p.xLo = lo - 1; p.xHi = hi + 1; }
    p.tstart = tstart.p; p.accept = accept.p; p.newCount = newCount.p; p.newStart = nullptr;
    p.o0 = p.o1 = p.o2 = nullptr;
    FB_LAUNCH(w, "reseed_decide", nv * 16) reseed_decide_kernel<<<(n + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, 0, w->stream>>>(p);
    check_launch("reseed_decide");
    uint64_t total = 0;
    exclusive_scan_u32(w, newCount.p, newStart.p, nv + 1, &total);
    Particles out;
    out.topo = topo;
    out.n = total;
    out.w0.alloc(total + 1, w->stream); out.w1.alloc(total + 1, w->stream); out.w2.alloc(total + 1, w->stream);
    p.newStart = newStart.p;
    p.o0 = out.w0.p; p.o1 = out.w1.p; p.o2 = out.w2.p;
    FB_LAUNCH(w, "reseed_write", (w->pts.n + total) * 12 + nv * 14) reseed_write_kernel<<<nblk(nv, 256), 256, 0, w->stream>>>(p);
    check_launch("reseed_write");
    out.voxelStart = std::move(newStart);
    w->pts = std::move(out);
}

}  // namespace fb
