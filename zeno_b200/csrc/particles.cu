// libflipb200 -- particle store: initial binning (K1) and the stable counting sort by
// (leaf slot, voxel offset) used by K1 and by the re-binning after advection (K2).
//
// The sort is bit-exact and deterministic: it produces exactly the order of a stable sort by
// key of the source order (what openvdb's PointPartitioner and the reference's per-leaf
// counting sort, FF/FLIP_vdb.cpp:3406-3476, produce for a sequential traversal).
//   histogram (integer atomics)  ->  exclusive scan  ->  scatter source indices  ->
//   per-voxel ascending fix-up of the (few) indices  ->  coalesced payload gather.
#include "world.cuh"
#include "next_kernels.cuh"

namespace fb {
namespace {

constexpr uint32_t KEY_DROPPED = 0xffffffffu;
constexpr uint32_t VOXEL_CAP = 28;  // "existing_par > 27 -> drop" (FF/FLIP_vdb.cpp:711-714,742-745)

// K1: particleArrayToGrid (projects/zenvdb/SetVDBPointDataGrid.cpp:17-72)
//  index = double(pos) * (1.0/double(dx))       (openvdb/math/Maps.h:688,751-753)
//  ijk   = floor(index + 0.5)                   (math/Coord.h:50-53, math/Math.h:822-823)
//  P     = fxpt16(float(index - ijk))           (points/PointConversion.h:700-718)
//  v     = half(vel)                            (TruncateCodec)
__global__ void bin_encode_kernel(const float* __restrict__ pos, const float* __restrict__ vel, uint64_t n,
                                  double inv, int3* __restrict__ ijkOut, uint32_t* __restrict__ w0,
                                  uint32_t* __restrict__ w1, uint32_t* __restrict__ w2) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3];
    uint32_t P[3], V[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double idx = __dmul_rn((double)pos[3 * i + a], inv);
        c[a] = (int)floor(__dadd_rn(idx, 0.5));
        float local = (float)__dsub_rn(idx, (double)c[a]);
        P[a] = fx_encode(local);
        V[a] = h_encode(vel ? vel[3 * i + a] : 0.f);
    }
    ijkOut[i] = make_int3(c[0], c[1], c[2]);
    w0[i] = P[0] | (P[1] << 16);
    w1[i] = P[2] | (V[0] << 16);
    w2[i] = V[1] | (V[2] << 16);
}
__global__ void ijk_to_origin_kernel(const int3* __restrict__ ijk, uint64_t n, int3* __restrict__ origins) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int3 c = ijk[i];
    origins[i] = make_int3(c.x & ~7, c.y & ~7, c.z & ~7);
}
__global__ void ijk_to_key_kernel(TopoView t, const int3* __restrict__ ijk, const uint8_t* __restrict__ alive,
                                  uint64_t n, uint32_t* __restrict__ keys) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (alive && !alive[i]) { keys[i] = KEY_DROPPED; return; }
    int3 c = ijk[i];
    int l = topo_find(t, c.x, c.y, c.z);
    keys[i] = l < 0 ? KEY_DROPPED : (uint32_t)l * LEAF + (uint32_t)voxel_off(c.x, c.y, c.z);
}
__global__ void hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = keys[i];
    if (k != KEY_DROPPED) atomicAdd(&count[k], 1u);
}
__global__ void cap_kernel(const uint32_t* __restrict__ count, uint32_t* __restrict__ capped, size_t nv, uint32_t cap) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) capped[i] = min(count[i], cap);
}
__global__ void scatter_kernel(const uint32_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ start,
                               uint32_t* __restrict__ fill, uint32_t* __restrict__ perm) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = keys[i];
    if (k == KEY_DROPPED) return;
    uint32_t r = atomicAdd(&fill[k], 1u);
    perm[start[k] + r] = (uint32_t)i;
}
// per voxel: ascending order of the source indices (= stable), then apply the per-voxel cap
__global__ void fixup_kernel(const uint32_t* __restrict__ startU, const uint32_t* __restrict__ startC,
                             size_t nv, uint32_t* perm, uint32_t* __restrict__ perm2) {
    size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint32_t b = startU[v], e = startU[v + 1];
    const uint32_t cb = startC[v], keep = startC[v + 1] - cb;
    if (e - b <= 64u) {
        // rank sort: an entry goes to (number of smaller entries); the source indices are distinct. No store feeds a later load,
        // unlike the in-place insertion sort this replaces (214 us: every step waited for its own write to come back).
        for (uint32_t i = b; i < e; i++) {
            const uint32_t x = __ldg(&perm[i]);
            uint32_t r = 0;
            for (uint32_t j = b; j < e; j++) r += __ldg(&perm[j]) < x ? 1u : 0u;
            if (r < keep) perm2[cb + r] = x;
        }
        return;
    }
    for (uint32_t i = b + 1; i < e; i++) {
        uint32_t x = perm[i];
        uint32_t j = i;
        while (j > b && perm[j - 1] > x) { perm[j] = perm[j - 1]; j--; }
        perm[j] = x;
    }
    for (uint32_t j = 0; j < keep; j++) perm2[cb + j] = perm[b + j];
}
__global__ void gather_kernel(const uint32_t* __restrict__ perm, uint64_t m, const uint32_t* __restrict__ i0,
                              const uint32_t* __restrict__ i1, const uint32_t* __restrict__ i2,
                              uint32_t* __restrict__ o0, uint32_t* __restrict__ o1, uint32_t* __restrict__ o2) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s = perm[i];
    o0[i] = __ldg(&i0[s]);
    o1[i] = __ldg(&i1[s]);
    o2[i] = __ldg(&i2[s]);
}
inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

// keys -> sorted store on topo. cap = 0 disables the per-voxel cap.
void sort_by_key(World* w, const TopoPtr& topo, const uint32_t* keys, uint64_t n, uint32_t cap,
                 const uint32_t* i0, const uint32_t* i1, const uint32_t* i2, Particles& out, uint64_t* keptOut) {
    size_t nv = (size_t)topo->n * LEAF;
    DBuf<uint32_t> count(nv + 1, w->stream), startU(nv + 1, w->stream), fillc(nv + 1, w->stream);
    count.zero();
    fillc.zero();
    if (n) {
        FB_LAUNCH(w, "rebin_hist", n * 8) hist_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(keys, n, count.p);
        check_launch("hist");
    }
    uint64_t totalU = 0, totalC = 0;
    exclusive_scan_u32(w, count.p, startU.p, nv + 1, cap ? nullptr : &totalU);
    DBuf<uint32_t> startC;
    if (cap) {
        startC.alloc(nv + 1, w->stream);
        FB_LAUNCH(w, "rebin_cap", nv * 8) cap_kernel<<<nblk(nv + 1, 256), 256, 0, w->stream>>>(count.p, startC.p, nv + 1, cap);
        check_launch("cap");
        exclusive_scan_u32(w, startC.p, startC.p, nv + 1, &totalC);
    } else totalC = totalU;
    const uint32_t* sC = cap ? startC.p : startU.p;
    DBuf<uint32_t> perm(n ? n : 1, w->stream), perm2(totalC ? totalC : 1, w->stream);
    if (n) {
        FB_LAUNCH(w, "rebin_scatter", n * 12) scatter_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(keys, n, startU.p, fillc.p, perm.p);
        check_launch("scatter");
        FB_LAUNCH(w, "rebin_fixup", nv * 8 + n * 12) fixup_kernel<<<nblk(nv, 256), 256, 0, w->stream>>>(startU.p, sC, nv, perm.p, perm2.p);
        check_launch("fixup");
    }
    out.topo = topo;
    out.n = totalC;
    DBuf<uint32_t> o0(totalC ? totalC : 1, w->stream), o1(totalC ? totalC : 1, w->stream), o2(totalC ? totalC : 1, w->stream);
    if (totalC) {
        FB_LAUNCH(w, "rebin_gather", totalC * 28) gather_kernel<<<nblk(totalC, 256), 256, 0, w->stream>>>(perm2.p, totalC, i0, i1, i2, o0.p, o1.p, o2.p);
        check_launch("gather");
    }
    out.w0 = std::move(o0); out.w1 = std::move(o1); out.w2 = std::move(o2);
    if (cap) out.voxelStart = std::move(startC);
    else out.voxelStart = std::move(startU);
    if (keptOut) *keptOut = totalC;
}
}  // namespace

// the stable counting sort without the per-voxel cap, for the other translation units (reseed.cu)
void sort_store_by_key(World* w, const TopoPtr& topo, const uint32_t* keys, uint64_t n, const uint32_t* i0, const uint32_t* i1, const uint32_t* i2,
                       Particles& out) {
    uint64_t kept = 0;
    sort_by_key(w, topo, keys, n, /*cap=*/0, i0, i1, i2, out, &kept);
}

void origins_from_ijk(World* w, const int3* ijk, uint64_t n, int3* origins);

void bin_from_points(World* w, const float* pos_host, const float* vel_host, uint64_t n) {
    reserve_pool(w, ((uint64_t)1 << 30) + 512 * n);
    DBuf<float> pos(3 * n + 1, w->stream), vel;
    FB_CUDA(cudaMemcpyAsync(pos.p, pos_host, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, w->stream));
    if (vel_host) {
        vel.alloc(3 * n + 1, w->stream);
        FB_CUDA(cudaMemcpyAsync(vel.p, vel_host, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, w->stream));
    }
    DBuf<int3> ijk(n + 1, w->stream), origins(n + 1, w->stream);
    DBuf<uint32_t> w0(n + 1, w->stream), w1(n + 1, w->stream), w2(n + 1, w->stream), keys(n + 1, w->stream);
    const double inv = 1.0 / (double)w->dx;
    if (n) {
        FB_LAUNCH(w, "bin_encode", n * 48) bin_encode_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(pos.p, vel.p, n, inv, ijk.p, w0.p, w1.p, w2.p);
        check_launch("bin_encode");
        FB_LAUNCH(w, "bin_origins", n * 24) ijk_to_origin_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(ijk.p, n, origins.p);
        check_launch("ijk_to_origin");
    }
    if (dd_on(w)) {
        // slab decomposition: the caller hands each rank (at least) its own points; whatever lies in a neighbour's
        // slab moves there and the ghost layers are filled before the store is sorted
        DBuf<uint32_t> m0, m1, m2;
        DBuf<int3> mijk;
        uint64_t nm = 0;
        dd_migrate(w, 0, n, w0.p, w1.p, w2.p, ijk.p, nullptr, m0, m1, m2, mijk, &nm);
        n = nm;
        w0 = std::move(m0); w1 = std::move(m1); w2 = std::move(m2); ijk = std::move(mijk);
        origins.alloc(n + 1, w->stream);
        keys.alloc(n + 1, w->stream);
        origins_from_ijk(w, ijk.p, n, origins.p);
    }
    TopoPtr pool = topo_from_origins_dev(w, origins.p, (int)n, /*ring=*/true);
    if (n) {
        FB_LAUNCH(w, "bin_keys", n * 16) ijk_to_key_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(pool->view(), ijk.p, nullptr, n, keys.p);
        check_launch("ijk_to_key");
    }
    Particles out;
    sort_by_key(w, pool, keys.p, n, /*cap=*/0, w0.p, w1.p, w2.p, out, nullptr);
    w->pts = std::move(out);
    w->pool = pool;
    w->dropped = 0;
}

void rebin_particles(World* w, const TopoPtr& newPool, const uint32_t* keys_dev, uint64_t nOld,
                     DBuf<uint32_t>& w0, DBuf<uint32_t>& w1, DBuf<uint32_t>& w2) {
    Particles out;
    uint64_t kept = 0;
    sort_by_key(w, newPool, keys_dev, nOld, VOXEL_CAP, w0.p, w1.p, w2.p, out, &kept);
    w->dropped = nOld - kept;
    w->pts = std::move(out);
}

// ---------------------------------------------------------------- nodes beyond the substep chain (bodies: next_kernels.cuh)
namespace {
// KillParticlesInSDF: one CTA per store leaf
__global__ void __launch_bounds__(256) kill_keys_kernel(TopoView pt, const uint32_t* __restrict__ voxelStart, uint32_t* __restrict__ w0,
                                                        uint32_t* __restrict__ w1, TopoView st, const float* __restrict__ sval, float sbg,
                                                        int keep, uint32_t* __restrict__ keys) {
    __shared__ uint32_t sStart[LEAF + 1];
    const int leaf = blockIdx.x;
    const size_t vbase = (size_t)leaf * LEAF;
    for (int i = threadIdx.x; i <= LEAF; i += blockDim.x) sStart[i] = __ldg(&voxelStart[vbase + i]);
    __syncthreads();
    const uint32_t beg = sStart[0], end = sStart[LEAF];
    for (uint32_t gi = beg + threadIdx.x; gi < end; gi += blockDim.x)
        nextk::kill_keys_one(pt, sStart, w0, w1, st, sval, sbg, keep, keys, leaf, gi);
}
// ParticleAddDV
__global__ void add_dv_kernel(uint32_t* __restrict__ w1, uint32_t* __restrict__ w2, uint64_t n, double dx, double dy, double dz) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nextk::add_dv_one(w1, w2, i, dx, dy, dz);
}
}  // namespace

void particles_add_dv(World* w, float dvx, float dvy, float dvz) {
    FB_REQUIRE(w->pts.topo != nullptr, FLIPB200_ERR_STATE, "ParticleAddDV: no particles");
    const uint64_t n = w->pts.n;
    if (!n) return;
    FB_LAUNCH(w, "particles_add_dv", n * 16) add_dv_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(w->pts.w1.p, w->pts.w2.p, n, (double)dvx, (double)dvy, (double)dvz);
    check_launch("add_dv");
}

void kill_particles_in_sdf(World* w, int sdfGrid, bool keep) {
    FB_REQUIRE(w->pts.topo != nullptr, FLIPB200_ERR_STATE, "KillParticlesInSDF: no particles");
    FB_REQUIRE(is_float_grid(sdfGrid) && w->F(sdfGrid).topo != nullptr, FLIPB200_ERR_STATE, "KillParticlesInSDF: the killer SDF grid was not uploaded");
    GridF& g = w->F(sdfGrid);
    const TopoPtr topo = w->pts.topo;
    const uint64_t n = w->pts.n;
    if (topo->n == 0 || n == 0) return;
    DBuf<uint32_t> keys(n + 1, w->stream);
    FB_LAUNCH(w, "kill_keys", n * 16) kill_keys_kernel<<<topo->n, 256, 0, w->stream>>>(topo->view(), w->pts.voxelStart.p, w->pts.w0.p, w->pts.w1.p,
                                                                                     g.topo->view(), g.val.p, g.bg, keep ? 1 : 0, keys.p);
    check_launch("kill_keys");
    DBuf<uint32_t> i0 = std::move(w->pts.w0), i1 = std::move(w->pts.w1), i2 = std::move(w->pts.w2);
    Particles out;
    uint64_t kept = 0;
    sort_by_key(w, topo, keys.p, n, /*cap=*/0, i0.p, i1.p, i2.p, out, &kept);
    w->pts = std::move(out);
}

// VDBPointsToPrimitive (projects/zenvdb/GetVDBPoints.cpp:76-258): world position = float((double(P) + double(voxel)) * dx) and the decoded
// velocity of every particle, store order; one thread per voxel. Written into device staging, copied to the caller's host arrays.
namespace {
__global__ void points_export_kernel(TopoView t, const uint32_t* __restrict__ voxelStart, const uint32_t* __restrict__ w0, const uint32_t* __restrict__ w1,
                                     const uint32_t* __restrict__ w2, double s, float* __restrict__ pos, float* __restrict__ vel) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (size_t)t.n * LEAF) return;
    const uint32_t b = voxelStart[v], e = voxelStart[v + 1];
    if (b == e) return;
    const int3 o = t.origin[v >> 9];
    const int off = (int)(v & 511);
    const double cx = (double)(o.x + (off >> 6)), cy = (double)(o.y + ((off >> 3) & 7)), cz = (double)(o.z + (off & 7));
    for (uint32_t i = b; i < e; i++) {
        const uint32_t a0 = w0[i], a1 = w1[i], a2 = w2[i];
        pos[3 * (size_t)i] = __double2float_rn(__dmul_rn(__dadd_rn((double)fx_decode(a0 & 0xffffu), cx), s));
        pos[3 * (size_t)i + 1] = __double2float_rn(__dmul_rn(__dadd_rn((double)fx_decode(a0 >> 16), cy), s));
        pos[3 * (size_t)i + 2] = __double2float_rn(__dmul_rn(__dadd_rn((double)fx_decode(a1 & 0xffffu), cz), s));
        if (vel) { vel[3 * (size_t)i] = h_decode(a1 >> 16); vel[3 * (size_t)i + 1] = h_decode(a2 & 0xffffu); vel[3 * (size_t)i + 2] = h_decode(a2 >> 16); }
    }
}
}  // namespace
void particles_to_points(World* w, float* posHost, float* velHost) {
    FB_REQUIRE(w->pts.topo != nullptr, FLIPB200_ERR_STATE, "VDBPointsToPrimitive: no particles");
    const uint64_t n = w->pts.n;
    if (n == 0) return;
    DBuf<float> pos(3 * n, w->stream), vel;
    if (velHost) vel.alloc(3 * n, w->stream);
    const size_t nv = (size_t)w->pts.topo->n * LEAF;
    FB_LAUNCH(w, "points_export", n * (12 + 24)) points_export_kernel<<<nblk(nv, 256), 256, 0, w->stream>>>(w->pts.topo->view(), w->pts.voxelStart.p, w->pts.w0.p, w->pts.w1.p, w->pts.w2.p,
                                                                                                       (double)w->dx, pos.p, velHost ? vel.p : nullptr);
    check_launch("points_export");
    FB_CUDA(cudaMemcpyAsync(posHost, pos.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, w->stream));
    if (velHost) FB_CUDA(cudaMemcpyAsync(velHost, vel.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, w->stream));
    sync(w);
}

// helpers shared with g2p.cu
void origins_from_ijk(World* w, const int3* ijk, uint64_t n, int3* origins) {
    if (!n) return;
    FB_LAUNCH(w, "bin_origins", n * 24) ijk_to_origin_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(ijk, n, origins);
    check_launch("ijk_to_origin");
}
void keys_from_ijk(World* w, const TopoPtr& t, const int3* ijk, const uint8_t* alive, uint64_t n, uint32_t* keys) {
    if (!n) return;
    FB_LAUNCH(w, "bin_keys", n * 17) ijk_to_key_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(t->view(), ijk, alive, n, keys);
    check_launch("ijk_to_key");
}

}  // namespace fb
