// libflipb200 -- leaf topology (dense directory), grid storage, pool management, mask morphology.
// Replaces what the reference gets from openvdb::tree::Tree4 + tools::dilateActiveValues +
// TopologyCopy/topologyUnion on this path.
#include "world.cuh"
#include <algorithm>
#include <climits>

namespace fb {
namespace {

// bounding box of the candidate leaves: grid-stride over a fixed number of CTAs, block-level reduction,
// then six integer atomics per CTA (the first version issued them per warp over 16.8 M particles and
// spent 2 ms per substep serialising on six addresses)
constexpr int BBOX_THREADS = 256;
__global__ void __launch_bounds__(BBOX_THREADS) bbox_kernel(const int3* __restrict__ origins, int count, int* __restrict__ bb) {
    __shared__ int sLo[3][BBOX_THREADS / 32], sHi[3][BBOX_THREADS / 32];
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        int3 o = origins[i];
        int c[3] = {o.x >> 3, o.y >> 3, o.z >> 3};
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = min(lo[a], c[a]); hi[a] = max(hi[a], c[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        for (int d = 16; d > 0; d >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
        }
        if ((threadIdx.x & 31) == 0) { sLo[a][threadIdx.x >> 5] = lo[a]; sHi[a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int a = threadIdx.x, l = INT_MAX, h = INT_MIN;
        for (int k = 0; k < BBOX_THREADS / 32; k++) { l = min(l, sLo[a][k]); h = max(h, sHi[a][k]); }
        if (l != INT_MAX) { atomicMin(&bb[a], l); atomicMax(&bb[3 + a], h); }
    }
}
// flags the directory cell of every candidate leaf; consecutive candidates usually share a leaf (the
// particle store is leaf-sorted), so a lane only writes when its cell differs from the previous lane's
__global__ void mark_kernel(const int3* __restrict__ origins, int count, int3 dmin, int3 ddim,
                            uint32_t* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cell = -1;
    if (i < count) {
        int3 o = origins[i];
        cell = (((o.x >> 3) - dmin.x) * ddim.y + ((o.y >> 3) - dmin.y)) * ddim.z + ((o.z >> 3) - dmin.z);
    }
    int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    if (cell >= 0 && ((threadIdx.x & 31) == 0 || prev != cell)) flags[cell] = 1u;
}
// ring = 1: also flag the 26 neighbour cells of every flagged cell (gather over the directory volume)
__global__ void ring_kernel(const uint32_t* __restrict__ in, int3 ddim, uint32_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = ddim.x * ddim.y * ddim.z;
    if (i >= total) return;
    int lz = i % ddim.z, ly = (i / ddim.z) % ddim.y, lx = i / (ddim.z * ddim.y);
    uint32_t on = 0;
    for (int a = -1; a <= 1; a++)
        for (int b = -1; b <= 1; b++)
            for (int c = -1; c <= 1; c++) {
                int x = lx + a, y = ly + b, z = lz + c;
                if ((unsigned)x < (unsigned)ddim.x && (unsigned)y < (unsigned)ddim.y && (unsigned)z < (unsigned)ddim.z)
                    on |= in[(x * ddim.y + y) * ddim.z + z];
            }
    out[i] = on ? 1u : 0u;
}
__global__ void assign_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ slots,
                              int3 dmin, int3 ddim, int* __restrict__ dir, int3* __restrict__ origin) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = ddim.x * ddim.y * ddim.z;
    if (i >= total) return;
    if (flags[i]) {
        int s = (int)slots[i];
        dir[i] = s;
        int lz = i % ddim.z, ly = (i / ddim.z) % ddim.y, lx = i / (ddim.z * ddim.y);
        origin[s] = make_int3((lx + dmin.x) * 8, (ly + dmin.y) * 8, (lz + dmin.z) * 8);
    } else dir[i] = -1;
}
__global__ void nbr27_kernel(TopoView t, int* __restrict__ nbr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n * 27) return;
    int l = i / 27, k = i % 27;
    int3 o = t.origin[l];
    int a = k / 9 - 1, b = (k / 3) % 3 - 1, c = k % 3 - 1;
    nbr[i] = topo_find(t, o.x + 8 * a, o.y + 8 * b, o.z + 8 * c);
}
__global__ void fill_kernel(float* __restrict__ p, float v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// copy leaves of an old grid into the storage of a new topology
__global__ void rebase_kernel(TopoView oldT, TopoView newT, const float* __restrict__ oldVal,
                              float* __restrict__ newVal, const uint64_t* __restrict__ oldMask,
                              uint64_t* __restrict__ newMask, const uint8_t* __restrict__ oldAlloc,
                              uint8_t* __restrict__ newAlloc) {
    int l = blockIdx.x;
    int3 o = oldT.origin[l];
    int nl = topo_find(newT, o.x, o.y, o.z);
    if (nl < 0) return;
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x) newVal[(size_t)nl * LEAF + i] = oldVal[(size_t)l * LEAF + i];
    if (oldMask && threadIdx.x < 8) newMask[(size_t)nl * 8 + threadIdx.x] = oldMask[(size_t)l * 8 + threadIdx.x];
    if (oldAlloc && threadIdx.x == 0) newAlloc[nl] = oldAlloc[l];
}
__global__ void nonempty_origins_kernel(TopoView t, const uint64_t* __restrict__ mask, const uint8_t* __restrict__ alloc,
                                        int3* __restrict__ out, uint32_t* __restrict__ counter) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= t.n) return;
    uint64_t any = alloc ? alloc[l] : 0;
    for (int k = 0; k < 8; k++) any |= mask[(size_t)l * 8 + k];
    if (any) out[atomicAdd(counter, 1u)] = t.origin[l];
}
__global__ void particle_leaf_origins_kernel(TopoView t, const uint32_t* __restrict__ voxelStart,
                                             int3* __restrict__ out, uint32_t* __restrict__ counter) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= t.n) return;
    if (voxelStart[(size_t)(l + 1) * LEAF] > voxelStart[(size_t)l * LEAF]) out[atomicAdd(counter, 1u)] = t.origin[l];
}
__global__ void pts_counts_remap_kernel(TopoView oldT, TopoView newT, const uint32_t* __restrict__ oldStart,
                                        uint32_t* __restrict__ newCount) {
    int l = blockIdx.x;
    int3 o = oldT.origin[l];
    int nl = topo_find(newT, o.x, o.y, o.z);
    if (nl < 0) return;
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x) {
        size_t k = (size_t)l * LEAF + i;
        newCount[(size_t)nl * LEAF + i] = oldStart[k + 1] - oldStart[k];
    }
}
// dilation: one CTA of 512 threads per leaf, neighbour masks staged in shared memory
// out = dilate(in) by one voxel, 26- or 6-neighbourhood, on whole 64-bit mask words (word X = the x slice, bit y << 3 | z).
// The 26-neighbourhood dilation is separable: OR the x slices X-1, X, X+1, then spread the 8x8 slice along y (shifts by 8 bits,
// rows 7 / 0 of the +-y neighbour leaves carried in), then along z (shifts by 1 bit inside each byte, columns carried in from the
// +-z neighbours) -- each stage on the stage before, for the leaf AND the neighbour leaves it borrows from, so edge and corner
// leaves are covered. One warp per leaf, ~100 word operations (round 1: one thread per voxel, 27 shared-memory bit probes each,
// 34 us per launch at 5800 leaves).
__global__ void __launch_bounds__(128) dilate_kernel(TopoView t, const uint64_t* __restrict__ in,
                                                     uint64_t* __restrict__ out, int nn26) {
    __shared__ uint64_t sIn[4][27 * 8];
    __shared__ uint64_t sA[4][9 * 8];     // after the x stage: (ly, lz) columns of leaves, lx = centre
    __shared__ uint64_t sB[4][3 * 8];     // after the y stage: lz columns, lx = ly = centre
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * 4 + g;
    const bool live = l < t.n;
    uint64_t* I = sIn[g]; uint64_t* A = sA[g]; uint64_t* B = sB[g];
    if (live)
        for (int i = lane; i < 27 * 8; i += 32) {
            const int nb = t.nbr27[(size_t)l * 27 + i / 8];
            I[i] = nb >= 0 ? in[(size_t)nb * 8 + (i & 7)] : 0ull;
        }
    __syncwarp();
    if (!live) return;
    if (!nn26) {
        // faces only: the leaf's own shifts + one plane / row / column of each face neighbour (nbr27 index lx * 9 + ly * 3 + lz)
        if (lane < 8) {
            const int X = lane;
            const uint64_t w = I[13 * 8 + X];
            uint64_t o = w | (w << 8) | (w >> 8) | ((w << 1) & 0xfefefefefefefefeull) | ((w >> 1) & 0x7f7f7f7f7f7f7f7full);
            o |= X > 0 ? I[13 * 8 + X - 1] : I[4 * 8 + 7];
            o |= X < 7 ? I[13 * 8 + X + 1] : I[22 * 8 + 0];
            o |= I[10 * 8 + X] >> 56;                          // row 7 of the -y leaf -> row 0
            o |= (I[16 * 8 + X] & 0xffull) << 56;               // row 0 of the +y leaf -> row 7
            o |= (I[12 * 8 + X] >> 7) & 0x0101010101010101ull;  // column 7 of the -z leaf -> column 0
            o |= (I[14 * 8 + X] & 0x0101010101010101ull) << 7;  // column 0 of the +z leaf -> column 7
            out[(size_t)l * 8 + X] = o;
        }
        return;
    }
    for (int i = lane; i < 72; i += 32) {          // x stage
        const int col = i >> 3, X = i & 7;          // col = ly * 3 + lz
        const uint64_t* c0 = I + (0 * 9 + col) * 8;
        const uint64_t* c1 = I + (1 * 9 + col) * 8;
        const uint64_t* c2 = I + (2 * 9 + col) * 8;
        A[i] = c1[X] | (X > 0 ? c1[X - 1] : c0[7]) | (X < 7 ? c1[X + 1] : c2[0]);
    }
    __syncwarp();
    if (lane < 24) {                               // y stage
        const int lz = lane >> 3, X = lane & 7;
        const uint64_t m = A[(1 * 3 + lz) * 8 + X], lo = A[(0 * 3 + lz) * 8 + X], hi = A[(2 * 3 + lz) * 8 + X];
        B[lane] = m | (m << 8) | (m >> 8) | (lo >> 56) | ((hi & 0xffull) << 56);
    }
    __syncwarp();
    if (lane < 8) {                                // z stage
        const uint64_t m = B[8 + lane], lo = B[lane], hi = B[16 + lane];
        out[(size_t)l * 8 + lane] = m | ((m << 1) & 0xfefefefefefefefeull) | ((m >> 1) & 0x7f7f7f7f7f7f7f7full) |
                                    ((lo >> 7) & 0x0101010101010101ull) | ((hi & 0x0101010101010101ull) << 7);
    }
}
__global__ void popcount_kernel(const uint64_t* __restrict__ mask, size_t nWords, unsigned long long* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned c = i < nWords ? __popcll(mask[i]) : 0u;
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}
__global__ void solid_view_kernel(TopoView pool, TopoView st, const float* __restrict__ sVal, float bg,
                                  float* __restrict__ view, uint8_t* __restrict__ exists) {
    int l = blockIdx.x;
    int3 o = pool.origin[l];
    int sl = st.n > 0 ? topo_find(st, o.x, o.y, o.z) : -1;
    if (exists && threadIdx.x == 0) exists[l] = sl >= 0;
    for (int i = threadIdx.x; i < LEAF; i += blockDim.x)
        view[(size_t)l * LEAF + i] = sl >= 0 ? sVal[(size_t)sl * LEAF + i] : bg;
}
}  // namespace

static inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

TopoPtr topo_from_origins_dev(World* w, const int3* origins_dev, int count, bool ring, const int* bbox) {
    auto t = std::make_shared<Topo>();
    t->epoch = ++w->epochCounter;
    if (count == 0) {
        t->n = 0; t->dmin = make_int3(0, 0, 0); t->ddim = make_int3(0, 0, 0);
        return t;
    }
    int h[6];
    if (bbox) {
        for (int k = 0; k < 6; k++) h[k] = bbox[k];
    } else {
        DBuf<int> bb(6, w->stream);
        int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
        FB_CUDA(cudaMemcpyAsync(bb.p, init, sizeof(init), cudaMemcpyHostToDevice, w->stream));
        FB_LAUNCH(w, "topo_bbox", count * 12) bbox_kernel<<<std::min<unsigned>(nblk(count, BBOX_THREADS), 148 * 8), BBOX_THREADS, 0, w->stream>>>(origins_dev, count, bb.p);
        check_launch("bbox");
        read_back(w, h, bb.p, sizeof(h));
    }
    int r = ring ? 1 : 0;
    t->dmin = make_int3(h[0] - r, h[1] - r, h[2] - r);
    t->ddim = make_int3(h[3] - h[0] + 1 + 2 * r, h[4] - h[1] + 1 + 2 * r, h[5] - h[2] + 1 + 2 * r);
    size_t vol = (size_t)t->ddim.x * t->ddim.y * t->ddim.z;
    FB_REQUIRE(vol <= ((size_t)1 << 28), FLIPB200_ERR_DOMAIN,
               "leaf bounding box exceeds the dense directory limit (2^28 leaves)");
    DBuf<uint32_t> flags(vol, w->stream), slots(vol, w->stream);
    flags.zero();
    FB_LAUNCH(w, "topo_mark", count * 12) mark_kernel<<<nblk(count, 256), 256, 0, w->stream>>>(origins_dev, count, t->dmin, t->ddim, flags.p);
    check_launch("mark");
    if (ring) {
        DBuf<uint32_t> ringed(vol, w->stream);
        FB_LAUNCH(w, "topo_ring", vol * 8) ring_kernel<<<nblk(vol, 256), 256, 0, w->stream>>>(flags.p, t->ddim, ringed.p);
        check_launch("ring");
        flags = std::move(ringed);
    }
    uint64_t total = 0;
    exclusive_scan_u32(w, flags.p, slots.p, vol, &total);
    t->n = (int)total;
    t->dir.alloc(vol, w->stream);
    t->origin.alloc(t->n, w->stream);
    FB_LAUNCH(w, "topo_assign", vol * 12) assign_kernel<<<nblk(vol, 256), 256, 0, w->stream>>>(flags.p, slots.p, t->dmin, t->ddim, t->dir.p, t->origin.p);
    check_launch("assign");
    t->nbr27.alloc((size_t)t->n * 27, w->stream);
    FB_LAUNCH(w, "topo_nbr27", t->n * 27 * 8) nbr27_kernel<<<nblk((size_t)t->n * 27, 256), 256, 0, w->stream>>>(t->view(), t->nbr27.p);
    check_launch("nbr27");
    return t;
}

TopoPtr topo_from_origins_host(World* w, const int32_t* origins, int count, bool ring) {
    DBuf<int3> d(count, w->stream);
    if (count) FB_CUDA(cudaMemcpyAsync(d.p, origins, sizeof(int3) * (size_t)count, cudaMemcpyHostToDevice, w->stream));
    return topo_from_origins_dev(w, d.p, count, ring);
}

static void fill(World* w, float* p, float v, size_t n) {
    if (!n) return;
    if (v == 0.f) { FB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), w->stream)); return; }
    FB_LAUNCH(w, "fill", n * 4) fill_kernel<<<nblk(n, 256), 256, 0, w->stream>>>(p, v, n);
    check_launch("fill");
}

void grid_alloc(World* w, GridF& g, const TopoPtr& t, float bg) {
    g.topo = t; g.bg = bg;
    size_t n = (size_t)t->n;
    g.val.alloc(n * LEAF, w->stream);
    g.mask.alloc(n * 8, w->stream);
    g.alloc.alloc(n ? n : 1, w->stream);
    fill(w, g.val.p, bg, n * LEAF);
    g.mask.zero();
    g.alloc.zero();
}
void grid_alloc(World* w, GridV& g, const TopoPtr& t, const float bg[3]) {
    g.topo = t;
    size_t n = (size_t)t->n;
    for (int c = 0; c < 3; c++) {
        g.bg[c] = bg[c];
        g.val[c].alloc(n * LEAF, w->stream);
        fill(w, g.val[c].p, bg[c], n * LEAF);
    }
    g.mask.alloc(n * 8, w->stream);
    g.mask.zero();
}
void grid_rebase(World* w, GridF& g, const TopoPtr& t) {
    if (g.topo == t) return;
    GridF ng;
    grid_alloc(w, ng, t, g.bg);
    if (g.topo && g.topo->n > 0) {
        FB_LAUNCH(w, "grid_rebase", (size_t)g.topo->n * 4096) rebase_kernel<<<g.topo->n, 256, 0, w->stream>>>(g.topo->view(), t->view(), g.val.p, ng.val.p, g.mask.p, ng.mask.p, g.alloc.p, ng.alloc.p);
        check_launch("rebase");
    }
    g = std::move(ng);
}
void grid_rebase(World* w, GridV& g, const TopoPtr& t) {
    if (g.topo == t) return;
    GridV ng;
    grid_alloc(w, ng, t, g.bg);
    if (g.topo && g.topo->n > 0) {
        for (int c = 0; c < 3; c++) {
            FB_LAUNCH(w, "grid_rebase", (size_t)g.topo->n * 4096) rebase_kernel<<<g.topo->n, 256, 0, w->stream>>>(g.topo->view(), t->view(), g.val[c].p, ng.val[c].p, c == 0 ? g.mask.p : nullptr, ng.mask.p, nullptr, nullptr);
            check_launch("rebase");
        }
    }
    g = std::move(ng);
}
void grid_copy(World* w, GridF& dst, const GridF& src) {
    dst.topo = src.topo; dst.bg = src.bg;
    dst.val.alloc(src.val.n, w->stream);
    dst.mask.alloc(src.mask.n, w->stream);
    dst.alloc.alloc(src.alloc.n, w->stream);
    if (src.alloc.n) FB_CUDA(cudaMemcpyAsync(dst.alloc.p, src.alloc.p, src.alloc.n, cudaMemcpyDeviceToDevice, w->stream));
    if (src.val.n) FB_CUDA(cudaMemcpyAsync(dst.val.p, src.val.p, src.val.n * 4, cudaMemcpyDeviceToDevice, w->stream));
    if (src.mask.n) FB_CUDA(cudaMemcpyAsync(dst.mask.p, src.mask.p, src.mask.n * 8, cudaMemcpyDeviceToDevice, w->stream));
}
void grid_copy(World* w, GridV& dst, const GridV& src) {
    dst.topo = src.topo;
    for (int c = 0; c < 3; c++) {
        dst.bg[c] = src.bg[c];
        dst.val[c].alloc(src.val[c].n, w->stream);
        if (src.val[c].n) FB_CUDA(cudaMemcpyAsync(dst.val[c].p, src.val[c].p, src.val[c].n * 4, cudaMemcpyDeviceToDevice, w->stream));
    }
    dst.mask.alloc(src.mask.n, w->stream);
    if (src.mask.n) FB_CUDA(cudaMemcpyAsync(dst.mask.p, src.mask.p, src.mask.n * 8, cudaMemcpyDeviceToDevice, w->stream));
}

void ensure_pool(World* w, std::initializer_list<int> gridIds, bool includeParticles) {
    bool ok = (bool)w->pool;
    auto topoOf = [&](int id) -> TopoPtr { return is_vec_grid(id) ? w->V(id).topo : w->F(id).topo; };
    if (ok) {
        for (int id : gridIds) { TopoPtr t = topoOf(id); if (t && t != w->pool) ok = false; }
        if (includeParticles && w->pts.topo && w->pts.topo != w->pool) ok = false;
    }
    if (!ok) {
        // candidate leaves: every non-empty leaf of the listed grids + every leaf holding particles
        size_t cap = 0;
        for (int id : gridIds) { TopoPtr t = topoOf(id); if (t) cap += t->n; }
        if (w->pts.topo) cap += w->pts.topo->n;
        DBuf<int3> cand(cap ? cap : 1, w->stream);
        DBuf<uint32_t> counter(1, w->stream);
        counter.zero();
        for (int id : gridIds) {
            TopoPtr t = topoOf(id);
            if (!t || t->n == 0) continue;
            const uint64_t* m = is_vec_grid(id) ? w->V(id).mask.p : w->F(id).mask.p;
            const uint8_t* al = is_vec_grid(id) ? nullptr : w->F(id).alloc.p;
            FB_LAUNCH(w, "pool_candidates", t->n * 76) nonempty_origins_kernel<<<nblk(t->n, 128), 128, 0, w->stream>>>(t->view(), m, al, cand.p, counter.p);
            check_launch("nonempty_origins");
        }
        if (w->pts.topo && w->pts.topo->n > 0 && w->pts.n > 0) {
            FB_LAUNCH(w, "pool_candidates", w->pts.topo->n * 20) particle_leaf_origins_kernel<<<nblk(w->pts.topo->n, 128), 128, 0, w->stream>>>(w->pts.topo->view(), w->pts.voxelStart.p, cand.p, counter.p);
            check_launch("particle_leaf_origins");
        }
        uint32_t cnt = 0;
        read_back(w, &cnt, counter.p, 4);
        w->pool = topo_from_origins_dev(w, cand.p, (int)cnt, /*ring=*/true);
    }
    for (int id : gridIds) {
        if (is_vec_grid(id)) {
            if (!w->V(id).topo) grid_alloc(w, w->V(id), w->pool, w->V(id).bg);
            else grid_rebase(w, w->V(id), w->pool);
        } else {
            if (!w->F(id).topo) grid_alloc(w, w->F(id), w->pool, w->F(id).bg);
            else grid_rebase(w, w->F(id), w->pool);
        }
    }
    if (includeParticles && w->pts.topo != w->pool) {
        // same lexicographic leaf order in both topologies: only the per-voxel offsets move
        size_t nv = (size_t)w->pool->n * LEAF;
        DBuf<uint32_t> ns(nv + 1, w->stream);
        ns.zero();
        if (w->pts.topo && w->pts.topo->n > 0 && w->pts.n > 0) {
            FB_LAUNCH(w, "pts_remap", (size_t)w->pts.topo->n * 4096) pts_counts_remap_kernel<<<w->pts.topo->n, 256, 0, w->stream>>>(w->pts.topo->view(), w->pool->view(), w->pts.voxelStart.p, ns.p);
            check_launch("pts_remap");
        }
        exclusive_scan_u32(w, ns.p, ns.p, nv + 1, nullptr);
        w->pts.voxelStart = std::move(ns);
        w->pts.topo = w->pool;
    }
}

void refresh_solid_views(World* w) {
    if (!w->pool) return;
    if (w->solidViewEpoch == w->pool->epoch) return;
    size_t n = (size_t)w->pool->n;
    w->solidSdfView.alloc(n * LEAF, w->stream);
    w->solidLeafExists.alloc(n ? n : 1, w->stream);
    for (int c = 0; c < 3; c++) w->solidVelView[c].alloc(n * LEAF, w->stream);
    if (n) {
        GridF& s = w->F(FLIPB200_SOLID_SDF);
        TopoView sv = s.topo ? s.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
        FB_LAUNCH(w, "solid_view", n * 4096) solid_view_kernel<<<(unsigned)n, 256, 0, w->stream>>>(w->pool->view(), sv, s.val.p, s.bg, w->solidSdfView.p, w->solidLeafExists.p);
        check_launch("solid_view");
        GridV& u = w->V(FLIPB200_SOLID_VELOCITY);
        TopoView uv = u.topo ? u.topo->view() : TopoView{0, make_int3(0, 0, 0), make_int3(0, 0, 0), nullptr, nullptr, nullptr};
        for (int c = 0; c < 3; c++) {
            FB_LAUNCH(w, "solid_view", n * 4096) solid_view_kernel<<<(unsigned)n, 256, 0, w->stream>>>(w->pool->view(), uv, u.val[c].p, u.bg[c], w->solidVelView[c].p, nullptr);
            check_launch("solid_view");
        }
    }
    w->solidViewEpoch = w->pool->epoch;
}

void mask_dilate(World* w, const Topo& t, const uint64_t* in, uint64_t* out, bool nn26) {
    if (t.n == 0) return;
    FB_LAUNCH(w, "mask_dilate", (size_t)t.n * 128) dilate_kernel<<<(t.n + 3) / 4, 128, 0, w->stream>>>(t.view(), in, out, nn26 ? 1 : 0);
    check_launch("dilate");
}
void mask_count_async(World* w, const uint64_t* mask, int nLeaves, unsigned long long* out) {
    size_t nw = (size_t)nLeaves * 8;
    if (!nw) return;
    FB_LAUNCH(w, "mask_count", nw * 8) popcount_kernel<<<nblk(nw, 256), 256, 0, w->stream>>>(mask, nw, out);
    check_launch("popcount");
}
uint64_t mask_count(World* w, const uint64_t* mask, int nLeaves) {
    DBuf<unsigned long long> c(1, w->stream);
    c.zero();
    size_t nw = (size_t)nLeaves * 8;
    if (nw) {
        FB_LAUNCH(w, "mask_count", nw * 8) popcount_kernel<<<nblk(nw, 256), 256, 0, w->stream>>>(mask, nw, c.p);
        check_launch("popcount");
    }
    unsigned long long h = 0;
    read_back(w, &h, c.p, 8);
    return h;
}

}  // namespace fb
