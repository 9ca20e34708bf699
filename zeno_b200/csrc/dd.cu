// libflipb200 -- slab decomposition of one FLIP world over the GPUs of a box.
//
// Rank r owns the leaf layers [lo_r, hi_r) along x (leaf coordinate = voxel >> 3). Because pool slots are numbered
// lexicographically in (x,y,z) leaf coordinates, every leaf layer -- and therefore the owned region, the boundary
// layers a neighbour needs and the ghost layers received from it -- is ONE contiguous slot range: boundary layers are
// sent straight out of the grid arrays, no pack kernel.
//
//   particles : a rank stores its owned particles plus one ghost leaf layer on each side (what the collect-style P2G
//               reads, FF/FLIP_vdb.cpp:1137-1263); after every move (advection, initial binning) dd_migrate sends
//               migrants + ghost copies to the two neighbours and returns [from left | kept | from right]. Left ranks
//               own lower x, i.e. lower slots of the undecomposed store, so a stable sort of that sequence reproduces
//               the single-GPU order inside every voxel (bit-identical P2G sums, same particles dropped at the cap).
//   grids     : every node runs unchanged on the local pool (owned + ghost + ring layers); what it computes within a
//               few voxels of the pool's outer x faces is incomplete and is overwritten by dd_refresh, which copies
//               whole leaves of the two layers next to each slab face from their owner.
//   solver    : poisson.cu (level 0 sharded with one ghost-layer exchange per 8 colour passes, levels >= 1 assembled
//               globally and replicated).
#include "world.cuh"
#include <climits>
#include <algorithm>

namespace fb {

constexpr int DD_OPEN = 1 << 29;   // "no neighbour on this side"

struct DDMaps {
    uint64_t epoch = ~0ull;
    int b[6] = {0, 0, 0, 0, 0, 0};   // slots with leaf x < lo, lo+1, lo+2, hi-2, hi-1, hi
    int recvCnt[2][2] = {{0, 0}, {0, 0}};   // [0 = from left, 1 = from right][1 layer, 2 layers]
    DBuf<int> map[2];                // received leaf (sender's 2-layer list order) -> my slot or -1
};
struct DDState {
    bool on = false;
    int lo = -DD_OPEN, hi = DD_OPEN;
    bool boundsChecked = false;          // the neighbours' slabs were compared with mine (gap-free, no overlap)
    int leftLo = -DD_OPEN, rightHi = DD_OPEN;   // the neighbours' far bounds: a particle beyond them cannot be routed in one hop
    DBuf<int> strayed;                   // device counter of such particles
    DDMaps maps;
};

namespace {
inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

__global__ void layer_bounds_kernel(const int3* __restrict__ origin, int n, const int* __restrict__ xs, int* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= 6) return;
    const int x = xs[k];
    int a = 0, b = n;   // first slot whose leaf x >= x (origins are sorted by x first)
    while (a < b) { int m = (a + b) >> 1; if ((origin[m].x >> 3) < x) a = m + 1; else b = m; }
    out[k] = a;
}
__global__ void map_kernel(TopoView t, const int3* __restrict__ origins, int n, int* __restrict__ map) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int3 o = origins[i];
    map[i] = topo_find(t, o.x, o.y, o.z);
}
constexpr int DD_MAX_ARRAYS = 16;
struct UnpackArgs { int nArr; char* dst[DD_MAX_ARRAYS]; const char* src[DD_MAX_ARRAYS]; int bpl[DD_MAX_ARRAYS]; };
__global__ void unpack_kernel(UnpackArgs a, const int* __restrict__ map) {
    const int s = map[blockIdx.x];
    if (s < 0) return;
    const int arr = blockIdx.y, bpl = a.bpl[arr];
    const char* src = a.src[arr] + (size_t)blockIdx.x * bpl;
    char* dst = a.dst[arr] + (size_t)s * bpl;
    if ((bpl & 15) == 0) {
        const uint4* sp = reinterpret_cast<const uint4*>(src);
        uint4* dp = reinterpret_cast<uint4*>(dst);
        for (int i = threadIdx.x; i < (bpl >> 4); i += blockDim.x) dp[i] = sp[i];
    } else {
        for (int i = threadIdx.x; i < bpl; i += blockDim.x) dst[i] = src[i];
    }
}
// migration: 1 = goes to the left neighbour, 2 = to the right, 4 = stays in this rank's extended region
__global__ void migrate_flags_kernel(const int3* __restrict__ ijk, const uint8_t* __restrict__ alive, uint64_t pLo, uint64_t m,
                                     int lo, int hi, int hasLeft, int hasRight, uint32_t* __restrict__ fL,
                                     uint32_t* __restrict__ fR, uint32_t* __restrict__ fK, int leftLo, int rightHi, int* __restrict__ strayed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t g = pLo + i;
    const bool a = alive ? alive[g] != 0 : true;
    const int lx = ijk[g].x >> 3;
    // beyond the neighbour's own slab: it would sit there as a ghost nobody advects (migration is one hop)
    if (a && ((hasLeft && lx < leftLo) || (hasRight && lx >= rightHi))) atomicAdd(strayed, 1);
    fL[i] = (a && hasLeft && lx < lo + 1) ? 1u : 0u;
    fR[i] = (a && hasRight && lx >= hi - 1) ? 1u : 0u;
    fK[i] = (a && (!hasLeft || lx >= lo - 1) && (!hasRight || lx < hi + 1)) ? 1u : 0u;
}
struct CompactDst { uint32_t *w0, *w1, *w2; int3* ijk; };
__global__ void migrate_compact_kernel(const uint32_t* __restrict__ w0, const uint32_t* __restrict__ w1, const uint32_t* __restrict__ w2,
                                       const int3* __restrict__ ijk, uint64_t pLo, uint64_t m, const uint32_t* __restrict__ fL,
                                       const uint32_t* __restrict__ fR, const uint32_t* __restrict__ fK, const uint32_t* __restrict__ pL,
                                       const uint32_t* __restrict__ pR, const uint32_t* __restrict__ pK, CompactDst L, CompactDst R,
                                       CompactDst K) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t g = pLo + i;
    const uint32_t a = w0[g], b = w1[g], c = w2[g];
    const int3 v = ijk[g];
    if (fL[i]) { uint32_t d = pL[i]; L.w0[d] = a; L.w1[d] = b; L.w2[d] = c; L.ijk[d] = v; }
    if (fR[i]) { uint32_t d = pR[i]; R.w0[d] = a; R.w1[d] = b; R.w2[d] = c; R.ijk[d] = v; }
    if (fK[i]) { uint32_t d = pK[i]; K.w0[d] = a; K.w1[d] = b; K.w2[d] = c; K.ijk[d] = v; }
}

// two ints to each neighbour, two from each
void exchange_counts(World* w, const int toLeft[2], const int toRight[2], int fromLeft[2], int fromRight[2]) {
    const bool hasL = w->rank > 0, hasR = w->rank < w->nRanks - 1;
    DBuf<int> buf(8, w->stream);
    int h[8] = {toLeft[0], toLeft[1], toRight[0], toRight[1], 0, 0, 0, 0};
    FB_CUDA(cudaMemcpyAsync(buf.p, h, sizeof(h), cudaMemcpyHostToDevice, w->stream));
    comm_group_begin(w);
    if (hasL) { comm_send(w, w->rank - 1, buf.p, 8); comm_recv(w, w->rank - 1, buf.p + 4, 8); }
    if (hasR) { comm_send(w, w->rank + 1, buf.p + 2, 8); comm_recv(w, w->rank + 1, buf.p + 6, 8); }
    comm_group_end(w);
    FB_CUDA(cudaMemcpyAsync(h, buf.p, sizeof(h), cudaMemcpyDeviceToHost, w->stream));
    sync(w);
    fromLeft[0] = h[4]; fromLeft[1] = h[5]; fromRight[0] = h[6]; fromRight[1] = h[7];
}

DDMaps& dd_maps(World* w) {
    DDState& D = *w->dd;
    DDMaps& M = D.maps;
    FB_REQUIRE(w->pool != nullptr, FLIPB200_ERR_STATE, "slab decomposition: no pool yet");
    if (M.epoch == w->pool->epoch) return M;
    FB_PHASE(w, "dd_maps rebuild");
    const Topo& t = *w->pool;
    const bool hasL = w->rank > 0, hasR = w->rank < w->nRanks - 1;
    const int lo = hasL ? D.lo : -DD_OPEN, hi = hasR ? D.hi : DD_OPEN;
    int xs[6] = {lo, std::min(lo + 1, hi), std::min(lo + 2, hi), std::max(hi - 2, lo), std::max(hi - 1, lo), hi};
    if (t.n > 0) {
        DBuf<int> d(12, w->stream);
        FB_CUDA(cudaMemcpyAsync(d.p, xs, sizeof(xs), cudaMemcpyHostToDevice, w->stream));
        FB_LAUNCH(w, "dd_layer_bounds", 64) layer_bounds_kernel<<<1, 32, 0, w->stream>>>(t.origin.p, t.n, d.p, d.p + 6);
        check_launch("layer_bounds");
        FB_CUDA(cudaMemcpyAsync(M.b, d.p + 6, sizeof(M.b), cudaMemcpyDeviceToHost, w->stream));
        sync(w);
    } else {
        for (int k = 0; k < 6; k++) M.b[k] = 0;
    }
    const int toLeft[2] = {M.b[1] - M.b[0], M.b[2] - M.b[0]};
    const int toRight[2] = {M.b[5] - M.b[4], M.b[5] - M.b[3]};
    exchange_counts(w, toLeft, toRight, M.recvCnt[0], M.recvCnt[1]);
    // origins of the two-layer lists, then where each received leaf lives in my pool
    DBuf<int3> in[2];
    in[0].alloc(M.recvCnt[0][1] + 1, w->stream);
    in[1].alloc(M.recvCnt[1][1] + 1, w->stream);
    comm_group_begin(w);
    if (hasL) {
        comm_send(w, w->rank - 1, t.origin.p + M.b[0], (size_t)toLeft[1] * sizeof(int3));
        comm_recv(w, w->rank - 1, in[0].p, (size_t)M.recvCnt[0][1] * sizeof(int3));
    }
    if (hasR) {
        comm_send(w, w->rank + 1, t.origin.p + M.b[3], (size_t)toRight[1] * sizeof(int3));
        comm_recv(w, w->rank + 1, in[1].p, (size_t)M.recvCnt[1][1] * sizeof(int3));
    }
    comm_group_end(w);
    for (int s = 0; s < 2; s++) {
        const int n = M.recvCnt[s][1];
        M.map[s].alloc(n + 1, w->stream);
        if (n) {
            FB_LAUNCH(w, "dd_map", (size_t)n * 16) map_kernel<<<nblk(n, 128), 128, 0, w->stream>>>(t.view(), in[s].p, n, M.map[s].p);
            check_launch("dd_map");
        }
    }
    sync(w);
    M.epoch = t.epoch;
    return M;
}
}  // namespace

bool dd_on(World* w) { return w->dd && w->dd->on && comm_active(w); }
void dd_destroy(World* w) { delete w->dd; w->dd = nullptr; }
void dd_set_slab(World* w, int lo, int hi) {
    FB_REQUIRE(comm_active(w), FLIPB200_ERR_COMM, "dd_set_slab: initialise the communicator first");
    FB_REQUIRE(hi - lo >= 2, FLIPB200_ERR_ARG, "dd_set_slab: a slab must hold at least two leaf layers (ghost exchange reads the two layers next to each face)");
    if (!w->dd) w->dd = new DDState();
    w->dd->on = true;
    w->dd->lo = lo; w->dd->hi = hi;
    w->dd->boundsChecked = false;
    w->dd->maps.epoch = ~0ull;
}
namespace {
// once per dd_set_slab (collective, at the first exchange): my neighbours' slabs must continue mine without gap or overlap
void check_bounds(World* w) {
    DDState& D = *w->dd;
    if (D.boundsChecked) return;
    const int mine[2] = {D.lo, D.hi};
    int fromLeft[2] = {0, 0}, fromRight[2] = {0, 0};
    exchange_counts(w, mine, mine, fromLeft, fromRight);
    const bool hasL = w->rank > 0, hasR = w->rank < w->nRanks - 1;
    FB_REQUIRE(!hasL || fromLeft[1] == D.lo, FLIPB200_ERR_ARG, "dd_set_slab: rank " + std::to_string(w->rank - 1) + " owns leaf layers [" + std::to_string(fromLeft[0]) + ", " +
               std::to_string(fromLeft[1]) + "), rank " + std::to_string(w->rank) + " starts at " + std::to_string(D.lo) + ": the slabs must be contiguous");
    FB_REQUIRE(!hasR || fromRight[0] == D.hi, FLIPB200_ERR_ARG, "dd_set_slab: rank " + std::to_string(w->rank + 1) + " starts at leaf layer " + std::to_string(fromRight[0]) +
               ", rank " + std::to_string(w->rank) + " ends at " + std::to_string(D.hi) + ": the slabs must be contiguous");
    D.leftLo = hasL ? (w->rank - 1 > 0 ? fromLeft[0] : -DD_OPEN) : -DD_OPEN;
    D.rightHi = hasR ? (w->rank + 1 < w->nRanks - 1 ? fromRight[1] : DD_OPEN) : DD_OPEN;
    D.boundsChecked = true;
}
}  // namespace
void dd_owned_slots(World* w, int* ownLo, int* ownHi) {
    DDMaps& M = dd_maps(w);
    *ownLo = M.b[0]; *ownHi = M.b[5];
}
void dd_owned_coords(World* w, int* lo, int* hi) {
    *lo = w->rank > 0 ? w->dd->lo : -DD_OPEN;
    *hi = w->rank < w->nRanks - 1 ? w->dd->hi : DD_OPEN;
}

void dd_refresh(World* w, const std::vector<DDArray>& arrays, int layers) {
    if (!dd_on(w) || arrays.empty()) return;
    FB_REQUIRE((int)arrays.size() <= DD_MAX_ARRAYS && (layers == 1 || layers == 2), FLIPB200_ERR_ARG, "dd_refresh: bad argument");
    DDMaps& M = dd_maps(w);
    FB_PHASE(w, layers == 1 ? "dd_refresh 1 layer" : "dd_refresh 2 layers");
    const bool has[2] = {w->rank > 0, w->rank < w->nRanks - 1};
    const int peer[2] = {w->rank - 1, w->rank + 1};
    // what I send: to the left the first layers of my slab, to the right the last ones
    const int sendStart[2] = {M.b[0], layers == 1 ? M.b[4] : M.b[3]};
    const int sendCnt[2] = {layers == 1 ? M.b[1] - M.b[0] : M.b[2] - M.b[0], layers == 1 ? M.b[5] - M.b[4] : M.b[5] - M.b[3]};
    // what I receive: from the left its last layers (the single layer is the tail of its two-layer list), from the right its first
    const int recvCnt[2] = {M.recvCnt[0][layers - 1], M.recvCnt[1][layers - 1]};
    const int mapOff[2] = {layers == 1 ? M.recvCnt[0][1] - M.recvCnt[0][0] : 0, 0};
    DBuf<char> stage[2];
    std::vector<size_t> off[2];
    uint64_t bytes = 0;
    for (int s = 0; s < 2; s++) {
        size_t cur = 0;
        for (auto& a : arrays) { off[s].push_back(cur); cur += (((size_t)recvCnt[s] * a.bytesPerLeaf) + 255) & ~(size_t)255; }
        stage[s].alloc(cur + 256, w->stream);
        if (has[s]) bytes += cur;
    }
    comm_group_begin(w);
    for (int s = 0; s < 2; s++) {
        if (!has[s]) continue;
        for (size_t k = 0; k < arrays.size(); k++) {
            const auto& a = arrays[k];
            comm_send(w, peer[s], (const char*)a.base + (size_t)sendStart[s] * a.bytesPerLeaf, (size_t)sendCnt[s] * a.bytesPerLeaf);
            comm_recv(w, peer[s], stage[s].p + off[s][k], (size_t)recvCnt[s] * a.bytesPerLeaf);
        }
    }
    comm_group_end(w);
    for (int s = 0; s < 2; s++) {
        if (!has[s] || recvCnt[s] == 0) continue;
        UnpackArgs u;
        u.nArr = (int)arrays.size();
        for (size_t k = 0; k < arrays.size(); k++) { u.dst[k] = (char*)arrays[k].base; u.src[k] = stage[s].p + off[s][k]; u.bpl[k] = arrays[k].bytesPerLeaf; }
        FB_LAUNCH(w, "dd_unpack", 2 * bytes) unpack_kernel<<<dim3(recvCnt[s], u.nArr), 128, 0, w->stream>>>(u, M.map[s].p + mapOff[s]);
        check_launch("dd_unpack");
    }
}
void dd_refresh(World* w, GridF& g, int layers) {
    if (!dd_on(w) || !g.topo) return;
    FB_REQUIRE(g.topo == w->pool, FLIPB200_ERR_STATE, "dd_refresh: grid is not on the pool");
    dd_refresh(w, {DDArray{g.val.p, LEAF * 4}, DDArray{g.mask.p, 64}, DDArray{g.alloc.p, 1}}, layers);
}
void dd_refresh(World* w, GridV& g, int layers) {
    if (!dd_on(w) || !g.topo) return;
    FB_REQUIRE(g.topo == w->pool, FLIPB200_ERR_STATE, "dd_refresh: grid is not on the pool");
    dd_refresh(w, {DDArray{g.val[0].p, LEAF * 4}, DDArray{g.val[1].p, LEAF * 4}, DDArray{g.val[2].p, LEAF * 4}, DDArray{g.mask.p, 64}}, layers);
}

void dd_migrate(World* w, uint64_t pLo, uint64_t pHi, const uint32_t* w0, const uint32_t* w1, const uint32_t* w2, const int3* ijk,
                const uint8_t* alive, DBuf<uint32_t>& o0, DBuf<uint32_t>& o1, DBuf<uint32_t>& o2, DBuf<int3>& oijk, uint64_t* nOut) {
    const bool hasL = w->rank > 0, hasR = w->rank < w->nRanks - 1;
    const uint64_t m = pHi - pLo;
    DBuf<uint32_t> fL(m + 1, w->stream), fR(m + 1, w->stream), fK(m + 1, w->stream);
    DBuf<uint32_t> pL(m + 1, w->stream), pR(m + 1, w->stream), pK(m + 1, w->stream);
    fL.zero(); fR.zero(); fK.zero();
    check_bounds(w);
    DDState& D = *w->dd;
    if (!D.strayed.p) D.strayed.alloc(1, w->stream);
    D.strayed.zero();
    if (m) {
        FB_LAUNCH(w, "dd_migrate_flags", m * 25) migrate_flags_kernel<<<nblk(m, 256), 256, 0, w->stream>>>(ijk, alive, pLo, m, w->dd->lo, w->dd->hi, hasL ? 1 : 0, hasR ? 1 : 0, fL.p, fR.p, fK.p,
                                                                                                       D.leftLo, D.rightHi, D.strayed.p);
        check_launch("migrate_flags");
    }
    uint64_t cL = 0, cR = 0, cK = 0;
    FB_PHASE(w, "dd_migrate after flags");
    exclusive_scan_u32(w, fL.p, pL.p, m + 1, &cL);
    exclusive_scan_u32(w, fR.p, pR.p, m + 1, &cR);
    exclusive_scan_u32(w, fK.p, pK.p, m + 1, &cK);
    const int toLeft[2] = {(int)cL, 0}, toRight[2] = {(int)cR, 0};
    int fromLeft[2], fromRight[2];
    comm_allreduce(w, D.strayed.p, 1, CT_I32, false);     // every rank must see the same verdict (a lone throw would hang the others)
    d2h_words(w, w->hostScratch + 512, D.strayed.p, 4);   // arrives with the wait inside exchange_counts
    exchange_counts(w, toLeft, toRight, fromLeft, fromRight);
    const int strayed = *reinterpret_cast<const int*>(w->hostScratch + 512);
    FB_REQUIRE(strayed == 0, FLIPB200_ERR_DOMAIN, std::to_string(strayed) + " particles moved past a neighbour's whole slab in one step (migration is one hop): use thicker slabs or a smaller time step");
    const uint64_t rL = hasL ? (uint64_t)fromLeft[0] : 0, rR = hasR ? (uint64_t)fromRight[0] : 0;
    const uint64_t n = rL + cK + rR;
    o0.alloc(n + 1, w->stream); o1.alloc(n + 1, w->stream); o2.alloc(n + 1, w->stream); oijk.alloc(n + 1, w->stream);
    DBuf<uint32_t> sL0(cL + 1, w->stream), sL1(cL + 1, w->stream), sL2(cL + 1, w->stream), sR0(cR + 1, w->stream), sR1(cR + 1, w->stream), sR2(cR + 1, w->stream);
    DBuf<int3> sLi(cL + 1, w->stream), sRi(cR + 1, w->stream);
    if (m) {
        CompactDst L{sL0.p, sL1.p, sL2.p, sLi.p}, R{sR0.p, sR1.p, sR2.p, sRi.p}, K{o0.p + rL, o1.p + rL, o2.p + rL, oijk.p + rL};
        FB_LAUNCH(w, "dd_migrate_compact", m * 48) migrate_compact_kernel<<<nblk(m, 256), 256, 0, w->stream>>>(w0, w1, w2, ijk, pLo, m, fL.p, fR.p, fK.p, pL.p, pR.p, pK.p, L, R, K);
        check_launch("migrate_compact");
    }
    comm_group_begin(w);
    if (hasL) {
        comm_send(w, w->rank - 1, sL0.p, cL * 4); comm_send(w, w->rank - 1, sL1.p, cL * 4); comm_send(w, w->rank - 1, sL2.p, cL * 4); comm_send(w, w->rank - 1, sLi.p, cL * 12);
        comm_recv(w, w->rank - 1, o0.p, rL * 4); comm_recv(w, w->rank - 1, o1.p, rL * 4); comm_recv(w, w->rank - 1, o2.p, rL * 4); comm_recv(w, w->rank - 1, oijk.p, rL * 12);
    }
    if (hasR) {
        const uint64_t at = rL + cK;
        comm_send(w, w->rank + 1, sR0.p, cR * 4); comm_send(w, w->rank + 1, sR1.p, cR * 4); comm_send(w, w->rank + 1, sR2.p, cR * 4); comm_send(w, w->rank + 1, sRi.p, cR * 12);
        comm_recv(w, w->rank + 1, o0.p + at, rR * 4); comm_recv(w, w->rank + 1, o1.p + at, rR * 4); comm_recv(w, w->rank + 1, o2.p + at, rR * 4); comm_recv(w, w->rank + 1, oijk.p + at, rR * 12);
    }
    comm_group_end(w);
    sync(w);   // the send buffers are temporaries of this call
    *nOut = n;
}

}  // namespace fb

extern "C" {
int flipb200_dd_set_slab(flipb200_world* w, int leafLo, int leafHi) {
    try {
        if (!w) return FLIPB200_ERR_ARG;
        cudaSetDevice(w->device);
        fb::dd_set_slab(w, leafLo, leafHi);
        return FLIPB200_OK;
    } catch (const fb::Error& e) { fb::set_last_error(e.what()); return e.code; } catch (...) { return FLIPB200_ERR_ARG; }
}
int flipb200_dd_owned_particles(flipb200_world* w, uint64_t* n) {
    try {
        if (!w || !n || !fb::dd_on(w)) return FLIPB200_ERR_STATE;
        cudaSetDevice(w->device);
        *n = 0;
        if (!w->pts.topo || w->pts.topo != w->pool) return FLIPB200_ERR_STATE;
        int a = 0, b = 0;
        fb::dd_owned_slots(w, &a, &b);
        uint32_t r[2] = {0, 0};
        if (w->pool->n == 0) return FLIPB200_OK;
        FB_CUDA(cudaMemcpyAsync(&r[0], w->pts.voxelStart.p + (size_t)a * fb::LEAF, 4, cudaMemcpyDeviceToHost, w->stream));
        FB_CUDA(cudaMemcpyAsync(&r[1], w->pts.voxelStart.p + (size_t)b * fb::LEAF, 4, cudaMemcpyDeviceToHost, w->stream));
        fb::sync(w);
        *n = r[1] - r[0];
        return FLIPB200_OK;
    } catch (const fb::Error& e) { fb::set_last_error(e.what()); return e.code; } catch (...) { return FLIPB200_ERR_ARG; }
}
int flipb200_dd_owned(flipb200_world* w, int* leafLo, int* leafHi) {
    if (!w || !w->dd || !leafLo || !leafHi) return FLIPB200_ERR_STATE;
    fb::dd_owned_coords(w, leafLo, leafHi);
    return FLIPB200_OK;
}
}
