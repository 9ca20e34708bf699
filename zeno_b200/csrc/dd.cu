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
// Peer-memory ghost exchange (dd_refresh): every rank owns one device allocation [4 boxes | flags] -- a receive box per
// (side, parity) and a 32-bit epoch flag per box -- that its two neighbours map (cudaIpc between processes, the plain pointer
// inside one process). An exchange is then two launches per side instead of an NCCL send/recv group + unpack: a push kernel
// copies my boundary leaves STRAIGHT INTO the neighbour's box over NVLink and, from its last CTA, stores the exchange's epoch
// into the neighbour's flag (system-scope fence before it); a one-thread kernel waits for MY flag to reach the epoch, then the
// usual unpack kernel scatters the box into my ghost leaves. Boxes alternate by epoch parity: the neighbour's push k+2 into a box
// is ordered after its wait for my push k+1, which my stream issued after my unpack k of that box.
struct DDPeer {
    bool tried = false, ready = false;
    size_t boxBytes = 0;
    char* mine = nullptr;              // [2 sides][2 parities] boxes, then the flags
    char* remote[2] = {nullptr, nullptr};   // the neighbour's allocation (left, right)
    bool opened[2] = {false, false};   // remote[s] came from cudaIpcOpenMemHandle
    unsigned epoch[2] = {0, 0};        // exchanges done with each side
    DBuf<unsigned> counter;            // last-CTA counters of the push kernels, one per side
    int* errHost = nullptr;            // a word of the world's mapped scratch block: a wait ran into its time limit
    char* box(char* base, int side, int parity) const { return base + (size_t)(side * 2 + parity) * boxBytes; }
    unsigned* flag(char* base, int side, int parity) const { return reinterpret_cast<unsigned*>(base + 4 * boxBytes) + (side * 2 + parity) * 32; }
    size_t total() const { return 4 * boxBytes + 4 * 32 * sizeof(unsigned); }
    ~DDPeer() {
        for (int s = 0; s < 2; s++) if (opened[s] && remote[s]) cudaIpcCloseMemHandle(remote[s]);
        if (mine) cudaFree(mine);
    }
};
struct DDState {
    DDPeer peer;
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
    __syncwarp();
    // what the neighbours receive from me: one / two layers to the left, one / two to the right (out[6..9])
    if (k == 0) { out[6] = out[1] - out[0]; out[7] = out[2] - out[0]; out[8] = out[5] - out[4]; out[9] = out[5] - out[3]; }
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
// ---- peer-memory exchange
struct PushArgs { int nArr; const char* src[DD_MAX_ARRAYS]; size_t bytes[DD_MAX_ARRAYS]; size_t off[DD_MAX_ARRAYS]; };
__global__ void __launch_bounds__(256) dd_push_kernel(PushArgs a, char* __restrict__ box, unsigned* __restrict__ remoteFlag, unsigned epoch,
                                                      unsigned* __restrict__ counter) {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < a.nArr; k++) {
        char* dst = box + a.off[k];   // box offsets are 256-byte aligned
        if ((reinterpret_cast<uintptr_t>(a.src[k]) & 15) == 0) {
            const size_t n16 = a.bytes[k] >> 4;
            const uint4* sp = reinterpret_cast<const uint4*>(a.src[k]);
            uint4* dp = reinterpret_cast<uint4*>(dst);
            size_t i = gtid;
            for (; i + 3 * gsz < n16; i += 4 * gsz) {   // four 16-byte stores in flight per thread
                const uint4 v0 = sp[i], v1 = sp[i + gsz], v2 = sp[i + 2 * gsz], v3 = sp[i + 3 * gsz];
                dp[i] = v0; dp[i + gsz] = v1; dp[i + 2 * gsz] = v2; dp[i + 3 * gsz] = v3;
            }
            for (; i < n16; i += gsz) dp[i] = sp[i];
            for (size_t i = (n16 << 4) + gtid; i < a.bytes[k]; i += gsz) dst[i] = a.src[k][i];
        } else {   // the one-byte-per-leaf arrays start anywhere
            for (size_t i = gtid; i < a.bytes[k]; i += gsz) dst[i] = a.src[k][i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // one fence per CTA, behind the barrier: cumulative over the stores of all its threads
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {   // every CTA's stores are fenced: publish
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(remoteFlag) = epoch;
            *counter = 0;   // for the next push to this side (a later kernel of this stream)
        }
    }
}
__global__ void dd_wait_kernel(const unsigned* __restrict__ flag, unsigned epoch, int* __restrict__ err) {
    const volatile unsigned* f = reinterpret_cast<const volatile unsigned*>(flag);
    long long t0 = clock64();
    // epochs only grow; a (wrapping) difference >= 0 means the neighbour's push of this exchange has landed
    while ((int)(*f - epoch) < 0) {
        if (clock64() - t0 > 30000000000ll) { *err = 1; break; }   // ~15 s: a peer that never pushes must not hang the device
        __nanosleep(100);
    }
    __threadfence_system();
}

// migration: a particle goes to the left neighbour (1), to the right one (2) and / or stays in this rank's extended region (4).
// One stable three-way partition in two passes over the particles (count per block of 1024, scatter) with a one-CTA scan of the
// block counts between them; round 1 wrote three flag arrays, scanned each over all particles (three host-read totals) and
// compacted in a fourth pass.
constexpr int MG_THREADS = 256, MG_ITEMS = 4, MG_BLOCK = MG_THREADS * MG_ITEMS;
struct MigrateArgs {
    const int3* ijk; const uint8_t* alive; uint64_t pLo, m;
    int lo, hi, hasLeft, hasRight, leftLo, rightHi;
};
__device__ __forceinline__ unsigned migrate_flags(const MigrateArgs& A, uint64_t i, bool& stray) {
    const uint64_t g = A.pLo + i;
    const bool a = A.alive ? A.alive[g] != 0 : true;
    const int lx = A.ijk[g].x >> 3;
    // beyond the neighbour's own slab: it would sit there as a ghost nobody advects (migration is one hop)
    stray = a && ((A.hasLeft && lx < A.leftLo) || (A.hasRight && lx >= A.rightHi));
    unsigned f = 0;
    if (a && A.hasLeft && lx < A.lo + 1) f |= 1u;
    if (a && A.hasRight && lx >= A.hi - 1) f |= 2u;
    if (a && (!A.hasLeft || lx >= A.lo - 1) && (!A.hasRight || lx < A.hi + 1)) f |= 4u;
    return f;
}
__global__ void __launch_bounds__(MG_THREADS) migrate_count_kernel(MigrateArgs A, uint32_t* __restrict__ blockCnt, unsigned nb, int* __restrict__ strayed) {
    __shared__ uint32_t sm[3][MG_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * MG_BLOCK + (uint64_t)threadIdx.x * MG_ITEMS;
    uint32_t c[3] = {0, 0, 0};
    int ns = 0;
#pragma unroll
    for (int k = 0; k < MG_ITEMS; k++) {
        if (base + k < A.m) {
            bool st;
            const unsigned f = migrate_flags(A, base + k, st);
            c[0] += f & 1u; c[1] += (f >> 1) & 1u; c[2] += (f >> 2) & 1u;
            ns += st ? 1 : 0;
        }
    }
    if (ns) atomicAdd(strayed, ns);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        uint32_t v = c[a];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) sm[a][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t t = 0;
        for (int k = 0; k < MG_THREADS / 32; k++) t += sm[threadIdx.x][k];
        blockCnt[(size_t)threadIdx.x * nb + blockIdx.x] = t;
    }
}
// exclusive scan of the three block-count rows in place, totals to tot[0..2] (one CTA: a row holds m / 1024 entries)
__global__ void __launch_bounds__(1024) migrate_scan_kernel(uint32_t* __restrict__ blockCnt, unsigned nb, int* __restrict__ tot) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = 0; a < 3; a++) {
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (unsigned base = 0; base < nb; base += 1024) {
            const unsigned i = base + threadIdx.x;
            const uint32_t v = i < nb ? blockCnt[(size_t)a * nb + i] : 0u;
            uint32_t incl = v;
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            if (lane == 31) ws[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                uint32_t t = ws[lane], ti = t;
                for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti += u; }
                ws[lane] = ti - t;
            }
            __syncthreads();
            const uint32_t excl = carry + ws[wid] + incl - v;
            if (i < nb) blockCnt[(size_t)a * nb + i] = excl;
            __syncthreads();
            if (threadIdx.x == 1023) carry = excl + v;
            __syncthreads();
        }
        if (threadIdx.x == 0) tot[a] = (int)carry;
        __syncthreads();
    }
}
struct CompactDst { uint32_t *w0, *w1, *w2; int3* ijk; };
__global__ void __launch_bounds__(MG_THREADS) migrate_scatter_kernel(MigrateArgs A, const uint32_t* __restrict__ w0, const uint32_t* __restrict__ w1,
                                                                     const uint32_t* __restrict__ w2, const uint32_t* __restrict__ blockOff, unsigned nb,
                                                                     CompactDst L, CompactDst R, CompactDst K) {
    __shared__ uint32_t sm[3][MG_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * MG_BLOCK + (uint64_t)threadIdx.x * MG_ITEMS;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned f[MG_ITEMS];
    uint32_t c[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < MG_ITEMS; k++) {
        f[k] = 0;
        if (base + k < A.m) { bool st; f[k] = migrate_flags(A, base + k, st); }
        c[0] += f[k] & 1u; c[1] += (f[k] >> 1) & 1u; c[2] += (f[k] >> 2) & 1u;
    }
    uint32_t pos[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {   // exclusive scan over the block's threads (items of a thread are consecutive: the order is kept)
        uint32_t incl = c[a];
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) sm[a][wid] = incl;
        pos[a] = incl - c[a];
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; a++) {
        uint32_t before = 0;
        for (unsigned k = 0; k < wid; k++) before += sm[a][k];
        pos[a] += before + blockOff[(size_t)a * nb + blockIdx.x];
    }
#pragma unroll
    for (int k = 0; k < MG_ITEMS; k++) {
        if (!f[k]) continue;
        const uint64_t g = A.pLo + base + k;
        const uint32_t a = w0[g], b = w1[g], c2 = w2[g];
        const int3 v = A.ijk[g];
        if (f[k] & 1u) { const uint32_t d = pos[0]++; L.w0[d] = a; L.w1[d] = b; L.w2[d] = c2; L.ijk[d] = v; }
        if (f[k] & 2u) { const uint32_t d = pos[1]++; R.w0[d] = a; R.w1[d] = b; R.w2[d] = c2; R.ijk[d] = v; }
        if (f[k] & 4u) { const uint32_t d = pos[2]++; K.w0[d] = a; K.w1[d] = b; K.w2[d] = c2; K.ijk[d] = v; }
    }
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
    // layer bounds and the neighbours' layer sizes in ONE read-back: the bounds stay on the device, the counts derived from
    // them travel device to device (round 1: a read-back of the bounds, then a host round trip for the counts)
    DBuf<int> d(24, w->stream);   // [0..5] xs, [6..11] b, [12..15] toLeft[2] toRight[2], [16..19] fromLeft[2] fromRight[2]
    d.zero();
    FB_CUDA(cudaMemcpyAsync(d.p, xs, sizeof(xs), cudaMemcpyHostToDevice, w->stream));
    FB_LAUNCH(w, "dd_layer_bounds", 64) layer_bounds_kernel<<<1, 32, 0, w->stream>>>(t.origin.p, t.n, d.p, d.p + 6);
    check_launch("layer_bounds");
    comm_group_begin(w);
    if (hasL) { comm_send(w, w->rank - 1, d.p + 12, 8); comm_recv(w, w->rank - 1, d.p + 16, 8); }
    if (hasR) { comm_send(w, w->rank + 1, d.p + 14, 8); comm_recv(w, w->rank + 1, d.p + 18, 8); }
    comm_group_end(w);
    int hb[14];
    read_back(w, hb, d.p + 6, sizeof(hb));
    for (int k = 0; k < 6; k++) M.b[k] = hb[k];
    const int toLeft[2] = {M.b[1] - M.b[0], M.b[2] - M.b[0]};
    const int toRight[2] = {M.b[5] - M.b[4], M.b[5] - M.b[3]};
    FB_REQUIRE(hb[6] == toLeft[0] && hb[7] == toLeft[1] && hb[8] == toRight[0] && hb[9] == toRight[1], FLIPB200_ERR_STATE, "dd_maps: layer counts");
    M.recvCnt[0][0] = hasL ? hb[10] : 0; M.recvCnt[0][1] = hasL ? hb[11] : 0;
    M.recvCnt[1][0] = hasR ? hb[12] : 0; M.recvCnt[1][1] = hasR ? hb[13] : 0;
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
    M.epoch = t.epoch;   // everything above is ordered on the world's stream; the temporaries are freed stream-ordered
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

namespace {
// collective, once per world: allocate my boxes, swap the allocation's handle (IPC) or address (in-process) with both neighbours
void peer_setup(World* w) {
    DDPeer& P = w->dd->peer;
    if (P.tried) return;
    P.tried = true;
    const char* e = getenv("FLIPB200_DD_P2P");
    int want = e ? atoi(e) : 1;
    const bool has[2] = {w->rank > 0, w->rank < w->nRanks - 1};
    const bool local = comm_is_local(w);
    size_t mb = 64;   // per box; four boxes per rank (a fused two-layer refresh of Velocity + PostAdvVelocity + LiquidSDF across a 32 x 32-leaf face is 30 MB)
    if (const char* b = getenv("FLIPB200_DD_P2P_MB")) mb = (size_t)std::max(1, atoi(b));
    struct Blob { unsigned char h[64]; unsigned long long ptr; int ok; int pad; };   // 80 bytes
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    Blob mine;
    memset(&mine, 0, sizeof(mine));
    if (want) {
        P.boxBytes = mb << 20;
        if (cudaMalloc((void**)&P.mine, P.total()) != cudaSuccess) { cudaGetLastError(); P.mine = nullptr; }
        if (P.mine) {
            FB_CUDA(cudaMemsetAsync(P.mine + 4 * P.boxBytes, 0, P.total() - 4 * P.boxBytes, w->stream));
            mine.ptr = (unsigned long long)(uintptr_t)P.mine;
            mine.ok = 1;
            if (!local) {
                cudaIpcMemHandle_t h;
                if (cudaIpcGetMemHandle(&h, P.mine) == cudaSuccess) memcpy(mine.h, &h, 64);
                else { cudaGetLastError(); mine.ok = 0; }
            }
        }
    }
    // every rank takes part in the swap, whatever it could allocate: the verdict must be the same on both sides of a face
    DBuf<unsigned char> d(3 * sizeof(Blob), w->stream);
    FB_CUDA(cudaMemsetAsync(d.p, 0, 3 * sizeof(Blob), w->stream));
    FB_CUDA(cudaMemcpyAsync(d.p, &mine, sizeof(Blob), cudaMemcpyHostToDevice, w->stream));
    comm_group_begin(w);
    if (has[0]) { comm_send(w, w->rank - 1, d.p, sizeof(Blob)); comm_recv(w, w->rank - 1, d.p + sizeof(Blob), sizeof(Blob)); }
    if (has[1]) { comm_send(w, w->rank + 1, d.p, sizeof(Blob)); comm_recv(w, w->rank + 1, d.p + 2 * sizeof(Blob), sizeof(Blob)); }
    comm_group_end(w);
    Blob theirs[2];
    read_back(w, theirs, d.p + sizeof(Blob), 2 * sizeof(Blob));
    bool ok = mine.ok != 0;
    for (int s = 0; s < 2 && ok; s++) {
        if (!has[s]) continue;
        if (!theirs[s].ok) { ok = false; break; }
        if (local) P.remote[s] = reinterpret_cast<char*>((uintptr_t)theirs[s].ptr);
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, theirs[s].h, 64);
            void* rp = nullptr;
            if (cudaIpcOpenMemHandle(&rp, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            P.remote[s] = (char*)rp; P.opened[s] = true;
        }
    }
    // both ends of every face must agree: all-reduce the verdict (a rank whose neighbour failed must not push into nothing)
    DBuf<int> v(1, w->stream);
    const int mineOk = ok ? 0 : 1;
    FB_CUDA(cudaMemcpyAsync(v.p, &mineOk, 4, cudaMemcpyHostToDevice, w->stream));
    comm_allreduce(w, v.p, 1, CT_I32, false);
    int bad = 0;
    read_back(w, &bad, v.p, 4);
    P.counter.alloc(2, w->stream);
    P.counter.zero();
    P.errHost = reinterpret_cast<int*>(w->hostScratch + 800);
    *P.errHost = 0;
    P.ready = bad == 0;
}
}  // namespace

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
    peer_setup(w);
    DDPeer& P = w->dd->peer;
    DBuf<char> stage[2];
    std::vector<size_t> off[2], soff[2];   // layout of what I receive from side s / of what I send to it (the neighbour's layout)
    size_t rbytes[2] = {0, 0}, sbytes[2] = {0, 0};
    uint64_t bytes = 0;
    bool viaPeer[2] = {false, false};
    for (int s = 0; s < 2; s++) {
        size_t cur = 0, scur = 0;
        for (auto& a : arrays) {
            off[s].push_back(cur); cur += (((size_t)recvCnt[s] * a.bytesPerLeaf) + 255) & ~(size_t)255;
            soff[s].push_back(scur); scur += (((size_t)sendCnt[s] * a.bytesPerLeaf) + 255) & ~(size_t)255;
        }
        rbytes[s] = cur; sbytes[s] = scur;
        // both ends of a face see the same two sizes (my send is its receive), so they take the same route
        viaPeer[s] = has[s] && P.ready && std::max(cur, scur) <= P.boxBytes;
        if (!viaPeer[s]) stage[s].alloc(cur + 256, w->stream);
        if (has[s]) bytes += cur;
    }
    // peer route: push my boundary leaves into the neighbour's box, then wait for its push into mine
    for (int s = 0; s < 2; s++) {
        if (!viaPeer[s]) continue;
        const unsigned ep = ++P.epoch[s];
        const int par = (int)(ep & 1u);
        PushArgs pa;
        pa.nArr = (int)arrays.size();
        size_t most = 0;
        for (size_t k = 0; k < arrays.size(); k++) {
            pa.src[k] = (const char*)arrays[k].base + (size_t)sendStart[s] * arrays[k].bytesPerLeaf;
            pa.bytes[k] = (size_t)sendCnt[s] * arrays[k].bytesPerLeaf;
            pa.off[k] = soff[s][k];
            most = std::max(most, pa.bytes[k]);
        }
        // the neighbour's box for ITS side facing me: I am its right neighbour when it is my left one
        char* rbox = P.box(P.remote[s], 1 - s, par);
        unsigned* rflag = P.flag(P.remote[s], 1 - s, par);
        const unsigned grid = (unsigned)std::min<size_t>(148, std::max<size_t>(1, (most / 64 + 255) / 256));   // <= one CTA per SM, 64 bytes per thread and round
        FB_LAUNCH(w, "dd_push", sbytes[s]) dd_push_kernel<<<grid, 256, 0, w->stream>>>(pa, rbox, rflag, ep, P.counter.p + s);
        check_launch("dd_push");
    }
    bool anyNccl = false;
    for (int s = 0; s < 2; s++) anyNccl = anyNccl || (has[s] && !viaPeer[s]);
    if (anyNccl) {
        comm_group_begin(w);
        for (int s = 0; s < 2; s++) {
            if (!has[s] || viaPeer[s]) continue;
            for (size_t k = 0; k < arrays.size(); k++) {
                const auto& a = arrays[k];
                comm_send(w, peer[s], (const char*)a.base + (size_t)sendStart[s] * a.bytesPerLeaf, (size_t)sendCnt[s] * a.bytesPerLeaf);
                comm_recv(w, peer[s], stage[s].p + off[s][k], (size_t)recvCnt[s] * a.bytesPerLeaf);
            }
        }
        comm_group_end(w);
    }
    // In-process worlds share ONE device: a kernel that spins for a neighbour's push can sit in front of that very push in a
    // hardware work queue two streams happen to share (measured: 3 worlds deadlock until the time limit). There the wait is a
    // host rendezvous -- a 4-byte message each way, which in that backend synchronises both streams -- and the device-side wait
    // is exercised where it belongs, between GPUs (tests/test_nccl_gpu.py, bench.py --gpus N).
    const bool hostWait = comm_is_local(w);
    if (hostWait && (viaPeer[0] || viaPeer[1])) {
        DBuf<int> token(4, w->stream);
        token.zero();
        comm_group_begin(w);
        for (int s = 0; s < 2; s++) if (viaPeer[s]) { comm_send(w, peer[s], token.p + s, 4); comm_recv(w, peer[s], token.p + 2 + s, 4); }
        comm_group_end(w);
    }
    for (int s = 0; s < 2; s++) {
        if (!has[s]) continue;
        const char* src = stage[s].p;
        if (viaPeer[s]) {
            const int par = (int)(P.epoch[s] & 1u);
            src = P.box(P.mine, s, par);
            if (!hostWait) {
                FB_LAUNCH(w, "dd_wait", 4) dd_wait_kernel<<<1, 1, 0, w->stream>>>(P.flag(P.mine, s, par), P.epoch[s], reinterpret_cast<int*>(w->hostScratchDev + 800));
                check_launch("dd_wait");
            }
        }
        if (recvCnt[s] == 0) continue;
        UnpackArgs u;
        u.nArr = (int)arrays.size();
        for (size_t k = 0; k < arrays.size(); k++) { u.dst[k] = (char*)arrays[k].base; u.src[k] = src + off[s][k]; u.bpl[k] = arrays[k].bytesPerLeaf; }
        FB_LAUNCH(w, "dd_unpack", 2 * bytes) unpack_kernel<<<dim3(recvCnt[s], u.nArr), 128, 0, w->stream>>>(u, M.map[s].p + mapOff[s]);
        check_launch("dd_unpack");
    }
    if (P.ready && P.errHost && *P.errHost) {
        *P.errHost = 0;
        throw Error(FLIPB200_ERR_COMM, "ghost exchange: a neighbour's push did not arrive within the time limit");
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
    check_bounds(w);
    DDState& D = *w->dd;
    const unsigned nb = (unsigned)std::max<uint64_t>(1, (m + MG_BLOCK - 1) / MG_BLOCK);
    DBuf<uint32_t> blockCnt((size_t)3 * nb, w->stream);
    DBuf<int> tot(8, w->stream);   // [0..2] to left / to right / kept, [3] strayed, [4] from left, [5] from right
    blockCnt.zero(); tot.zero();
    MigrateArgs A{ijk, alive, pLo, m, D.lo, D.hi, hasL ? 1 : 0, hasR ? 1 : 0, D.leftLo, D.rightHi};
    if (m) {
        FB_LAUNCH(w, "dd_migrate_count", m * 13) migrate_count_kernel<<<nb, MG_THREADS, 0, w->stream>>>(A, blockCnt.p, nb, tot.p + 3);
        check_launch("migrate_count");
    }
    FB_LAUNCH(w, "dd_migrate_scan", (size_t)nb * 24) migrate_scan_kernel<<<1, 1024, 0, w->stream>>>(blockCnt.p, nb, tot.p);
    check_launch("migrate_scan");
    FB_PHASE(w, "dd_migrate after counts");
    // the counts travel device to device; one read-back brings mine, the neighbours' and the (all-reduced) stray verdict
    comm_allreduce(w, tot.p + 3, 1, CT_I32, false);     // every rank must see the same verdict (a lone throw would hang the others)
    comm_group_begin(w);
    if (hasL) { comm_send(w, w->rank - 1, tot.p + 0, 4); comm_recv(w, w->rank - 1, tot.p + 4, 4); }
    if (hasR) { comm_send(w, w->rank + 1, tot.p + 1, 4); comm_recv(w, w->rank + 1, tot.p + 5, 4); }
    comm_group_end(w);
    int h[6] = {0, 0, 0, 0, 0, 0};
    read_back(w, h, tot.p, sizeof(h));
    const uint64_t cL = (uint64_t)h[0], cR = (uint64_t)h[1], cK = (uint64_t)h[2];
    const int strayed = h[3];
    const int fromLeft[2] = {h[4], 0}, fromRight[2] = {h[5], 0};
    FB_REQUIRE(strayed == 0, FLIPB200_ERR_DOMAIN, std::to_string(strayed) + " particles moved past a neighbour's whole slab in one step (migration is one hop): use thicker slabs or a smaller time step");
    const uint64_t rL = hasL ? (uint64_t)fromLeft[0] : 0, rR = hasR ? (uint64_t)fromRight[0] : 0;
    const uint64_t n = rL + cK + rR;
    o0.alloc(n + 1, w->stream); o1.alloc(n + 1, w->stream); o2.alloc(n + 1, w->stream); oijk.alloc(n + 1, w->stream);
    DBuf<uint32_t> sL0(cL + 1, w->stream), sL1(cL + 1, w->stream), sL2(cL + 1, w->stream), sR0(cR + 1, w->stream), sR1(cR + 1, w->stream), sR2(cR + 1, w->stream);
    DBuf<int3> sLi(cL + 1, w->stream), sRi(cR + 1, w->stream);
    if (m) {
        CompactDst L{sL0.p, sL1.p, sL2.p, sLi.p}, R{sR0.p, sR1.p, sR2.p, sRi.p}, K{o0.p + rL, o1.p + rL, o2.p + rL, oijk.p + rL};
        FB_LAUNCH(w, "dd_migrate_scatter", m * 49) migrate_scatter_kernel<<<nb, MG_THREADS, 0, w->stream>>>(A, w0, w1, w2, blockCnt.p, nb, L, R, K);
        check_launch("migrate_scatter");
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
