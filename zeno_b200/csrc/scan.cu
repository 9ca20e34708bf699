// libflipb200 -- device-wide exclusive prefix sum (u32), three-phase and deterministic.
// Used for leaf-slot numbering (dense directory -> slots) and for the particle store's
// per-voxel offsets (K1/K2: the histogram -> offsets step of the counting sort).
#include "world.cuh"

namespace fb {
namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// phase 1/3: per-tile local exclusive scan; tile totals to sums[]
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const uint32_t* __restrict__ in,
                                                           uint32_t* __restrict__ out,
                                                           uint32_t* __restrict__ sums, size_t n) {
    __shared__ uint32_t warpTotals[SCAN_THREADS / 32];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        local += v[i];
    }
    uint32_t incl = warp_incl_scan(local);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 31) warpTotals[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = lane < SCAN_THREADS / 32 ? warpTotals[lane] : 0u;
        uint32_t ti = warp_incl_scan(t);
        if (lane < SCAN_THREADS / 32) warpTotals[lane] = ti - t;
        if (lane == SCAN_THREADS / 32 - 1 && sums) sums[blockIdx.x] = ti;
    }
    __syncthreads();
    uint32_t run = warpTotals[wid] + incl - local;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_add(uint32_t* __restrict__ out,
                                                         const uint32_t* __restrict__ sums, size_t n) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t add = sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] += add;
}

void scan_rec(World* w, const uint32_t* in, uint32_t* out, size_t n, uint32_t* totalDev) {
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles <= 1) {
        FB_LAUNCH(w, "scan_tiles", n * 8) scan_tiles<<<1, SCAN_THREADS, 0, w->stream>>>(in, out, totalDev, n);
        check_launch("scan_tiles");
        return;
    }
    DBuf<uint32_t> sums(tiles, w->stream);
    FB_LAUNCH(w, "scan_tiles", n * 8) scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, w->stream>>>(in, out, sums.p, n);
    check_launch("scan_tiles");
    scan_rec(w, sums.p, sums.p, tiles, totalDev);
    FB_LAUNCH(w, "scan_add", n * 8) scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, w->stream>>>(out, sums.p, n);
    check_launch("scan_add");
}
}  // namespace

void exclusive_scan_u32(World* w, const uint32_t* in, uint32_t* out, size_t n, uint64_t* total) {
    if (n == 0) { if (total) *total = 0; return; }
    DBuf<uint32_t> tot(1, w->stream);
    scan_rec(w, in, out, n, tot.p);
    if (total) {
        uint32_t h = 0;
        read_back(w, &h, tot.p, sizeof(uint32_t));
        *total = h;
    }
}

}  // namespace fb
