// libflipb200 -- particle-to-grid transfer (K3/K4), velocity extrapolation (K5).
//
// p2g_gather_kernel: one CTA per pool leaf, "collect style" like the reference
// (FF/FLIP_vdb.cpp:1137-1263): every target voxel gathers the particles of its 27
// neighbouring cells. The 10x10x10 cell shell around the leaf is streamed through shared
// memory one x-plane at a time (decoded once per CTA), the per-voxel accumulators
// (sum w*v, sum w per channel, min d^2, count) live in shared memory, and each voxel visits
// its source cells in (x,y,z) order and the particles of a cell in store order. That is
// exactly the order of the reference's all_particle_iterator (:1021-1132), so the fp32 sums
// are reproducible and there are no float atomics anywhere.
#include "world.cuh"

namespace fb {
void mark_alloc_from_mask(World* w, const uint64_t* mask, int n, uint8_t* alloc);
namespace {

constexpr int P2G_THREADS = 192;   // three x-slices of 64 target voxels are live per plane
constexpr int PLANE_CELLS = 100;   // 10 x 10 cells (y,z in [-1,8])
constexpr int PLANE_CAP = 1400;    // staged particles per batch (a plane at 8 ppc holds ~800)

struct P2GParams {
    TopoView t;
    const uint32_t* voxelStart;
    const uint32_t *w0, *w1, *w2;
    float* vel[3];        // out: normalised velocity, 0 where the channel is off
    uint64_t* chMask;     // out: [3][n][8]
    float* sdf;           // out: liquid sdf values
    uint64_t* topoMask;   // out: dilated-occupancy mask [n][8]
    float dx, radius, sdfBg;
    int* overflow;      // [0] a cell beyond the gather kernel's staging depth; [2..3] / [4..5] 64-bit sums of (records per plane)^2 and
                        // of records per plane over all leaves: their ratio sizes the next call's staging buffer
    int cap;            // p2g_xrow_kernel: staged records per batch (dynamic shared memory = cap * 25 bytes)
};

__global__ void __launch_bounds__(P2G_THREADS) p2g_gather_kernel(P2GParams p, const uint8_t* __restrict__ todo) {
    extern __shared__ float smem[];
    if (todo && !todo[blockIdx.x]) return;   // fallback pass: only the leaves p2g_xrow_kernel handed back
    float* sP = smem;                          // [PLANE_CAP][6] px,py,pz,vx,vy,vz
    float* acc = sP + PLANE_CAP * 6;           // [8][512]: wv0..2, w0..2, mind2, count
    __shared__ uint32_t cBeg[PLANE_CELLS];
    __shared__ uint32_t cCnt[PLANE_CELLS];
    __shared__ int cOff[PLANE_CELLS + 1];      // smem offset of the cell in the current batch (-1: not staged)
    __shared__ int sBatchEnd, sBatchTotal;

    const int leaf = blockIdx.x;
    const int tid = threadIdx.x;
    const int* nbr = p.t.nbr27 + (size_t)leaf * 27;

    for (int i = tid; i < 8 * LEAF; i += P2G_THREADS) acc[i] = (i >= 6 * LEAF && i < 7 * LEAF) ? 3.0e38f : 0.f;

    for (int cx = -1; cx <= 8; cx++) {
        __syncthreads();
        // particle ranges of the plane's 100 cells
        if (tid < PLANE_CELLS) {
            int cy = tid / 10 - 1, cz = tid % 10 - 1;
            int li = (cx < 0 ? 0 : (cx < 8 ? 1 : 2)) * 9 + (cy < 0 ? 0 : (cy < 8 ? 1 : 2)) * 3 + (cz < 0 ? 0 : (cz < 8 ? 1 : 2));
            int nl = nbr[li];
            uint32_t b = 0, c = 0;
            if (nl >= 0) {
                size_t v = (size_t)nl * LEAF + (((cx & 7) << 6) | ((cy & 7) << 3) | (cz & 7));
                b = __ldg(&p.voxelStart[v]);
                c = __ldg(&p.voxelStart[v + 1]) - b;
            }
            cBeg[tid] = b; cCnt[tid] = c;
        }
        __syncthreads();
        int batchStart = 0;
        while (batchStart < PLANE_CELLS) {
            // warp 0: the maximal run of cells starting at batchStart whose particles fit PLANE_CAP
            if (tid < 32) {
                int run = 0, end = PLANE_CELLS;
                for (int base = batchStart; base < PLANE_CELLS; base += 32) {
                    int c = base + tid;
                    int cnt = 0;
                    if (c < PLANE_CELLS) {
                        cnt = (int)cCnt[c];
                        if (cnt > PLANE_CAP) { cnt = PLANE_CAP; atomicExch(p.overflow, 1); }
                    }
                    int incl = cnt;
                    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
                    int excl = run + incl - cnt;
                    bool fits = excl + cnt <= PLANE_CAP;  // monotone in c: a prefix property
                    unsigned fm = __ballot_sync(0xffffffffu, fits);
                    int nfit = fm == 0xffffffffu ? 32 : __ffs(~fm) - 1;
                    if (tid < nfit && c < PLANE_CELLS) cOff[c] = excl;
                    if (nfit < 32) { run = __shfl_sync(0xffffffffu, excl, nfit); end = base + nfit; break; }   // run = particles of the cells that fit
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (end > PLANE_CELLS) end = PLANE_CELLS;
                if (tid == 0) { sBatchEnd = end; sBatchTotal = run; }
            }
            __syncthreads();
            const int batchEnd = sBatchEnd;
            const int total = sBatchTotal;
            // stage + decode the batch's particles, one thread per staged slot (round 1 used one thread per cell with a serial loop over
            // its particles: 100 of 192 threads busy, strided loads; ncu showed a third of all warp samples waiting at the barriers).
            // The slot's cell is the last staged cell c with cOff[c] <= slot; consecutive threads read consecutive particles of a cell.
            for (int sl = tid; sl < total; sl += P2G_THREADS) {
                int lo = batchStart, hi = batchEnd;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cOff[mid] <= sl) lo = mid; else hi = mid; }
                const uint32_t gi = cBeg[lo] + (uint32_t)(sl - cOff[lo]);
                const uint32_t a0 = __ldg(&p.w0[gi]), a1 = __ldg(&p.w1[gi]), a2 = __ldg(&p.w2[gi]);
                float* dst = sP + (size_t)sl * 6;
                dst[0] = fx_decode(a0 & 0xffffu);
                dst[1] = fx_decode(a0 >> 16);
                dst[2] = fx_decode(a1 & 0xffffu);
                dst[3] = h_decode(a1 >> 16);
                dst[4] = h_decode(a2 & 0xffffu);
                dst[5] = h_decode(a2 >> 16);
            }
            __syncthreads();
            // accumulate: slice si handles target x = cx + 1 - si... (ox = source - target)
            {
                int si = tid >> 6;                 // 0,1,2
                int x = cx - 1 + si;               // si=0 -> ox=+1, si=1 -> ox=0, si=2 -> ox=-1
                int ox = cx - x;
                int y = (tid >> 3) & 7, z = tid & 7;
                if (x >= 0 && x <= 7) {
                    int vo = (x << 6) | (y << 3) | z;
                    float a0 = acc[vo], a1 = acc[LEAF + vo], a2 = acc[2 * LEAF + vo];
                    float g0 = acc[3 * LEAF + vo], g1 = acc[4 * LEAF + vo], g2 = acc[5 * LEAF + vo];
                    float md = acc[6 * LEAF + vo], cn = acc[7 * LEAF + vo];
                    const float fx = (float)(-ox);
                    for (int oy = -1; oy <= 1; oy++) {
                        const float fy = (float)(-oy);
                        for (int oz = -1; oz <= 1; oz++) {
                            const float fz = (float)(-oz);
                            int c = (y + oy + 1) * 10 + (z + oz + 1);
                            if (c < batchStart || c >= batchEnd) continue;
                            int cnt = min((int)cCnt[c], PLANE_CAP);
                            const float* src = sP + (size_t)cOff[c] * 6;
                            for (int j = 0; j < cnt; j++) {
                                float px = src[6 * j], py = src[6 * j + 1], pz = src[6 * j + 2];
                                float tx = fabsf(__fsub_rn(fx, px)), ty = fabsf(__fsub_rn(fy, py)), tz = fabsf(__fsub_rn(fz, pz));
                                float d2 = __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz));
                                md = fminf(md, d2);
                                cn += 1.f;
                                float hx = fmaxf(0.f, __fsub_rn(1.0f, tx)), hy = fmaxf(0.f, __fsub_rn(1.0f, ty)), hz = fmaxf(0.f, __fsub_rn(1.0f, tz));
                                if (ox != 1) {  // u sample at -0.5 in x; a source cell at +1 is always out of reach
                                    float xs = fabsf(__fsub_rn(__fadd_rn(fx, -0.5f), px));
                                    float wgt = __fmul_rn(__fmul_rn(fmaxf(0.f, __fsub_rn(1.0f, xs)), hy), hz);
                                    a0 = __fadd_rn(__fmul_rn(src[6 * j + 3], wgt), a0);
                                    g0 = __fadd_rn(wgt, g0);
                                }
                                if (oy != 1) {
                                    float ys = fabsf(__fsub_rn(__fadd_rn(fy, -0.5f), py));
                                    float wgt = __fmul_rn(__fmul_rn(hx, fmaxf(0.f, __fsub_rn(1.0f, ys))), hz);
                                    a1 = __fadd_rn(__fmul_rn(src[6 * j + 4], wgt), a1);
                                    g1 = __fadd_rn(wgt, g1);
                                }
                                if (oz != 1) {
                                    float zs = fabsf(__fsub_rn(__fadd_rn(fz, -0.5f), pz));
                                    float wgt = __fmul_rn(__fmul_rn(hx, hy), fmaxf(0.f, __fsub_rn(1.0f, zs)));
                                    a2 = __fadd_rn(__fmul_rn(src[6 * j + 5], wgt), a2);
                                    g2 = __fadd_rn(wgt, g2);
                                }
                            }
                        }
                    }
                    acc[vo] = a0; acc[LEAF + vo] = a1; acc[2 * LEAF + vo] = a2;
                    acc[3 * LEAF + vo] = g0; acc[4 * LEAF + vo] = g1; acc[5 * LEAF + vo] = g2;
                    acc[6 * LEAF + vo] = md; acc[7 * LEAF + vo] = cn;
                }
            }
            __syncthreads();
            batchStart = batchEnd;
        }
    }
    __syncthreads();
    // normalize_p2g_velocity (FF/FLIP_vdb.cpp:120-165) + sdf transform (:1199-1204)
    for (int base = 0; base < LEAF; base += 32 * (P2G_THREADS / 32)) {
        int vo = base + tid;
        bool valid = vo < LEAF;
        bool on[3] = {false, false, false};
        bool touched = false;
        if (valid) {
            touched = acc[7 * LEAF + vo] > 0.f;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float wsum = acc[(3 + c) * LEAF + vo];
                float out = 0.f;
                if (touched && wsum != 0.f) { out = __fdiv_rn(acc[c * LEAF + vo], __fadd_rn(wsum, 0.001f)); on[c] = true; }
                p.vel[c][(size_t)leaf * LEAF + vo] = out;
            }
            float s = p.sdfBg;
            if (touched) s = fminf(s, __fsub_rn(__fmul_rn(p.dx, sqrtf(acc[6 * LEAF + vo])), p.radius));
            p.sdf[(size_t)leaf * LEAF + vo] = s;
        }
        unsigned bt = __ballot_sync(0xffffffffu, valid && touched);
        unsigned b0 = __ballot_sync(0xffffffffu, on[0]), b1 = __ballot_sync(0xffffffffu, on[1]), b2 = __ballot_sync(0xffffffffu, on[2]);
        if (valid && (tid & 31) == 0) {
            size_t wi = (size_t)leaf * 16 + (vo >> 5);
            reinterpret_cast<uint32_t*>(p.topoMask)[wi] = bt;
            size_t nW = (size_t)p.t.n * 16;
            reinterpret_cast<uint32_t*>(p.chMask)[wi] = b0;
            reinterpret_cast<uint32_t*>(p.chMask)[nW + wi] = b1;
            reinterpret_cast<uint32_t*>(p.chMask)[2 * nW + wi] = b2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// p2g_xrow_kernel (round 2): the same collect-style sums in the same order, a third of the instructions.
// One CTA per pool leaf, 64 threads; thread (y, z) owns the whole x-row of eight target voxels. The shell is still
// streamed one x-plane at a time, but a staged particle now serves the THREE targets x = cx+1, cx, cx-1 of its thread
// at once: everything that depends on y and z only (two of the three distances, hats, staggered hats) is computed
// once per particle instead of once per (particle, target), a particle record is loaded once instead of three times,
// and the accumulators of the three live targets stay in registers (they rotate as the plane advances; a target is
// finished after the plane at its ox = +1 and written straight to global memory -- no accumulator array in shared
// memory). Each target still sees its 27 source cells in (ox, oy, oz) order and a cell's particles in store order, and
// every product / sum is the same single-rounded operation as before, so the result is bit-identical to
// p2g_gather_kernel (tests/test_p2g_paths_gpu.py) and to the oracle.
// Staging: a plane's records are decoded once per CTA into 24-byte records [px py pz vx vy vz], cell after cell in
// (y, z) order with ONE PAD RECORD per cell: with exactly 8 particles in every voxel (the state a scene starts in)
// un-padded cell starts are 48 words apart and the 32 lanes of a warp would hit two bank pairs.
// Planes that do not fit the staging buffer are processed in batches of whole cell rows (ascending y keeps a target's
// oy order); a single row that does not fit hands the leaf back to p2g_gather_kernel through `todo`.
constexpr int XR_THREADS = 64;
#ifndef FB_XR_CAP
#define FB_XR_CAP 1024   // 8 CTAs per SM; measured 1178 us against 1304 us at 1344 (6 CTAs) and 1458 us at 1600 (5 CTAs)
#endif
#ifndef FB_XR_CAP_MAX
#define FB_XR_CAP_MAX 2304   // 4 CTAs per SM; beyond that a plane is split into row batches instead
#endif
#ifndef FB_XR_UNROLL
#define FB_XR_UNROLL 8
#endif
#ifndef FB_XR_JUNROLL
#define FB_XR_JUNROLL 2
#endif
constexpr int XR_CAP = FB_XR_CAP;   // staged records per batch, pad records included: the default, at 8 particles per voxel
constexpr int XR_CAP_MAX = FB_XR_CAP_MAX;
constexpr int XR_JUNROLL = FB_XR_JUNROLL;
constexpr int XR_MIN_CTAS = 8;
constexpr int XR_STAGE_UNROLL = FB_XR_UNROLL;

struct XAcc { float a0, a1, a2, g0, g1, g2, md; int cn; };
__device__ __forceinline__ void xacc_reset(XAcc& A) { A.a0 = A.a1 = A.a2 = A.g0 = A.g1 = A.g2 = 0.f; A.md = 3.0e38f; A.cn = 0; }

// one (particle, target) pair; OX = source cell - target voxel in x
template <int OX>
__device__ __forceinline__ void xr_pair(XAcc& A, float px, float vx, float vy, float vz, float ty2, float tz2, float hy, float hz,
                                        float ysh, float zsh) {
    const float fx = (float)(-OX);
    const float tx = fabsf(__fsub_rn(fx, px));
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), ty2), tz2);
    A.md = fminf(A.md, d2);
    const float hx = fmaxf(0.f, __fsub_rn(1.0f, tx));
    if (OX != 1) {   // u sample at -0.5 in x; a source cell at +1 is always out of reach
        const float xs = fabsf(__fsub_rn(fx - 0.5f, px));
        const float wgt = __fmul_rn(__fmul_rn(fmaxf(0.f, __fsub_rn(1.0f, xs)), hy), hz);
        A.a0 = __fadd_rn(__fmul_rn(vx, wgt), A.a0);
        A.g0 = __fadd_rn(wgt, A.g0);
    }
    {   // v / w from a source cell at oy / oz = +1: the staggered hat is exactly 0 there (|-1.5 - p| >= 1 for every decodable p),
        // the products are +-0 and the sums do not change -- the same bits as skipping the term, without a branch per cell offset
        const float wgt = __fmul_rn(__fmul_rn(hx, ysh), hz);
        A.a1 = __fadd_rn(__fmul_rn(vy, wgt), A.a1);
        A.g1 = __fadd_rn(wgt, A.g1);
    }
    {
        const float wgt = __fmul_rn(__fmul_rn(hx, hy), zsh);
        A.a2 = __fadd_rn(__fmul_rn(vz, wgt), A.a2);
        A.g2 = __fadd_rn(wgt, A.g2);
    }
}

// the particles of one source cell against the thread's three live targets (M: x = cx+1, Z: x = cx, P: x = cx-1).
// ONE loop body serves all nine (oy, oz) offsets (fy = -oy, fz = -oz at run time): nine specialised copies (x 2 variants) were
// 64 KB of code and the kernel spent 29 % of its warp samples waiting for instructions (ncu: stall_no_inst).
template <bool ALL>
__device__ __forceinline__ void xr_cell(const float* __restrict__ rec, int cnt, float fy, float fz, XAcc& M, XAcc& Z, XAcc& P, bool vM, bool vZ, bool vP) {
    const float2* __restrict__ src = reinterpret_cast<const float2*>(rec);
    const float fyh = fy - 0.5f, fzh = fz - 0.5f;
    if (ALL || vM) M.cn += cnt;
    if (ALL || vZ) Z.cn += cnt;
    if (ALL || vP) P.cn += cnt;
    // the record of the next visit is loaded while this one is evaluated; past the last particle that is the cell's pad record
    float2 n0 = src[0], n1 = src[1], n2 = src[2];
#pragma unroll XR_JUNROLL
    for (int j = 0; j < cnt; j++) {
        const float2 q0 = n0, q1 = n1, q2 = n2;
        n0 = src[3 * j + 3]; n1 = src[3 * j + 4]; n2 = src[3 * j + 5];
        const float px = q0.x, py = q0.y, pz = q1.x, vx = q1.y, vy = q2.x, vz = q2.y;
        const float ty = fabsf(__fsub_rn(fy, py)), tz = fabsf(__fsub_rn(fz, pz));
        const float ty2 = __fmul_rn(ty, ty), tz2 = __fmul_rn(tz, tz);
        const float hy = fmaxf(0.f, __fsub_rn(1.0f, ty)), hz = fmaxf(0.f, __fsub_rn(1.0f, tz));
        const float ysh = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(fyh, py))));
        const float zsh = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(fzh, pz))));
        if (ALL || vM) xr_pair<-1>(M, px, vx, vy, vz, ty2, tz2, hy, hz, ysh, zsh);
        if (ALL || vZ) xr_pair<0>(Z, px, vx, vy, vz, ty2, tz2, hy, hz, ysh, zsh);
        if (ALL || vP) xr_pair<1>(P, px, vx, vy, vz, ty2, tz2, hy, hz, ysh, zsh);
    }
}

template <bool ALL>
__device__ __forceinline__ void xr_rows(const float* __restrict__ sP, const int* __restrict__ cPre, const int* __restrict__ cCnt, int base,
                                        int r0, int r1, int y, int z, XAcc& M, XAcc& Z, XAcc& P, bool vM, bool vZ, bool vP) {
#pragma unroll 1
    for (int k = 0; k < 9; k++) {   // (oy, oz) in lexicographic order
        const int oy = k / 3 - 1, oz = k - (k / 3) * 3 - 1;
        const int row = y + oy + 1;
        if (row < r0 || row >= r1) continue;   // only a plane split into row batches diverges here
        const int c = row * 10 + (z + oz + 1);
        xr_cell<ALL>(sP + (size_t)(cPre[c] - base) * 6, cCnt[c], (float)(-oy), (float)(-oz), M, Z, P, vM, vZ, vP);
    }
}

// voxelStart range of plane cell c (c = (cy+1)*10 + cz+1) at source plane cx
__device__ __forceinline__ void xr_cell_range(const P2GParams& p, const int* __restrict__ sNbr, int cx, int c, uint32_t& b, uint32_t& n) {
    const int cy = c / 10 - 1, cz = c % 10 - 1;
    const int li = (cx < 0 ? 0 : (cx < 8 ? 1 : 2)) * 9 + (cy < 0 ? 0 : (cy < 8 ? 1 : 2)) * 3 + (cz < 0 ? 0 : (cz < 8 ? 1 : 2));
    const int nl = sNbr[li];
    b = 0; n = 0;
    if (nl >= 0 && c < PLANE_CELLS) {
        const size_t v = (size_t)nl * LEAF + (((cx & 7) << 6) | ((cy & 7) << 3) | (cz & 7));
        b = __ldg(&p.voxelStart[v]);
        n = __ldg(&p.voxelStart[v + 1]) - b;
    }
}

__global__ void __launch_bounds__(XR_THREADS, XR_MIN_CTAS) p2g_xrow_kernel(P2GParams p, uint8_t* __restrict__ todo) {
    extern __shared__ __align__(16) float xrSmem[];
    float* sP = xrSmem;                                              // [cap][6]
    uint8_t* cellOf = reinterpret_cast<uint8_t*>(xrSmem + (size_t)p.cap * 6);   // [cap]
    const int cap = p.cap;
    unsigned long long planeSq = 0, planeSum = 0;
    __shared__ uint32_t cBeg[PLANE_CELLS];
    __shared__ int cCnt[PLANE_CELLS];
    __shared__ int cPre[PLANE_CELLS + 1];   // first record of the cell, counted over the whole plane (pads included)
    __shared__ int sRow[12];                // batch b = cell rows [sRow[b], sRow[b+1])
    __shared__ int sNB;
    __shared__ int sNbr[27];

    const int leaf = blockIdx.x;
    const int tid = threadIdx.x;
    const int y = tid >> 3, z = tid & 7;
    if (tid < 27) sNbr[tid] = p.t.nbr27[(size_t)leaf * 27 + tid];
    __syncthreads();
    XAcc M, Z, P;
    xacc_reset(M); xacc_reset(Z); xacc_reset(P);
    // the cell ranges of a plane are fetched one plane ahead (two cells per thread), so that the plane's staging starts
    // with the particle loads instead of a dependent voxelStart round trip
    uint32_t nb0, nn0, nb1, nn1;
    xr_cell_range(p, sNbr, -1, tid, nb0, nn0);
    xr_cell_range(p, sNbr, -1, tid + XR_THREADS, nb1, nn1);

    for (int cx = -1; cx <= 8; cx++) {
        __syncthreads();   // the previous plane's readers are done
        cBeg[tid] = nb0; cCnt[tid] = (int)nn0;
        if (tid + XR_THREADS < PLANE_CELLS) { cBeg[tid + XR_THREADS] = nb1; cCnt[tid + XR_THREADS] = (int)nn1; }
        if (cx < 8) {
            xr_cell_range(p, sNbr, cx + 1, tid, nb0, nn0);
            xr_cell_range(p, sNbr, cx + 1, tid + XR_THREADS, nb1, nn1);
        }
        __syncthreads();
        if (tid < 32) {
            int run = 0;
            for (int base = 0; base < PLANE_CELLS; base += 32) {
                int c = base + tid;
                int v = c < PLANE_CELLS ? cCnt[c] + 1 : 0;
                int incl = v;
                for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += u; }
                if (c < PLANE_CELLS) cPre[c] = run + incl - v;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (tid == 0) cPre[PLANE_CELLS] = run;
            __syncwarp();
            if (tid == 0) {
                if (run > PLANE_CELLS) { planeSq += (unsigned long long)run * (unsigned)run; planeSum += (unsigned)run; }
                int nb = 0, r = 0;
                sRow[0] = 0;
                if (run > PLANE_CELLS) {   // a plane without particles has no batches
                    while (r < 10) {
                        int start = cPre[r * 10], e = r;
                        while (e < 10 && cPre[(e + 1) * 10] - start <= cap) e++;
                        if (e == r) { nb = -1; break; }
                        r = e; sRow[++nb] = r;
                    }
                }
                sNB = nb;
            }
        }
        __syncthreads();
        const int nB = sNB;
        if (nB < 0) { if (tid == 0) todo[leaf] = 1; return; }
        const bool vM = cx <= 6, vZ = cx >= 0 && cx <= 7, vP = cx >= 1;
        for (int b = 0; b < nB; b++) {
            const int r0 = sRow[b], r1 = sRow[b + 1];
            const int base = cPre[r0 * 10];
            const int total = cPre[r1 * 10] - base;
            if (b > 0) __syncthreads();
            // slot -> cell map: thread c marks the records of cell c (pad included)
            for (int c = r0 * 10 + tid; c < r1 * 10; c += XR_THREADS) {
                const int s0 = cPre[c] - base, n = cCnt[c];
                for (int j = 0; j <= n; j++) cellOf[s0 + j] = (uint8_t)c;
            }
            __syncthreads();
            for (int s0 = 0; s0 < total; s0 += XR_STAGE_UNROLL * XR_THREADS) {
                uint32_t a0[XR_STAGE_UNROLL], a1[XR_STAGE_UNROLL], a2[XR_STAGE_UNROLL];
                bool ok[XR_STAGE_UNROLL];
#pragma unroll
                for (int k = 0; k < XR_STAGE_UNROLL; k++) {   // every load of the round is issued before the first decode
                    const int sl = s0 + k * XR_THREADS + tid;
                    ok[k] = false;
                    if (sl < total) {
                        const int c = cellOf[sl];
                        const int j = sl + base - cPre[c];
                        if (j < cCnt[c]) {
                            const uint32_t gi = cBeg[c] + (uint32_t)j;
                            a0[k] = __ldg(&p.w0[gi]); a1[k] = __ldg(&p.w1[gi]); a2[k] = __ldg(&p.w2[gi]);
                            ok[k] = true;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < XR_STAGE_UNROLL; k++) {
                    if (ok[k]) {
                        float2* dst = reinterpret_cast<float2*>(sP + (size_t)(s0 + k * XR_THREADS + tid) * 6);
                        dst[0] = make_float2(fx_decode_fast(a0[k] & 0xffffu), fx_decode_fast(a0[k] >> 16));
                        dst[1] = make_float2(fx_decode_fast(a1[k] & 0xffffu), h_decode(a1[k] >> 16));
                        dst[2] = make_float2(h_decode(a2[k] & 0xffffu), h_decode(a2[k] >> 16));
                    }
                }
            }
            __syncthreads();
            if (vM && vZ && vP) xr_rows<true>(sP, cPre, cCnt, base, r0, r1, y, z, M, Z, P, true, true, true);
            else xr_rows<false>(sP, cPre, cCnt, base, r0, r1, y, z, M, Z, P, vM, vZ, vP);
        }
        // target x = cx-1 has seen its last plane: normalize_p2g_velocity (FF/FLIP_vdb.cpp:120-165) + sdf transform (:1199-1204)
        if (vP) {
            const int vo = ((cx - 1) << 6) | tid;
            const bool touched = P.cn > 0;
            bool on0 = false, on1 = false, on2 = false;
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
            if (touched && P.g0 != 0.f) { o0 = __fdiv_rn(P.a0, __fadd_rn(P.g0, 0.001f)); on0 = true; }
            if (touched && P.g1 != 0.f) { o1 = __fdiv_rn(P.a1, __fadd_rn(P.g1, 0.001f)); on1 = true; }
            if (touched && P.g2 != 0.f) { o2 = __fdiv_rn(P.a2, __fadd_rn(P.g2, 0.001f)); on2 = true; }
            const size_t gi = (size_t)leaf * LEAF + vo;
            p.vel[0][gi] = o0; p.vel[1][gi] = o1; p.vel[2][gi] = o2;
            float s = p.sdfBg;
            if (touched) s = fminf(s, __fsub_rn(__fmul_rn(p.dx, sqrtf(P.md)), p.radius));
            p.sdf[gi] = s;
            const unsigned bt = __ballot_sync(0xffffffffu, touched);
            const unsigned b0 = __ballot_sync(0xffffffffu, on0), b1 = __ballot_sync(0xffffffffu, on1), b2 = __ballot_sync(0xffffffffu, on2);
            if ((tid & 31) == 0) {
                const size_t wi = (size_t)leaf * 16 + (vo >> 5);
                const size_t nW = (size_t)p.t.n * 16;
                reinterpret_cast<uint32_t*>(p.topoMask)[wi] = bt;
                reinterpret_cast<uint32_t*>(p.chMask)[wi] = b0;
                reinterpret_cast<uint32_t*>(p.chMask)[nW + wi] = b1;
                reinterpret_cast<uint32_t*>(p.chMask)[2 * nW + wi] = b2;
            }
        }
        P = Z; Z = M; xacc_reset(M);
    }
    if (tid == 0 && planeSum) {   // the work-weighted mean plane (sum of squares / sum) sizes the next call's staging buffer
        atomicAdd(reinterpret_cast<unsigned long long*>(p.overflow + 2), planeSq);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.overflow + 4), planeSum);
    }
}

// air one-ring of the liquid sdf (FF/FLIP_vdb.cpp:1325-1382). ring = dilate26(topo) & ~topo.
__global__ void __launch_bounds__(512) air_ring_kernel(TopoView t, const uint64_t* __restrict__ topoMask,
                                                       const uint64_t* __restrict__ ringMask, float* __restrict__ sdf,
                                                       uint64_t* __restrict__ sdfMask, float dx, float bg) {
    int leaf = blockIdx.x, off = threadIdx.x;
    bool inTopo = mask_get(topoMask, leaf, off);
    bool inRing = mask_get(ringMask, leaf, off) && !inTopo;
    bool on = inTopo;
    float newSdf = 0.f;
    if (inRing) {
        int3 o = t.origin[leaf];
        int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
        bool hasLiquid = false;
        newSdf = sdf[(size_t)leaf * LEAF + off];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            int comp = i >> 1;
            int d = (i & 1) == 0 ? 1 : -1;
            int nx = x + (comp == 0 ? d : 0), ny = y + (comp == 1 ? d : 0), nz = z + (comp == 2 ? d : 0);
            // only voxels of the original topology can be negative; ring voxels hold bg or positive values
            int nl = topo_find(t, nx, ny, nz);
            float ns = bg;
            if (nl >= 0 && mask_get(topoMask, nl, voxel_off(nx, ny, nz))) ns = sdf[(size_t)nl * LEAF + voxel_off(nx, ny, nz)];
            if (ns < 0.f) { hasLiquid = true; newSdf = fminf(newSdf, __fadd_rn(dx, ns)); }
        }
        on = hasLiquid;
    }
    __syncthreads();  // all reads of neighbouring topo voxels in this leaf are done; ring writes touch ring voxels only
    if (inRing && on) sdf[(size_t)leaf * LEAF + off] = newSdf;
    unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<uint32_t*>(sdfMask)[(size_t)leaf * 16 + (threadIdx.x >> 5)] = b;
}

// one layer of union_extrapolate (FF/vdb_velocity_extrapolator.cpp:618-657). One WARP per (leaf, channel): a lane owns 16
// consecutive voxels and walks only the candidates among them (in the target topology, not valid yet) -- on all but the leaves at
// the rim of the valid region that is nothing, so a layer costs a read of the masks (round 1 and 2 launched 512 threads per leaf
// and channel: 45 us per layer, bound by the 17 K CTAs, for work on a few hundred leaves). Same arithmetic per voxel.
constexpr int EX_WARPS = 8;
__global__ void __launch_bounds__(EX_WARPS * 32) extrapolate_layer_kernel(TopoView t, const uint64_t* __restrict__ target,
                                                                           const uint64_t* __restrict__ validIn3,
                                                                           uint64_t* __restrict__ validOut3, float* __restrict__ val0,
                                                                           float* __restrict__ val1, float* __restrict__ val2, size_t stride) {
    const int item = blockIdx.x * EX_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (item >= t.n * 3) return;
    const int ch = item / t.n, leaf = item - ch * t.n;
    const uint64_t* validIn = validIn3 + ch * stride;
    uint64_t* validOut = validOut3 + ch * stride;
    float* val = ch == 0 ? val0 : (ch == 1 ? val1 : val2);
    const uint32_t tg = (reinterpret_cast<const uint32_t*>(target)[(size_t)leaf * 16 + (lane >> 1)] >> ((lane & 1) * 16)) & 0xffffu;
    const uint32_t vd = (reinterpret_cast<const uint32_t*>(validIn)[(size_t)leaf * 16 + (lane >> 1)] >> ((lane & 1) * 16)) & 0xffffu;
    uint32_t cand = tg & ~vd, newv = vd;
    if (__any_sync(0xffffffffu, cand != 0)) {
        // the warp's candidates as one list, dealt round-robin to the lanes (a rim leaf holds a few dozen to ~200, unevenly spread)
        __shared__ uint16_t sList[EX_WARPS][LEAF];
        __shared__ uint32_t sNew[EX_WARPS][16];
        uint16_t* list = sList[threadIdx.x >> 5];
        uint32_t* snew = sNew[threadIdx.x >> 5];
        const int mine = __popc(cand);
        int pre = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre += v; }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        int at = pre - mine;
        while (cand) { const int bit = __ffs(cand) - 1; cand &= cand - 1; list[at++] = (uint16_t)(lane * 16 + bit); }
        if (lane < 16) snew[lane] = 0u;
        __syncwarp();
        const int3 o = t.origin[leaf];
        for (int k = lane; k < total; k += 32) {
            const int off = list[k];
            const int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
            int tw = 0;
            float sum = 0.f;
#pragma unroll
            for (int d = 0; d < 6; d++) {
                const int dir = d >> 1, s = (d & 1) ? 1 : -1;
                const int nx = x + (dir == 0 ? s : 0), ny = y + (dir == 1 ? s : 0), nz = z + (dir == 2 ? s : 0);
                int nl, no;
                const int lx = (off >> 6) + (dir == 0 ? s : 0), ly = ((off >> 3) & 7) + (dir == 1 ? s : 0), lz = (off & 7) + (dir == 2 ? s : 0);
                if ((unsigned)lx < 8u && (unsigned)ly < 8u && (unsigned)lz < 8u) { nl = leaf; no = (lx << 6) | (ly << 3) | lz; }
                else { nl = topo_find(t, nx, ny, nz); no = voxel_off(nx, ny, nz); }
                if (nl >= 0 && mask_get(validIn, nl, no)) { tw++; sum = __fadd_rn(sum, val[(size_t)nl * LEAF + no]); }
            }
            if (tw != 0) { val[(size_t)leaf * LEAF + off] = __fdiv_rn(sum, (float)tw); atomicOr(&snew[off >> 5], 1u << (off & 31)); }
        }
        __syncwarp();
        newv |= (snew[lane >> 1] >> ((lane & 1) * 16)) & 0xffffu;
    }
    const uint32_t other = __shfl_xor_sync(0xffffffffu, newv, 1);
    if ((lane & 1) == 0) reinterpret_cast<uint32_t*>(validOut)[(size_t)leaf * 16 + (lane >> 1)] = newv | (other << 16);
}
// to_vec3 (packed3grids.cpp:49-83): union mask; a channel that is off contributes 0
__global__ void __launch_bounds__(512) to_vec3_kernel(int n, const uint64_t* __restrict__ chMask, float* __restrict__ v0,
                                                      float* __restrict__ v1, float* __restrict__ v2,
                                                      uint64_t* __restrict__ outMask) {
    int leaf = blockIdx.x, off = threadIdx.x;
    size_t stride = (size_t)n * 8;
    bool o0 = mask_get(chMask, leaf, off), o1 = mask_get(chMask + stride, leaf, off), o2 = mask_get(chMask + 2 * stride, leaf, off);
    size_t i = (size_t)leaf * LEAF + off;
    if (!o0) v0[i] = 0.f;
    if (!o1) v1[i] = 0.f;
    if (!o2) v2[i] = 0.f;
    unsigned b = __ballot_sync(0xffffffffu, o0 | o1 | o2);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<uint32_t*>(outMask)[(size_t)leaf * 16 + (threadIdx.x >> 5)] = b;
}
__global__ void andnot_kernel(uint64_t* __restrict__ a, const uint64_t* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] &= ~b[i];
}
}  // namespace

void union_extrapolate(World* w, int nLayer, GridV& vel, uint64_t* chMask, const uint64_t* targetMask) {
    const Topo& t = *vel.topo;
    if (t.n == 0 || nLayer <= 0) return;
    size_t stride = (size_t)t.n * 8;
    DBuf<uint64_t> tmp(3 * stride, w->stream);
    uint64_t* cur = chMask;
    uint64_t* nxt = tmp.p;
    for (int layer = 0; layer < nLayer; layer++) {
        FB_LAUNCH(w, "extrapolate_layer", (size_t)3 * t.n * (2048 + 192))
            extrapolate_layer_kernel<<<(3 * t.n + EX_WARPS - 1) / EX_WARPS, EX_WARPS * 32, 0, w->stream>>>(t.view(), targetMask, cur, nxt, vel.val[0].p, vel.val[1].p, vel.val[2].p, stride);
        check_launch("extrapolate_layer");
        std::swap(cur, nxt);
    }
    if (cur != chMask) FB_CUDA(cudaMemcpyAsync(chMask, cur, 3 * stride * 8, cudaMemcpyDeviceToDevice, w->stream));
}

void finish_vec3(World* w, GridV& vel, const uint64_t* chMask) {
    int n = vel.topo->n;
    if (!n) return;
    FB_LAUNCH(w, "to_vec3", (size_t)n * (6144 + 256)) to_vec3_kernel<<<n, 512, 0, w->stream>>>(n, chMask, vel.val[0].p, vel.val[1].p, vel.val[2].p, vel.mask.p);
    check_launch("to_vec3");
}

void check_p2g_overflow(World* w) {   // call right after a host wait on the stream
    if (w->p2gOverflowHost && *w->p2gOverflowHost) {
        *w->p2gOverflowHost = 0;
        throw Error(FLIPB200_ERR_ARG, "FLIP_P2G: a voxel holds more than 1400 particles");
    }
}

// FLIP_P2G::apply (FF/nosys/P2G.cpp:11-42)
void p2g(World* w, float dx, int velExtraLayer) {
    FB_REQUIRE(w->pts.topo != nullptr, FLIPB200_ERR_STATE, "FLIP_P2G: no particles");
    ensure_pool(w, {}, /*includeParticles=*/true);
    TopoPtr pool = w->pool;
    const int n = pool->n;
    GridV& vel = w->V(FLIPB200_VELOCITY);
    GridV& post = w->V(FLIPB200_POSTADV_VELOCITY);
    GridF& sdf = w->F(FLIPB200_LIQUID_SDF);
    const float zero3[3] = {0.f, 0.f, 0.f};
    // outputs are brand-new grids on the pool (the reference setTree()s new trees)
    GridV nvel; nvel.topo = pool;
    for (int c = 0; c < 3; c++) { nvel.bg[c] = 0.f; nvel.val[c].alloc((size_t)n * LEAF, w->stream); }
    nvel.mask.alloc((size_t)n * 8, w->stream);
    GridF nsdf; nsdf.topo = pool; nsdf.bg = sdf.bg;
    nsdf.val.alloc((size_t)n * LEAF, w->stream);
    nsdf.mask.alloc((size_t)n * 8, w->stream);
    nsdf.alloc.alloc(n ? n : 1, w->stream);
    nsdf.alloc.zero();
    (void)zero3;
    if (n == 0) {
        vel = std::move(nvel); grid_copy(w, post, vel); sdf = std::move(nsdf);
        if (dd_on(w)) { dd_refresh(w, vel, 2); dd_refresh(w, post, 2); dd_refresh(w, sdf, 2); }
        return;
    }

    DBuf<uint64_t> chMask((size_t)3 * n * 8, w->stream), topoMask((size_t)n * 8, w->stream), ring((size_t)n * 8, w->stream);
    if (!w->p2gOverflowHost) {
        w->p2gOverflowHost = reinterpret_cast<int*>(w->hostScratch + 768);   // two words of the mapped scratch block
        for (int k = 0; k < 6; k++) w->p2gOverflowHost[k] = 0;
        w->p2gOverflow.alloc(6, w->stream);
    }
    DBuf<int>& overflow = w->p2gOverflow;
    overflow.zero();
    P2GParams p;
    p.cap = 0;
    p.t = pool->view();
    p.voxelStart = w->pts.voxelStart.p;
    p.w0 = w->pts.w0.p; p.w1 = w->pts.w1.p; p.w2 = w->pts.w2.p;
    for (int c = 0; c < 3; c++) p.vel[c] = nvel.val[c].p;
    p.chMask = chMask.p;
    p.sdf = nsdf.val.p;
    p.topoMask = topoMask.p;
    p.dx = dx;
    p.radius = dx * 0.8f * 1.01f;  // FF/FLIP_vdb.cpp:1287
    p.sdfBg = sdf.bg;
    p.overflow = overflow.p;
    size_t smemBytes = (size_t)(PLANE_CAP * 6 + 8 * LEAF) * sizeof(float);
    static bool attrSet = false;
    if (!attrSet) {
        FB_CUDA(cudaFuncSetAttribute(p2g_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        attrSet = true;
    }
    // algorithmic bytes: 12 B/particle + 4 B/voxel offsets + 16 B/voxel outputs
    const bool oldPath = getenv("FLIPB200_P2G_OLD") != nullptr;   // read per call: tests/test_p2g_paths_gpu.py compares both
    if (oldPath) {
        FB_LAUNCH(w, "p2g_gather", w->pts.n * 12 + (size_t)n * LEAF * 20)
            p2g_gather_kernel<<<n, P2G_THREADS, smemBytes, w->stream>>>(p, nullptr);
        check_launch("p2g_gather");
    } else {
        DBuf<uint8_t> todo((size_t)n, w->stream);
        todo.zero();
        // staging depth: the work-weighted mean plane (sum of squares / sum of the per-plane record counts) of the PREVIOUS call (read back without a wait, so it is one call old) plus head
        // room; 1024 records = 8 CTAs per SM at 8 particles per voxel, up to 2304 (4 CTAs) for denser stores; a plane beyond the
        // depth is processed in row batches
        int cap = XR_CAP;
        unsigned long long sq = 0, sum = 0;
        memcpy(&sq, w->p2gOverflowHost + 2, 8); memcpy(&sum, w->p2gOverflowHost + 4, 8);
        const int seen = sum ? (int)(sq / sum) : 0;
        if (seen + 64 > cap) cap = std::min(XR_CAP_MAX, (seen + 64 + 127) & ~127);
        if (const char* e = getenv("FLIPB200_P2G_CAP")) cap = std::max(512, std::min(XR_CAP_MAX, atoi(e) & ~63));
        p.cap = cap;
        const size_t xrSmem = (size_t)cap * 25;
        static size_t attrSmem = 0;
        if (xrSmem > attrSmem) {
            FB_CUDA(cudaFuncSetAttribute(p2g_xrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)XR_CAP_MAX * 25)));
            attrSmem = (size_t)XR_CAP_MAX * 25;
        }
        FB_LAUNCH(w, "p2g_gather", w->pts.n * 12 + (size_t)n * LEAF * 20)
            p2g_xrow_kernel<<<n, XR_THREADS, xrSmem, w->stream>>>(p, todo.p);
        check_launch("p2g_xrow");
        // leaves with a cell row beyond the staging buffer (only an un-capped initial binning can produce one)
        FB_LAUNCH(w, "p2g_gather_fallback", 0)
            p2g_gather_kernel<<<n, P2G_THREADS, smemBytes, w->stream>>>(p, todo.p);
        check_launch("p2g_gather_fallback");
    }

    // air ring: airmask = dilate26(topo) \ topo ; final sdf mask = topo + ring voxels with a liquid face neighbour
    mask_dilate(w, *pool, topoMask.p, ring.p, true);
    FB_LAUNCH(w, "p2g_air_ring", (size_t)n * (2048 + 256)) air_ring_kernel<<<n, 512, 0, w->stream>>>(pool->view(), topoMask.p, ring.p, nsdf.val.p, nsdf.mask.p, dx, sdf.bg);
    check_launch("air_ring");
    // leaves of the reference's sdf tree: everything touched by the dilated topology (:1380)
    mark_alloc_from_mask(w, ring.p, n, nsdf.alloc.p);

    // post-P2G copy before extrapolation (FF/FLIP_vdb.cpp:1386), then extrapolate + to_vec3
    GridV npost;
    grid_copy(w, npost, nvel);
    finish_vec3(w, npost, chMask.p);
    union_extrapolate(w, velExtraLayer, nvel, chMask.p, nsdf.mask.p);
    finish_vec3(w, nvel, chMask.p);

    d2h_words(w, w->p2gOverflowHost, overflow.p, 24);   // [0] checked by check_p2g_overflow, [2..5] size the next call's staging
    vel = std::move(nvel);
    post = std::move(npost);
    sdf = std::move(nsdf);
    if (dd_on(w)) {
        // the two leaf layers beyond each slab face were computed from an incomplete particle set: take the owner's
        // (one exchange for the three grids: 11 arrays of the same leaf layers)
        dd_refresh(w, {DDArray{vel.val[0].p, LEAF * 4}, DDArray{vel.val[1].p, LEAF * 4}, DDArray{vel.val[2].p, LEAF * 4}, DDArray{vel.mask.p, 64},
                       DDArray{post.val[0].p, LEAF * 4}, DDArray{post.val[1].p, LEAF * 4}, DDArray{post.val[2].p, LEAF * 4}, DDArray{post.mask.p, 64},
                       DDArray{sdf.val.p, LEAF * 4}, DDArray{sdf.mask.p, 64}, DDArray{sdf.alloc.p, 1}}, 2);
    }
}

}  // namespace fb
