// libflipb200 -- the bottom of the mu-cycle inside ONE thread-block cluster (included by poisson.cu).
//
// Below the first one or two levels a level is a few dozen 8^3 leaves, and the mu = 2 cycle visits it 4-16 times per
// preconditioner application, 18 colour passes per visit: the work per pass is a few thousand stencil updates and the
// cost is the synchronisation between dependent passes. mg_cluster_kernel runs every level from `first` down as ONE
// cluster of C CTAs (C = 16 where the part allows it, else 8):
//   * CTA 0 holds the coarsest level as compact rows and runs its Jacobi-CG;
//   * CTAs 1..C-1 hold the leaves of all levels in their shared memory for the whole launch: iterate x, right-hand
//     side b, the six 8x8 HALO faces of the neighbour leaves, and -- for leaves with non-default coefficients --
//     invdiag, the three "minus" face arrays and the three "plus" faces of the neighbours. Every array is stored SPLIT BY
//     COLOUR (red = (x + y + z) even first): voxel (X, Y, Z) of colour c sits at c * 256 + (X << 5 | Y << 2 | Z >> 1), so
//     the four voxels of one colour in a z-row are one aligned float4, and so are their x / y neighbours of the other colour;
//   * a colour pass is 64 threads per leaf, ONE Z-ROW (four updates) PER THREAD, all operands as 128-bit shared-memory
//     loads: ~30 issued instructions per update instead of ~65 with one voxel per thread. The passes are instruction-issue
//     bound (measured: a version with the same loads but one voxel per thread and four leaves interleaved per thread was
//     not faster than the plain one), so the instruction count is what matters;
//   * a colour pass reads local shared memory only; every updated face row is PUSHED into the neighbour leaf's halo
//     with st.shared::cluster (asynchronous: measured on B200 a dependent ld.shared::cluster costs > 1 us per pass at
//     16 x 1024 threads, a store costs its issue slot) -- the classic halo exchange, between SMs;
//   * passes are separated by barrier.cluster (measured 0.29 / 0.38 us for 16 x 256 / 1024 threads) instead of a
//     device-wide barrier (1.4 us) or a launch boundary (0.95 us in a graph, 3.5 us in a stream). The barrier's acquire
//     invalidates the SM's L1 (CCTL.IVALL), so nothing in the op loop reads global or constant memory;
//   * restriction stores into the parent leaf's b, prolongation reads the parent's x through the same window.
// The arithmetic of every op is that of rbgs_leaf / zero_red_leaf / ax_voxel / restrict_leaf / prolong_voxel; with
// FLIPB200_CG_COMPAT=1 (compact_cg's reduction trees in the coarsest CG) the result is bit-identical to the round-1
// cycle kernel (tests/test_mg_paths_gpu.py); the default CG sums in a cheaper, equally fixed order.

constexpr int CL_THREADS = 512;        // eight leaves (64 threads each) per sweep of the CTA
constexpr int CL_GROUPS = CL_THREADS / 64;
constexpr int CL_MAX_LEVELS = 6;
constexpr int CL_MAX_OPS = 2048;
constexpr int CL_XB_FLOATS = 512 + 512 + 6 * 64;    // x[2][256], b[2][256], halo faces -x +x -y +y -z +z as [6][2][32]
constexpr int CL_XB_BYTES = CL_XB_FLOATS * 4;
constexpr int CL_HALO = 1024;                        // float offset of the halo faces inside an XB block
constexpr int CL_COEF_FLOATS = 4 * 512 + 3 * 64;     // invdiag, xe, ye, ze as [2][256] each, then the +x, +y, +z neighbour faces of xe / ye / ze as [3][2][32]
constexpr int CL_COEF_BYTES = CL_COEF_FLOATS * 4;
constexpr uint32_t CL_NONE = 0xffffffffu;

struct ClMeta {            // one per local leaf slot; built once per solve in global memory (cl_meta_kernel), copied at launch
    uint32_t nb[6];        // assign word of the -x,+x,-y,+y,-z,+z neighbour leaf (CL_NONE = none)
    uint32_t parent;       // assign word of the parent leaf in the next coarser level
    uint32_t parentOff;    // offset of this leaf's octant inside the parent leaf
    int leaf;              // slot of the leaf in its level, -1 = unused
    uint32_t flags;        // LI_*
    int coef;              // coefficient block index or -1
    uint32_t pad;
    uint64_t mask[8];      // DOF mask
    uint64_t parentMask[8];
};
static_assert(sizeof(ClMeta) == 176, "ClMeta layout");
struct ClLevel {
    LevelView v;
    const uint32_t* assign;   // [n] rank | xb index << 8 | coef index << 20; CL_NONE = leaf without DOFs (no slot)
    const ClMeta* meta;       // [C][xbPer]
    int n, xbPer, coefPer;    // leaves of the level; XB / coefficient blocks per CTA
    int nN, nC;               // leaves with non-default / default coefficients: CTA r holds the slots [0, cntN) and [coefPer, coefPer + cntC)
    int xbOff, coefOff, metaOff;   // byte offsets into dynamic shared memory (the same in every CTA)
};
struct ClusterParams {
    ClLevel lv[CL_MAX_LEVELS];
    int nLevels, nOps;
    const uint8_t* prog;      // op | (level - first) << 3
    float* topX; const float* topB;   // level lv[0]'s vectors in global memory
    int loadX;                // the first op continues from topX (no zero guess)
    int xbBytes;              // total bytes of the XB region (starts at offset 0)
    float w, oneMinusW, prolongAlpha;
    CompactDev cg;            // the coarsest level's compact rows (blob in global memory)
    int cgOff;                // CTA 0: [diag np][minus 3 np][cols 6 np u16][x np][b np][P np][T np]
    int sresOff;              // CL_GROUPS x 512 floats of residual scratch
    int progOff;              // CL_MAX_OPS bytes
    int dumpOff;              // one XB block: halo pushes towards a missing neighbour land here
    int cgCompat;             // 1: compact_cg's summation order (bit-identical to the round-1 path)
    unsigned long long* trace;
    int dbg;                  // timing experiments: 1 = no halo pushes, 2 = colour passes do nothing, 4 = no stores
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cl_mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ float cl_ld(uint32_t a) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
// (no "memory" clobber: the stores only have to stay ahead of the next barrier, and volatile asm statements keep their order)
__device__ __forceinline__ void cl_st(uint32_t a, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void cl_st4(uint32_t a, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void bar_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ bool meta_bit(const uint64_t* m, int off) { return (m[off >> 6] >> (off & 63)) & 1ull; }

// natural voxel offset (x << 6 | y << 3 | z) -> colour-split index
__device__ __forceinline__ int split_idx(int off) {
    const int X = off >> 6, Y = (off >> 3) & 7, Z = off & 7;
    return (((X + Y + Z) & 1) << 8) | (X << 5) | (Y << 2) | (Z >> 1);
}
// the per-slot record as the passes use it (shared memory): ClMeta with addresses resolved + the colour-split DOF mask
struct ClSlot {
    uint32_t nb[6];        // shared::cluster address of the neighbour's XB block; a missing neighbour points at this CTA's dump block
    uint32_t parent;       // shared::cluster address of the parent's XB block, 0 = none
    uint32_t parentOff;
    int leaf;
    uint32_t flags;
    int coef;              // byte offset of the coefficient block, -1 = default coefficients
    uint32_t pad;
    uint32_t cmask[16];    // DOF bits, [colour][x] bit (y << 2 | z >> 1)
    uint64_t mask[8];      // natural DOF mask
    uint64_t parentMask[8];
};
static_assert(sizeof(ClSlot) == 240, "ClSlot layout");

// Everything about "the z-row (X, Y) of colour c" that does not depend on the leaf. A row is the four voxels
// Z = 2k + p, k = 0..3, p = (X + Y + c) & 1; its x / y neighbours are the same row indices of the other colour one row
// over (or a halo face row), its z neighbours are the other colour's row (X, Y) shifted by p, plus one halo voxel.
struct RowIdx {
    int own, xp, xm, yp, ym, zr, zh;   // float offsets into the XB block (b at 512 + own); zh = the one z-halo voxel
    int cxp, cyp, czr, czf;            // coefficient block: xe / ye rows of the +x / +y neighbour, ze row of the other colour, ze face voxel
    int p;                             // 0: the row starts at z = 0 (its -z neighbour is in the halo), 1: it ends at z = 7
    int pushX, pushY, pushZ;           // halo offsets written in the neighbours (-1: the row is not on that face)
    int nbX, nbY, nbZ;                 // which neighbour
    int word, shift;                   // DOF bits of the row: (cmask[word] >> shift) & 15
};
__device__ __forceinline__ RowIdx row_idx(int q, int c) {
    RowIdx r;
    const int X = q >> 3, Y = q & 7, p = (X + Y + c) & 1, t0 = (X << 5) | (Y << 2);
    const int O = (1 - c) * 256, H = CL_HALO + (1 - c) * 32, Hc = CL_HALO + c * 32;
    const int fz = (X << 2) | (Y >> 1);
    r.p = p;
    r.own = c * 256 + t0;
    r.xp = X < 7 ? O + t0 + 32 : H + 64 + (Y << 2);     r.xm = X > 0 ? O + t0 - 32 : H + (Y << 2);
    r.yp = Y < 7 ? O + t0 + 4 : H + 192 + (X << 2);     r.ym = Y > 0 ? O + t0 - 4 : H + 128 + (X << 2);
    r.zr = O + t0;
    r.zh = p == 0 ? H + 256 + fz : H + 320 + fz;
    r.cxp = X < 7 ? 512 + O + t0 + 32 : 2048 + (1 - c) * 32 + (Y << 2);
    r.cyp = Y < 7 ? 1024 + O + t0 + 4 : 2048 + 64 + (1 - c) * 32 + (X << 2);
    r.czr = 1536 + O + t0;
    r.czf = 2048 + 128 + (1 - c) * 32 + fz;
    // my x = 0 plane is the +x halo (face 1) of the -x neighbour (nb 0), and so on
    r.pushX = X == 0 ? Hc + 64 + (Y << 2) : (X == 7 ? Hc + (Y << 2) : -1);          r.nbX = X == 0 ? 0 : 1;
    r.pushY = Y == 0 ? Hc + 192 + (X << 2) : (Y == 7 ? Hc + 128 + (X << 2) : -1);   r.nbY = Y == 0 ? 2 : 3;
    r.pushZ = p == 0 ? Hc + 320 + fz : Hc + 256 + fz;                               r.nbZ = p == 0 ? 4 : 5;
    r.word = c * 8 + X; r.shift = Y << 2;
    return r;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void push_row(const ClSlot& m, const RowIdx& r, float4 v) {
    if (r.pushX >= 0) cl_st4(m.nb[r.nbX] + (uint32_t)r.pushX * 4u, v);
    if (r.pushY >= 0) cl_st4(m.nb[r.nbY] + (uint32_t)r.pushY * 4u, v);
    cl_st(m.nb[r.nbZ] + (uint32_t)r.pushZ * 4u, r.p == 0 ? v.x : v.w);
}
// off-diagonal sums of a row, each with the association of offdiag()
template <bool COEF>
__device__ __forceinline__ float4 row_offdiag(const float* xb, const float* cb, const RowIdx& r, float def) {
    const float4 xp = ld4(xb + r.xp), xm = ld4(xb + r.xm), yp = ld4(xb + r.yp), ym = ld4(xb + r.ym), zr = ld4(xb + r.zr);
    const float zh = xb[r.zh];
    const float4 zm = r.p == 0 ? make_float4(zh, zr.x, zr.y, zr.z) : zr;
    const float4 zp = r.p == 0 ? zr : make_float4(zr.y, zr.z, zr.w, zh);
    float4 cxp, cxm, cyp, cym, czp, czm;
    if (COEF) {
        cxp = ld4(cb + r.cxp); cxm = ld4(cb + 512 + r.own); cyp = ld4(cb + r.cyp); cym = ld4(cb + 1024 + r.own);
        czm = ld4(cb + 1536 + r.own);
        const float4 zo = ld4(cb + r.czr);
        const float zf = cb[r.czf];
        czp = r.p == 0 ? zo : make_float4(zo.y, zo.z, zo.w, zf);
    } else {
        cxp = cxm = cyp = cym = czp = czm = make_float4(def, def, def, def);
    }
    float4 o;
#define FB_ROW_OD(e) o.e = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xp.e, cxp.e), __fmul_rn(xm.e, cxm.e)), __fadd_rn(__fmul_rn(yp.e, cyp.e), __fmul_rn(ym.e, cym.e))), \
                                     __fadd_rn(__fmul_rn(zp.e, czp.e), __fmul_rn(zm.e, czm.e)))
    FB_ROW_OD(x); FB_ROW_OD(y); FB_ROW_OD(z); FB_ROW_OD(w);
#undef FB_ROW_OD
    return o;
}
// one colour of red-black SOR (rbgs_leaf) on the slots first + g, first + g + CL_GROUPS, ... (g = this thread's group of 64)
template <bool COEF>
__device__ __forceinline__ void colour_pass(unsigned char* sm, int xbOff, const ClSlot* slots, int first, int cnt, int g, const RowIdx& r,
                                            float def, float defInv, float w, float oneMinusW, int dbg) {
    for (int i = g; i < cnt; i += CL_GROUPS) {
        const ClSlot& m = slots[first + i];
        float* xb = reinterpret_cast<float*>(sm + xbOff + (size_t)(first + i) * CL_XB_BYTES);
        const float* cb = COEF ? reinterpret_cast<const float*>(sm + m.coef) : nullptr;
        const float4 od = row_offdiag<COEF>(xb, cb, r, def);
        const float4 xi = ld4(xb + r.own), bi = ld4(xb + 512 + r.own);
        const float4 inv = COEF ? ld4(cb + r.own) : make_float4(defInv, defInv, defInv, defInv);
        const uint32_t on = (m.cmask[r.word] >> r.shift) & 15u;
        float4 nv;
#define FB_ROW_UP(e, b) nv.e = (on & b) ? __fmaf_rn(xi.e, oneMinusW, __fmul_rn(__fmul_rn(__fsub_rn(bi.e, od.e), inv.e), w)) : xi.e
        FB_ROW_UP(x, 1u); FB_ROW_UP(y, 2u); FB_ROW_UP(z, 4u); FB_ROW_UP(w, 8u);
#undef FB_ROW_UP
        if (dbg & 4) continue;
        *reinterpret_cast<float4*>(xb + r.own) = nv;
        if (!(dbg & 1)) push_row(m, r, nv);
    }
}

// ranks the leaves of one level by class (non-default coefficients / default / no DOF) and deals them out in blocks
// over the CTAs 1..ranks; counts = {non-default, default}
__global__ void __launch_bounds__(1024) cl_assign_kernel(const LeafInfo* __restrict__ info, int n, int ranks, uint32_t* __restrict__ assign,
                                                         int* __restrict__ counts) {
    __shared__ int wsumN[32], wsumC[32];
    const int j = threadIdx.x;
    int cls = 0;
    if (j < n) {
        const uint32_t f = info[j].flags;
        if (f & LI_ANY) cls = ((f & LI_CONST) && (f & LI_DIAG)) ? 2 : 1;
    }
    const unsigned bn = __ballot_sync(0xffffffffu, cls == 1), bc = __ballot_sync(0xffffffffu, cls == 2);
    const int lane = j & 31, wid = j >> 5;
    const int pn = __popc(bn & ((1u << lane) - 1u)), pc = __popc(bc & ((1u << lane) - 1u));
    if (lane == 0) { wsumN[wid] = __popc(bn); wsumC[wid] = __popc(bc); }
    __syncthreads();
    int baseN = 0, baseC = 0, totN = 0, totC = 0;
    for (int k = 0; k < 32; k++) {
        if (k < wid) { baseN += wsumN[k]; baseC += wsumC[k]; }
        totN += wsumN[k]; totC += wsumC[k];
    }
    const int per = (totN + ranks - 1) / ranks, perC = (totC + ranks - 1) / ranks;
    if (j < n) {
        uint32_t a = CL_NONE;
        if (cls == 1) { const int k = baseN + pn; a = (uint32_t)(1 + k / per) | ((uint32_t)(k % per) << 8) | ((uint32_t)(k % per) << 20); }
        else if (cls == 2) { const int k = baseC + pc; a = (uint32_t)(1 + k / perC) | ((uint32_t)(per + k % perC) << 8) | (0xfffu << 20); }
        assign[j] = a;
    }
    if (j == 0) { counts[0] = totN; counts[1] = totC; }
}
// the per-slot records of one level (C = next coarser level of the cluster, or null)
__global__ void cl_meta_kernel(LevelView L, const uint32_t* __restrict__ assign, int n, int xbPer, ClMeta* __restrict__ meta,
                               LevelView C, const uint32_t* __restrict__ cassign) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t a = assign[j];
    if (a == CL_NONE) return;
    ClMeta m;
    const LeafInfo* li = L.info + j;
    for (int k = 0; k < 6; k++) { const int nb = li->nb[k]; m.nb[k] = nb >= 0 ? assign[nb] : CL_NONE; }
    m.leaf = j; m.flags = li->flags; m.pad = 0;
    m.coef = (a >> 20) == 0xfffu ? -1 : (int)(a >> 20);
    for (int k = 0; k < 8; k++) { m.mask[k] = li->mask[k]; m.parentMask[k] = 0; }
    m.parent = CL_NONE; m.parentOff = 0;
    if (cassign) {
        const int3 o = L.t.origin[j];
        const int cl = topo_find(C.t, o.x >> 1, o.y >> 1, o.z >> 1);
        if (cl >= 0 && cassign[cl] != CL_NONE) {
            m.parent = cassign[cl];
            m.parentOff = (uint32_t)((((o.x >> 1) & 7) << 6) | (((o.y >> 1) & 7) << 3) | ((o.z >> 1) & 7));
            for (int k = 0; k < 8; k++) m.parentMask[k] = C.info[cl].mask[k];
        }
    }
    meta[(size_t)(a & 0xffu) * xbPer + ((a >> 8) & 0xfffu)] = m;
}

__device__ __forceinline__ uint32_t cl_xb_addr_s(int xbOff, uint32_t smBase, uint32_t a) {
    return cl_mapa(smBase + (uint32_t)xbOff + ((a >> 8) & 0xfffu) * (uint32_t)CL_XB_BYTES, a & 0xffu);
}
__device__ __forceinline__ uint32_t cl_xb_addr(const ClLevel& L, uint32_t smBase, uint32_t a) {
    return cl_mapa(smBase + (uint32_t)L.xbOff + ((a >> 8) & 0xfffu) * (uint32_t)CL_XB_BYTES, a & 0xffu);
}

// Jacobi-preconditioned CG on the compact coarsest level with CL_THREADS threads. compact_cg runs on 1024 threads; here thread t
// plays its threads t, t + T, t + 2T, ... (K = 1024 / T of them): the same per-thread partial sums, the same butterfly inside
// each of the 32 warps, the same tree over the 32 warp partials -- bit-identical results, a quarter of the warps at every barrier.
template <int T>
struct CgSum {
    static constexpr int K = 1024 / T;
    // the butterfly over 32 warp partials, as lane 0 of compact_cg's second stage evaluates it
    static __device__ __forceinline__ float tree32(const float* p) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = __fadd_rn(p[i], p[i + 16]);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = __fadd_rn(a[i], a[i + 8]);
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = __fadd_rn(a[i], a[i + 4]);
        a[0] = __fadd_rn(a[0], a[2]); a[1] = __fadd_rn(a[1], a[3]);
        return __fadd_rn(a[0], a[1]);
    }
    static __device__ __forceinline__ float sum(float (&v)[K], float* red, unsigned& phase) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
#pragma unroll
            for (int j = 0; j < K; j++) v[j] = __fadd_rn(v[j], __shfl_xor_sync(0xffffffffu, v[j], d));
        float* buf = red + (phase & 1u) * 32;
        phase++;
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int j = 0; j < K; j++) buf[j * (T / 32) + (threadIdx.x >> 5)] = v[j];
        __syncthreads();
        return tree32(buf);
    }
    static __device__ __forceinline__ float2 sum2(float (&a)[K], float (&b)[K], float* red, unsigned& phase) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
#pragma unroll
            for (int j = 0; j < K; j++) {
                a[j] = __fadd_rn(a[j], __shfl_xor_sync(0xffffffffu, a[j], d));
                b[j] = __fadd_rn(b[j], __shfl_xor_sync(0xffffffffu, b[j], d));
            }
        float* buf = red + (phase & 1u) * 64;
        phase++;
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int j = 0; j < K; j++) { buf[j * (T / 32) + (threadIdx.x >> 5)] = a[j]; buf[32 + j * (T / 32) + (threadIdx.x >> 5)] = b[j]; }
        __syncthreads();
        return make_float2(tree32(buf), tree32(buf + 32));
    }
};
// The default order of the sums (any fixed order is as good as another here: the coarsest solve is Eigen's CG in the reference,
// whose order is its own): a thread's K rows in ascending order, one butterfly per warp, then EVERY thread folds the T / 32 warp
// partials itself as a balanced tree -- one bar.sync per sum, no second stage.
template <int T>
struct CgSumFast {
    static constexpr int K = 1024 / T, W = T / 32;
    static __device__ __forceinline__ float fold(const float* p) {
        float a[W];
#pragma unroll
        for (int i = 0; i < W; i += 4) { const float4 q = *reinterpret_cast<const float4*>(p + i); a[i] = q.x; a[i + 1] = q.y; a[i + 2] = q.z; a[i + 3] = q.w; }
#pragma unroll
        for (int h = W / 2; h > 0; h >>= 1)
#pragma unroll
            for (int i = 0; i < h; i++) a[i] = __fadd_rn(a[i], a[i + h]);
        return a[0];
    }
    static __device__ __forceinline__ float sum(float (&v)[K], float* red, unsigned& phase) {
        float x = v[0];
#pragma unroll
        for (int j = 1; j < K; j++) x = __fadd_rn(x, v[j]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, d));
        float* buf = red + (phase & 1u) * 64;
        phase++;
        if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = x;
        __syncthreads();
        return fold(buf);
    }
    static __device__ __forceinline__ float2 sum2(float (&a)[K], float (&b)[K], float* red, unsigned& phase) {
        float x = a[0], y = b[0];
#pragma unroll
        for (int j = 1; j < K; j++) { x = __fadd_rn(x, a[j]); y = __fadd_rn(y, b[j]); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, d)); y = __fadd_rn(y, __shfl_xor_sync(0xffffffffu, y, d)); }
        float* buf = red + (phase & 1u) * 64;
        phase++;
        if ((threadIdx.x & 31) == 0) { buf[threadIdx.x >> 5] = x; buf[32 + (threadIdx.x >> 5)] = y; }
        __syncthreads();
        return make_float2(fold(buf), fold(buf + 32));
    }
};
// x = S.x, the residual lives in S.b, P and T in `pt` (2 x np floats); red = 128 floats
template <int T>
__device__ void compact_cg_k(const CompactSm& S, float* pt, float* red, unsigned& phase) {
    constexpr int K = 1024 / T;
    using Sum = CgSum<T>;
    const int n = S.n, np = S.np, tid = threadIdx.x;
    float* X = S.x; float* R = S.b; float* P = pt; float* Tv = pt + np;
    auto dinv = [&](int r) { float d = S.diag[r]; return d != 0.f ? __fdiv_rn(1.0f, d) : 1.0f; };
    for (int r = tid; r < np; r += T) { X[r] = 0.f; P[r] = 0.f; if (r >= n) R[r] = 0.f; }
    float acc[K], acc2[K];
#pragma unroll
    for (int j = 0; j < K; j++) { acc[j] = 0.f; for (int r = j * T + tid; r < n; r += 1024) acc[j] = __fadd_rn(acc[j], __fmul_rn(R[r], R[r])); }
    const float rhsNorm2 = Sum::sum(acc, red, phase);
    if (rhsNorm2 != 0.f) {
        const float tol = 1.1920929e-07f;
        const float threshold = fmaxf(__fmul_rn(__fmul_rn(tol, tol), rhsNorm2), 1.17549435e-38f);
        float residualNorm2 = rhsNorm2;
        if (!(residualNorm2 < threshold)) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                acc[j] = 0.f;
                for (int r = j * T + tid; r < n; r += 1024) { const float pv = __fmul_rn(dinv(r), R[r]); P[r] = pv; acc[j] = __fadd_rn(acc[j], __fmul_rn(R[r], pv)); }
            }
            float absNew = Sum::sum(acc, red, phase);   // the bar.sync inside also publishes P
            for (int it = 0; it < 10; it++) {
#pragma unroll
                for (int j = 0; j < K; j++) {
                    acc[j] = 0.f;
                    for (int r = j * T + tid; r < n; r += 1024) {
                        float s = __fadd_rn(0.f, __fmul_rn(S.diag[r], P[r]));
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            const unsigned cm = S.cols[(2 * ch) * np + r], cp = S.cols[(2 * ch + 1) * np + r];
                            if (cm != (unsigned)n) s = __fadd_rn(s, __fmul_rn(S.minus[ch * np + r], P[cm]));
                            if (cp != (unsigned)n) s = __fadd_rn(s, __fmul_rn(S.minus[ch * np + cp], P[cp]));
                        }
                        Tv[r] = s;
                        acc[j] = __fadd_rn(acc[j], __fmul_rn(P[r], s));
                    }
                }
                const float pt2 = Sum::sum(acc, red, phase);
                const float alpha = __fdiv_rn(absNew, pt2);
#pragma unroll
                for (int j = 0; j < K; j++) {
                    acc[j] = 0.f; acc2[j] = 0.f;
                    for (int r = j * T + tid; r < n; r += 1024) {
                        X[r] = __fadd_rn(X[r], __fmul_rn(alpha, P[r]));
                        const float rv = __fsub_rn(R[r], __fmul_rn(alpha, Tv[r]));
                        R[r] = rv;
                        acc[j] = __fadd_rn(acc[j], __fmul_rn(rv, rv));
                        const float zv = __fmul_rn(dinv(r), rv);
                        Tv[r] = zv;
                        acc2[j] = __fadd_rn(acc2[j], __fmul_rn(rv, zv));
                    }
                }
                const float2 both = Sum::sum2(acc, acc2, red, phase);   // all SpMV reads of P are behind this bar.sync
                residualNorm2 = both.x;
                if (residualNorm2 < threshold) break;
                const float absOld = absNew;
                absNew = both.y;
                const float beta = __fdiv_rn(absNew, absOld);
                for (int r = tid; r < n; r += T) P[r] = __fadd_rn(Tv[r], __fmul_rn(beta, P[r]));
                __syncthreads();
            }
        }
    }
    __syncthreads();
}

// the same CG with every vector in registers (at most 1024 rows): thread t owns the rows t, t + T, ... of compact_cg's threads
template <int T, class Sum>
__device__ void compact_cg_regk(const CompactSm& S, float* Psm, float* red, unsigned& phase) {
    constexpr int K = 1024 / T;
    const int n = S.n, np = S.np, tid = threadIdx.x;
    float X[K], R[K], dg[K], di[K], Pv[K], cm[K][3], cp[K][3], s[K], zv[K], acc[K], acc2[K];
    unsigned col[K][6];
    bool on[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int r = j * T + tid;
        on[j] = r < n;
        X[j] = 0.f; R[j] = 0.f; dg[j] = 0.f; di[j] = 1.0f; Pv[j] = 0.f; s[j] = 0.f; zv[j] = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) { col[j][2 * ch] = col[j][2 * ch + 1] = (unsigned)n; cm[j][ch] = cp[j][ch] = 0.f; }
        if (on[j]) {
            R[j] = S.b[r];
            dg[j] = S.diag[r];
            di[j] = dg[j] != 0.f ? __fdiv_rn(1.0f, dg[j]) : 1.0f;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                col[j][2 * ch] = S.cols[(2 * ch) * np + r]; col[j][2 * ch + 1] = S.cols[(2 * ch + 1) * np + r];
                cm[j][ch] = S.minus[ch * np + r];
                cp[j][ch] = S.minus[ch * np + col[j][2 * ch + 1]];
            }
        }
    }
    for (int q = tid; q < np; q += T) Psm[q] = 0.f;
#pragma unroll
    for (int j = 0; j < K; j++) acc[j] = on[j] ? __fadd_rn(0.f, __fmul_rn(R[j], R[j])) : 0.f;
    const float rhsNorm2 = Sum::sum(acc, red, phase);
    if (rhsNorm2 != 0.f) {
        const float tol = 1.1920929e-07f;
        const float threshold = fmaxf(__fmul_rn(__fmul_rn(tol, tol), rhsNorm2), 1.17549435e-38f);
        float residualNorm2 = rhsNorm2;
        if (!(residualNorm2 < threshold)) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                acc[j] = 0.f;
                if (on[j]) { Pv[j] = __fmul_rn(di[j], R[j]); Psm[j * T + tid] = Pv[j]; acc[j] = __fadd_rn(0.f, __fmul_rn(R[j], Pv[j])); }
            }
            float absNew = Sum::sum(acc, red, phase);   // the bar.sync inside also publishes P
            for (int it = 0; it < 10; it++) {
#pragma unroll
                for (int j = 0; j < K; j++) {
                    acc[j] = 0.f;
                    if (on[j]) {
                        float sv = __fadd_rn(0.f, __fmul_rn(dg[j], Pv[j]));
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            if (col[j][2 * ch] != (unsigned)n) sv = __fadd_rn(sv, __fmul_rn(cm[j][ch], Psm[col[j][2 * ch]]));
                            if (col[j][2 * ch + 1] != (unsigned)n) sv = __fadd_rn(sv, __fmul_rn(cp[j][ch], Psm[col[j][2 * ch + 1]]));
                        }
                        s[j] = sv;
                        acc[j] = __fadd_rn(0.f, __fmul_rn(Pv[j], sv));
                    }
                }
                const float pt2 = Sum::sum(acc, red, phase);
                const float alpha = __fdiv_rn(absNew, pt2);
#pragma unroll
                for (int j = 0; j < K; j++) {
                    acc[j] = 0.f; acc2[j] = 0.f;
                    if (on[j]) {
                        X[j] = __fadd_rn(X[j], __fmul_rn(alpha, Pv[j]));
                        R[j] = __fsub_rn(R[j], __fmul_rn(alpha, s[j]));
                        acc[j] = __fadd_rn(0.f, __fmul_rn(R[j], R[j]));
                        zv[j] = __fmul_rn(di[j], R[j]);
                        acc2[j] = __fadd_rn(0.f, __fmul_rn(R[j], zv[j]));
                    }
                }
                const float2 both = Sum::sum2(acc, acc2, red, phase);   // all reads of P are behind this bar.sync
                residualNorm2 = both.x;
                if (residualNorm2 < threshold) break;
                const float absOld = absNew;
                absNew = both.y;
                const float beta = __fdiv_rn(absNew, absOld);
#pragma unroll
                for (int j = 0; j < K; j++)
                    if (on[j]) { Pv[j] = __fadd_rn(zv[j], __fmul_rn(beta, Pv[j])); Psm[j * T + tid] = Pv[j]; }
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int j = 0; j < K; j++) if (on[j]) S.x[j * T + tid] = X[j];
    __syncthreads();
}

// The default coarsest-level CG: threads 0..127 of CTA 0 only (the other warps go on to the cluster barrier), every thread
// sums its rows in ascending order, one butterfly per warp, the four warp partials folded as (p0 + p1) + (p2 + p3). Same
// algorithm and stopping rule as compact_cg (Eigen's ConjugateGradient with the diagonal preconditioner, <= 10 iterations,
// uaamg.cpp:2291-2303), a fixed but different summation order, a quarter of the synchronisation latency.
constexpr int CG_FAST_THREADS = 128;
__device__ __forceinline__ void cg_fast_sync() { bar_named(9, CG_FAST_THREADS); }
__device__ __forceinline__ float cg_fast_sum(float v, float* red, unsigned& phase) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
    float* buf = red + (phase & 1u) * 8;
    phase++;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    cg_fast_sync();
    return __fadd_rn(__fadd_rn(buf[0], buf[1]), __fadd_rn(buf[2], buf[3]));
}
__device__ __forceinline__ float2 cg_fast_sum2(float a, float b, float* red, unsigned& phase) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, d)); b = __fadd_rn(b, __shfl_xor_sync(0xffffffffu, b, d)); }
    float* buf = red + (phase & 1u) * 8;
    phase++;
    if ((threadIdx.x & 31) == 0) { buf[threadIdx.x >> 5] = a; buf[4 + (threadIdx.x >> 5)] = b; }
    cg_fast_sync();
    return make_float2(__fadd_rn(__fadd_rn(buf[0], buf[1]), __fadd_rn(buf[2], buf[3])), __fadd_rn(__fadd_rn(buf[4], buf[5]), __fadd_rn(buf[6], buf[7])));
}
__device__ void compact_cg_fast(const CompactSm& S, float* pt, float* red, unsigned& phase) {
    constexpr int T = CG_FAST_THREADS;
    const int n = S.n, np = S.np, tid = threadIdx.x;
    float* X = S.x; float* R = S.b; float* P = pt; float* Tv = pt + np; float* DI = pt + 2 * np;
    for (int r = tid; r < np; r += T) { X[r] = 0.f; P[r] = 0.f; if (r >= n) R[r] = 0.f; }
    float acc = 0.f, acc2 = 0.f;
    for (int r = tid; r < n; r += T) {
        const float d = S.diag[r];
        DI[r] = d != 0.f ? __fdiv_rn(1.0f, d) : 1.0f;
        acc = __fadd_rn(acc, __fmul_rn(R[r], R[r]));
    }
    const float rhsNorm2 = cg_fast_sum(acc, red, phase);
    if (rhsNorm2 != 0.f) {
        const float tol = 1.1920929e-07f;
        const float threshold = fmaxf(__fmul_rn(__fmul_rn(tol, tol), rhsNorm2), 1.17549435e-38f);
        float residualNorm2 = rhsNorm2;
        if (!(residualNorm2 < threshold)) {
            acc = 0.f;
            for (int r = tid; r < n; r += T) { const float pv = __fmul_rn(DI[r], R[r]); P[r] = pv; acc = __fadd_rn(acc, __fmul_rn(R[r], pv)); }
            float absNew = cg_fast_sum(acc, red, phase);   // the barrier inside also publishes P
            for (int it = 0; it < 10; it++) {
                acc = 0.f;
                for (int r = tid; r < n; r += T) {
                    float s = __fadd_rn(0.f, __fmul_rn(S.diag[r], P[r]));
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const unsigned cm = S.cols[(2 * ch) * np + r], cp = S.cols[(2 * ch + 1) * np + r];
                        if (cm != (unsigned)n) s = __fadd_rn(s, __fmul_rn(S.minus[ch * np + r], P[cm]));
                        if (cp != (unsigned)n) s = __fadd_rn(s, __fmul_rn(S.minus[ch * np + cp], P[cp]));
                    }
                    Tv[r] = s;
                    acc = __fadd_rn(acc, __fmul_rn(P[r], s));
                }
                const float pt2 = cg_fast_sum(acc, red, phase);
                const float alpha = __fdiv_rn(absNew, pt2);
                acc = 0.f; acc2 = 0.f;
                for (int r = tid; r < n; r += T) {
                    X[r] = __fadd_rn(X[r], __fmul_rn(alpha, P[r]));
                    const float rv = __fsub_rn(R[r], __fmul_rn(alpha, Tv[r]));
                    R[r] = rv;
                    acc = __fadd_rn(acc, __fmul_rn(rv, rv));
                    const float zv = __fmul_rn(DI[r], rv);
                    Tv[r] = zv;
                    acc2 = __fadd_rn(acc2, __fmul_rn(rv, zv));
                }
                const float2 both = cg_fast_sum2(acc, acc2, red, phase);   // all reads of P are behind this barrier
                residualNorm2 = both.x;
                if (residualNorm2 < threshold) break;
                const float absOld = absNew;
                absNew = both.y;
                const float beta = __fdiv_rn(absNew, absOld);
                for (int r = tid; r < n; r += T) P[r] = __fadd_rn(Tv[r], __fmul_rn(beta, P[r]));
                cg_fast_sync();
            }
        }
    }
    cg_fast_sync();
}

// what the op loop needs to know about a level, copied to shared memory once (see the note on CCTL.IVALL above)
struct ClLevelS {
    int xbOff, coefOff, metaOff, coefPer, cntN, cntC;
    float def, defInv, diag6;
    const float* diag;
    const uint32_t* assign;
};
__global__ void __launch_bounds__(CL_THREADS, 1) mg_cluster_kernel(const __grid_constant__ ClusterParams P) {
    extern __shared__ __align__(16) unsigned char clsm[];
    __shared__ __align__(16) float red[128];
    __shared__ ClLevelS lvs[CL_MAX_LEVELS];
    __shared__ float sW, sOmw, sAlpha;
    __shared__ int sNOps;
    __shared__ unsigned long long* sTrace;
    const int t = threadIdx.x, g = t >> 6, q = t & 63;
    const uint32_t rank = cl_rank();
    const uint32_t smBase = smem_addr(clsm);
    unsigned phase = 0;
    auto slots_of = [&](int l) -> ClSlot* { return reinterpret_cast<ClSlot*>(clsm + P.lv[l].metaOff); };
    auto xb_of = [&](int l, int i) -> float* { return reinterpret_cast<float*>(clsm + P.lv[l].xbOff + (size_t)i * CL_XB_BYTES); };
    // slots [0, cntN) hold leaves with a coefficient block, [coefPer, coefPer + cntC) leaves with default coefficients
    auto cnt_n = [&](int l) { const ClLevel& L = P.lv[l]; return max(0, min(L.coefPer, L.nN - ((int)rank - 1) * L.coefPer)); };
    auto cnt_c = [&](int l) { const ClLevel& L = P.lv[l]; const int perC = L.xbPer - L.coefPer; return max(0, min(perC, L.nC - ((int)rank - 1) * perC)); };
    uint8_t* prog = clsm + P.progOff;
    const RowIdx rRed = row_idx(q, 0), rBlack = row_idx(q, 1);
    const int X0 = q >> 3, Y0 = q & 7, par = (X0 + Y0) & 1;    // par: colour of this thread's voxel at z = 0
    const int nat0 = (X0 << 6) | (Y0 << 3);                      // natural offset of the z-row
    const uint32_t dump = cl_mapa(smBase + (uint32_t)P.dumpOff, rank);   // where pushes towards a missing neighbour go
    // natural z-row (8 floats) <-> red row, black row
    auto split8 = [&](const float4& a, const float4& b, float4& rr, float4& bb) {
        if (par == 0) { rr = make_float4(a.x, a.z, b.x, b.z); bb = make_float4(a.y, a.w, b.y, b.w); }
        else { bb = make_float4(a.x, a.z, b.x, b.z); rr = make_float4(a.y, a.w, b.y, b.w); }
    };
    auto merge8 = [&](const float4& rr, const float4& bb, float4& a, float4& b) {
        if (par == 0) { a = make_float4(rr.x, bb.x, rr.y, bb.y); b = make_float4(rr.z, bb.z, rr.w, bb.w); }
        else { a = make_float4(bb.x, rr.x, bb.y, rr.y); b = make_float4(bb.z, rr.z, bb.w, rr.w); }
    };

    // ---------------- staging
    if (P.trace && rank == 0 && t == 0) P.trace[2 * P.nOps + 2] = globaltimer();
    if (t < P.nLevels) {
        const ClLevel& L = P.lv[t];
        ClLevelS& S = lvs[t];
        S.xbOff = L.xbOff; S.coefOff = L.coefOff; S.metaOff = L.metaOff; S.coefPer = L.coefPer;
        S.cntN = rank > 0 ? cnt_n(t) : 0; S.cntC = rank > 0 ? cnt_c(t) : 0;
        S.def = -L.v.term; S.defInv = __fdiv_rn(1.0f, __fmul_rn(6.0f, L.v.term)); S.diag6 = __fmul_rn(6.0f, L.v.term);
        S.diag = L.v.diag; S.assign = L.assign;
    }
    if (t == 0) { sW = P.w; sOmw = P.oneMinusW; sAlpha = P.prolongAlpha; sNOps = P.nOps; sTrace = P.trace; }
    for (int i = t; i < P.nOps; i += CL_THREADS) prog[i] = __ldg(&P.prog[i]);
    if (rank > 0) {
        {   // x = b = halos = 0 everywhere
            float4* z4 = reinterpret_cast<float4*>(clsm);
            for (int i = t; i < (P.xbBytes >> 4); i += CL_THREADS) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // this CTA's slot records: one thread per (slot, field group)
        for (int l = 0; l < P.nLevels; l++) {
            const ClLevel& L = P.lv[l];
            ClSlot* slots = slots_of(l);
            const ClMeta* src = L.meta + (size_t)rank * L.xbPer;
            for (int u = t; u < L.xbPer * 32; u += CL_THREADS) {
                ClSlot& d = slots[u >> 5];
                const ClMeta& m = src[u >> 5];
                const int f = u & 31;
                const int leaf = m.leaf;
                if (f < 6) { const uint32_t a = m.nb[f]; d.nb[f] = (leaf < 0 || a == CL_NONE) ? dump : cl_xb_addr(L, smBase, a); }
                else if (f == 6) {
                    d.parent = (leaf < 0 || m.parent == CL_NONE || l + 1 >= P.nLevels) ? 0u : cl_xb_addr(P.lv[l + 1], smBase, m.parent);
                    d.parentOff = m.parentOff; d.leaf = leaf; d.flags = m.flags; d.pad = 0;
                    d.coef = (leaf < 0 || m.coef < 0) ? -1 : L.coefOff + m.coef * CL_COEF_BYTES;
                } else if (f < 15) { d.mask[f - 7] = leaf < 0 ? 0ull : m.mask[f - 7]; }
                else if (f < 23) { d.parentMask[f - 15] = leaf < 0 ? 0ull : m.parentMask[f - 15]; }
                else if (f < 31) {   // colour-split DOF bits of the x slice f - 23, both colours
                    const int X = f - 23;
                    const uint64_t wd = leaf < 0 ? 0ull : m.mask[X];
                    uint32_t rb = 0, bb = 0;
                    for (int Y = 0; Y < 8; Y++)
                        for (int k = 0; k < 4; k++) {
                            const int p = (X + Y) & 1;
                            rb |= (uint32_t)((wd >> ((Y << 3) | (2 * k + p))) & 1ull) << ((Y << 2) | k);
                            bb |= (uint32_t)((wd >> ((Y << 3) | (2 * k + (p ^ 1)))) & 1ull) << ((Y << 2) | k);
                        }
                    d.cmask[X] = rb; d.cmask[8 + X] = bb;
                }
            }
        }
        __syncthreads();
        // coefficient blocks of the leaves that have one: eight leaves at a time, a z-row per thread
        for (int l = 0; l < P.nLevels; l++) {
            const ClLevel& L = P.lv[l];
            const float def = -L.v.term;
            const ClSlot* slots = slots_of(l);
            const int cn = cnt_n(l);
            for (int i = g; i < cn; i += CL_GROUPS) {
                const ClSlot& m = slots[i];
                float* cb = reinterpret_cast<float*>(clsm + m.coef);
                const size_t gb = (size_t)m.leaf * LEAF + nat0;
                const float* src[4] = {L.v.invdiag, L.v.xe, L.v.ye, L.v.ze};
                float4 lo[4], hi[4];
#pragma unroll
                for (int a = 0; a < 4; a++) { lo[a] = __ldg(reinterpret_cast<const float4*>(src[a] + gb)); hi[a] = __ldg(reinterpret_cast<const float4*>(src[a] + gb + 4)); }
                float fv[3];
                const LeafInfo* li = L.v.info + m.leaf;
#pragma unroll
                for (int f = 0; f < 3; f++) {   // entry q of the +x / +y / +z neighbour's xe / ye / ze face
                    const int nb = f == 0 ? li->nb[1] : (f == 1 ? li->nb[3] : li->nb[5]);
                    const float* arr = f == 0 ? L.v.xe : (f == 1 ? L.v.ye : L.v.ze);
                    const int noff = f == 0 ? q : (f == 1 ? ((X0 << 6) | Y0) : ((X0 << 6) | (Y0 << 3)));
                    fv[f] = nb >= 0 ? __ldg(&arr[(size_t)nb * LEAF + noff]) : def;
                }
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    float4 rr, bb;
                    split8(lo[a], hi[a], rr, bb);
                    *reinterpret_cast<float4*>(cb + a * 512 + (X0 << 5) + (Y0 << 2)) = rr;
                    *reinterpret_cast<float4*>(cb + a * 512 + 256 + (X0 << 5) + (Y0 << 2)) = bb;
                }
                // face entry (a0, b0) = (q >> 3, q & 7): colour (a0 + b0) & 1, index a0 << 2 | b0 >> 1
#pragma unroll
                for (int f = 0; f < 3; f++) cb[2048 + f * 64 + par * 32 + ((X0 << 2) | (Y0 >> 1))] = fv[f];
            }
        }
        // the top level's right-hand side (and iterate)
        {
            const ClLevel& L = P.lv[0];
            const ClSlot* slots = slots_of(0);
            for (int i = g; i < L.xbPer; i += CL_GROUPS) {
                const ClSlot& m = slots[i];
                if (m.leaf < 0) continue;
                float* xb = xb_of(0, i);
                const size_t gb = (size_t)m.leaf * LEAF + nat0;
                float4 rr, bb;
                split8(*reinterpret_cast<const float4*>(P.topB + gb), *reinterpret_cast<const float4*>(P.topB + gb + 4), rr, bb);
                *reinterpret_cast<float4*>(xb + 512 + rRed.own) = rr; *reinterpret_cast<float4*>(xb + 512 + rBlack.own) = bb;
                if (P.loadX) {
                    split8(*reinterpret_cast<const float4*>(P.topX + gb), *reinterpret_cast<const float4*>(P.topX + gb + 4), rr, bb);
                    *reinterpret_cast<float4*>(xb + rRed.own) = rr; *reinterpret_cast<float4*>(xb + rBlack.own) = bb;
                }
            }
        }
    } else {
        // CTA 0: the coarsest level's compact rows
        const CompactDev& D = P.cg;
        const BlobLayout B = blob_layout(D.np, D.hasChild != 0);
        unsigned char* dst = clsm + P.cgOff;
        const int src[3] = {B.diag, B.minus, B.cols};
        const int len[3] = {4 * D.np, 12 * D.np, 12 * D.np};
        int o = 0;
        for (int u = 0; u < 3; u++) {
            const uint4* sp = reinterpret_cast<const uint4*>(D.blob + src[u]);
            uint4* dp = reinterpret_cast<uint4*>(dst + o);
            for (int i = t; i < (len[u] >> 4); i += CL_THREADS) dp[i] = __ldg(&sp[i]);
            o += len[u];
        }
    }
    cl_sync();
    if (P.loadX) {   // halos of the iterate that was loaded (every CTA's zero fill is behind the barrier above)
        if (rank > 0) {
            const ClLevel& L = P.lv[0];
            const ClSlot* slots = slots_of(0);
            for (int i = g; i < L.xbPer; i += CL_GROUPS) {
                const ClSlot& m = slots[i];
                if (m.leaf < 0) continue;
                const float* xb = xb_of(0, i);
                push_row(m, rRed, ld4(xb + rRed.own)); push_row(m, rBlack, ld4(xb + rBlack.own));
            }
        }
        cl_sync();
    }
    if (P.trace && rank == 0 && t == 0) P.trace[2 * P.nOps + 3] = globaltimer();

    // ---------------- the op list
    float* sres = reinterpret_cast<float*>(clsm + P.sresOff) + g * LEAF;
    const int cgOff = P.cgOff, cgCompat = P.cgCompat;
    const CompactDev cgD = P.cg;
    const int nOps = sNOps;
    const int dbg = P.dbg;
    const float w = sW, oneMinusW = sOmw, prolongAlpha = sAlpha;
    unsigned long long* const trace = sTrace;
    for (int k = 0; k < nOps; k++) {
        long long c0 = 0, c1 = 0, c2 = 0;
        if (trace) c0 = clock64();
        const int code = prog[k] & 7, l = prog[k] >> 3;
        const ClLevelS L = lvs[l];
        if (trace && rank == 0 && t == 0) trace[k] = globaltimer();
        if (trace) c1 = clock64();
        if (code == OP_COARSE) {
            if (rank == 0) {
                const CompactDev& D = cgD;
                unsigned char* base = clsm + cgOff;
                CompactSm S;
                S.n = D.n; S.np = D.np; S.nRed = D.nRed;
                S.diag = reinterpret_cast<const float*>(base);
                S.minus = reinterpret_cast<const float*>(base + 4 * D.np);
                S.cols = reinterpret_cast<const uint16_t*>(base + 16 * D.np);
                S.inv = S.diag; S.parent = nullptr; S.child = nullptr;
                S.x = reinterpret_cast<float*>(base + 28 * D.np);
                S.b = S.x + D.np;
                float* pt = S.b + D.np;
                for (int r = t; r < D.n; r += CL_THREADS) {
                    const uint32_t v = __ldg(&D.voxelOfRow[r]);
                    const uint32_t a = __ldg(&L.assign[v >> 9]);
                    S.b[r] = cl_ld(cl_xb_addr_s(L.xbOff, smBase, a) + (uint32_t)(512 + split_idx((int)(v & 511u))) * 4u);
                }
                __syncthreads();
                const long long g1 = trace ? clock64() : 0;
                if (cgCompat) {
                    if (D.n <= 1024) compact_cg_regk<CL_THREADS, CgSum<CL_THREADS>>(S, pt, red, phase);
                    else compact_cg_k<CL_THREADS>(S, pt, red, phase);
                } else if (D.n <= 1024) {
                    // every vector of the solve in registers, all 16 warps, three bar.sync per iteration
                    compact_cg_regk<CL_THREADS, CgSumFast<CL_THREADS>>(S, pt, red, phase);
                } else {
                    if (t < CG_FAST_THREADS) compact_cg_fast(S, pt, red, phase);
                    __syncthreads();
                }
                const long long g2 = trace ? clock64() : 0;
                for (int r = t; r < D.n; r += CL_THREADS) {
                    const uint32_t v = __ldg(&D.voxelOfRow[r]);
                    const uint32_t a = __ldg(&L.assign[v >> 9]);
                    cl_st(cl_xb_addr_s(L.xbOff, smBase, a) + (uint32_t)split_idx((int)(v & 511u)) * 4u, S.x[r]);
                }
                if (trace && t == 0) {
                    trace[2 * nOps + 4 + 3 * k] = (unsigned long long)(g1 - c1);
                    trace[2 * nOps + 4 + 3 * k + 1] = (unsigned long long)(g2 - g1);
                    trace[2 * nOps + 4 + 3 * k + 2] = (unsigned long long)(clock64() - g2);
                }
            }
        } else if (rank > 0) {
            const float def = L.def, defInv = L.defInv;
            const ClSlot* slots = reinterpret_cast<const ClSlot*>(clsm + L.metaOff);
            const int cn = L.cntN, cc = L.cntC;
            if ((code == OP_RED || code == OP_BLACK) && (dbg & 2)) {
                // timing experiment: barrier only
            } else if (code == OP_RED || code == OP_BLACK) {
                const RowIdx& r = code == OP_RED ? rRed : rBlack;
                colour_pass<true>(clsm, L.xbOff, slots, 0, cn, g, r, def, defInv, w, oneMinusW, dbg);
                colour_pass<false>(clsm, L.xbOff, slots, L.coefPer, cc, g, r, def, defInv, w, oneMinusW, dbg);
            } else if (code == OP_ZERO_RED) {
                for (int u = g; u < cn + cc; u += CL_GROUPS) {
                    const int i = u < cn ? u : L.coefPer + u - cn;
                    const ClSlot& m = slots[i];
                    float* xb = reinterpret_cast<float*>(clsm + L.xbOff + (size_t)i * CL_XB_BYTES);
                    const float4 inv = m.coef >= 0 ? ld4(reinterpret_cast<const float*>(clsm + m.coef) + rRed.own) : make_float4(defInv, defInv, defInv, defInv);
                    const float4 bi = ld4(xb + 512 + rRed.own);
                    const uint32_t on = (m.cmask[rRed.word] >> rRed.shift) & 15u;
                    float4 rv;
                    rv.x = (on & 1u) ? __fmul_rn(__fmul_rn(bi.x, inv.x), w) : 0.f;
                    rv.y = (on & 2u) ? __fmul_rn(__fmul_rn(bi.y, inv.y), w) : 0.f;
                    rv.z = (on & 4u) ? __fmul_rn(__fmul_rn(bi.z, inv.z), w) : 0.f;
                    rv.w = (on & 8u) ? __fmul_rn(__fmul_rn(bi.w, inv.w), w) : 0.f;
                    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(xb + rRed.own) = rv; *reinterpret_cast<float4*>(xb + rBlack.own) = zero;
                    push_row(m, rRed, rv); push_row(m, rBlack, zero);
                }
            } else if (code == OP_RESID_RESTRICT) {
                const int rounds = (cn + cc + CL_GROUPS - 1) / CL_GROUPS;
                for (int rd = 0; rd < rounds; rd++) {
                    const int u = rd * CL_GROUPS + g;
                    const bool have = u < cn + cc;   // uniform over the group of 64
                    const int i = have ? (u < cn ? u : L.coefPer + u - cn) : 0;
                    const ClSlot& m = slots[i];
                    if (have) {
                        const float* xb = reinterpret_cast<const float*>(clsm + L.xbOff + (size_t)i * CL_XB_BYTES);
                        const float* cb = m.coef >= 0 ? reinterpret_cast<const float*>(clsm + m.coef) : nullptr;
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const RowIdx& r = c == 0 ? rRed : rBlack;
                            const uint32_t on = (m.cmask[r.word] >> r.shift) & 15u;
                            const float4 od = cb ? row_offdiag<true>(xb, cb, r, def) : row_offdiag<false>(xb, cb, r, def);
                            const float4 xi = ld4(xb + r.own), bi = ld4(xb + 512 + r.own);
                            const float odv[4] = {od.x, od.y, od.z, od.w}, xv[4] = {xi.x, xi.y, xi.z, xi.w}, bv[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                const int nat = nat0 + 2 * kk + r.p;
                                float res = 0.f;
                                if ((on >> kk) & 1u) {
                                    const float dg = (m.flags & LI_DIAG) ? L.diag6 : __ldg(&L.diag[(size_t)m.leaf * LEAF + nat]);
                                    res = __fsub_rn(bv[kk], __fmaf_rn(xv[kk], dg, odv[kk]));
                                }
                                sres[nat] = res;
                            }
                        }
                    }
                    bar_named(1 + g, 64);
                    if (have && m.parent) {
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
                                    const int fo = fb + 64 * ii + 8 * jj + kk;
                                    if (meta_bit(m.mask, fo)) { sum = __fadd_rn(sum, sres[fo]); any = true; }
                                }
                        if (any) cl_st(m.parent + (uint32_t)(512 + split_idx((int)m.parentOff + ((cx << 6) | (cy << 3) | cz))) * 4u, __fmul_rn(sum, 0.125f));
                    }
                    bar_named(1 + g, 64);
                }
            } else if (code == OP_PROLONG) {
                for (int u = g; u < cn + cc; u += CL_GROUPS) {
                    const int i = u < cn ? u : L.coefPer + u - cn;
                    const ClSlot& m = slots[i];
                    if (!m.parent) continue;
                    float* xb = reinterpret_cast<float*>(clsm + L.xbOff + (size_t)i * CL_XB_BYTES);
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const RowIdx& r = c == 0 ? rRed : rBlack;
                        const uint32_t on = (m.cmask[r.word] >> r.shift) & 15u;
                        const float4 xi = ld4(xb + r.own);
                        float xv[4] = {xi.x, xi.y, xi.z, xi.w};
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            if (!((on >> kk) & 1u)) continue;
                            const int Z = 2 * kk + r.p;
                            const int co = (int)m.parentOff + (((X0 >> 1) << 6) | ((Y0 >> 1) << 3) | (Z >> 1));
                            if (!meta_bit(m.parentMask, co)) continue;
                            xv[kk] = __fadd_rn(xv[kk], __fmul_rn(prolongAlpha, cl_ld(m.parent + (uint32_t)split_idx(co) * 4u)));
                        }
                        const float4 nv = make_float4(xv[0], xv[1], xv[2], xv[3]);
                        *reinterpret_cast<float4*>(xb + r.own) = nv;
                        push_row(m, r, nv);
                    }
                }
            }
        }
        if (trace && rank == 1 && t == 0) trace[nOps + 1 + k] = globaltimer();
        if (trace) c2 = clock64();
        cl_sync();
        if (trace && rank == 1 && t == 0 && code != OP_COARSE) {
            const long long c3 = clock64();
            trace[2 * nOps + 4 + 3 * k] = (unsigned long long)(c1 - c0);
            trace[2 * nOps + 4 + 3 * k + 1] = (unsigned long long)(c2 - c1);
            trace[2 * nOps + 4 + 3 * k + 2] = (unsigned long long)(c3 - c2);
        }
    }
    if (trace && rank == 0 && t == 0) trace[nOps] = globaltimer();
    // ---------------- the top level's iterate back to global memory
    if (rank > 0) {
        const ClLevel& L = P.lv[0];
        const ClSlot* slots = slots_of(0);
        for (int i = g; i < L.xbPer; i += CL_GROUPS) {
            const ClSlot& m = slots[i];
            if (m.leaf < 0) continue;
            const float* xb = xb_of(0, i);
            float4 a, b;
            merge8(ld4(xb + rRed.own), ld4(xb + rBlack.own), a, b);
            float* dst = P.topX + (size_t)m.leaf * LEAF + nat0;
            *reinterpret_cast<float4*>(dst) = a; *reinterpret_cast<float4*>(dst + 4) = b;
        }
    }
}
