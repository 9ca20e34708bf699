// libflipb200 -- shared device/host definitions.
// Data model (DESIGN.md "Data layout in HBM"):
//   Topo   : a set of 8^3 leaves addressed through a DENSE leaf directory over the bounding
//            box of the leaf coordinates (slot = dir[linear leaf coord], -1 = no leaf). Slots
//            are numbered in lexicographic (x,y,z) leaf order, so every leaf list is sorted
//            and every reduction over leaves has a fixed order.
//   GridF/GridV: values [leaf][512] fp32 per channel (SoA) + one 512-bit active mask per leaf,
//            OpenVDB semantics (a voxel in no leaf reads as background / inactive).
//   Particles: SoA of three u32 words per particle (Px|Py<<16, Pz|vx<<16, vy|vz<<16) holding
//            the reference's 12-byte quantised state (fxpt16 position, fp16 velocity), sorted
//            by (leaf slot, voxel offset); voxelStart[slot*512+off] = global exclusive prefix.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>

namespace fb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            throw fb::Error(2, std::string(#call) + ": " + cudaGetErrorString(e_) + " at " +  \
                                   __FILE__ + ":" + std::to_string(__LINE__));                 \
    } while (0)

#define FB_REQUIRE(cond, code, msg)                   \
    do {                                              \
        if (!(cond)) throw fb::Error((code), (msg));  \
    } while (0)

#define FB_STR2(x) #x
#define FB_STR(x) FB_STR2(x)

constexpr int LEAF = 512;

// ------------------------------------------------------------------ device topology view
struct TopoView {
    int n;            // leaves
    int3 dmin;        // directory minimum, in leaf coordinates (voxel >> 3)
    int3 ddim;        // directory extent
    const int* dir;   // [ddim.x*ddim.y*ddim.z] slot or -1
    const int3* origin;  // [n] voxel coordinates of the leaf origin
    const int* nbr27;    // [n*27] slot of leaf + (i-1,j-1,k-1) leaves, index i*9+j*3+k, -1 = none
};

__host__ __device__ inline int floordiv8(int v) { return v >> 3; }

__device__ __forceinline__ int topo_find(const TopoView& t, int vx, int vy, int vz) {
    int lx = (vx >> 3) - t.dmin.x, ly = (vy >> 3) - t.dmin.y, lz = (vz >> 3) - t.dmin.z;
    if ((unsigned)lx >= (unsigned)t.ddim.x || (unsigned)ly >= (unsigned)t.ddim.y ||
        (unsigned)lz >= (unsigned)t.ddim.z)
        return -1;
    return __ldg(&t.dir[(lx * t.ddim.y + ly) * t.ddim.z + lz]);
}
__device__ __forceinline__ int voxel_off(int vx, int vy, int vz) {
    return ((vx & 7) << 6) | ((vy & 7) << 3) | (vz & 7);
}
__device__ __forceinline__ bool mask_get(const uint64_t* m, int leaf, int off) {
    return (m[(size_t)leaf * 8 + (off >> 6)] >> (off & 63)) & 1ull;
}

// accessor-style reads (value regardless of the active bit; background where no leaf)
__device__ __forceinline__ float grid_get(const TopoView& t, const float* val, float bg, int vx, int vy, int vz) {
    int l = topo_find(t, vx, vy, vz);
    if (l < 0) return bg;
    return __ldg(&val[(size_t)l * LEAF + voxel_off(vx, vy, vz)]);
}
__device__ __forceinline__ bool grid_on(const TopoView& t, const uint64_t* mask, int vx, int vy, int vz) {
    int l = topo_find(t, vx, vy, vz);
    if (l < 0) return false;
    return mask_get(mask, l, voxel_off(vx, vy, vz));
}

// ------------------------------------------------------------------ codecs (device)
// FixedPointCodec<false, PositionRange> (openvdb/points/AttributeArray.h:47-65,953-976)
__device__ __forceinline__ float fx_decode(uint32_t u) { return __fsub_rn(__fdiv_rn((float)u, 65535.0f), 0.5f); }
// The same value without the IEEE-division sequence: q = u * fl(1/65535), one exact remainder and one correction FMA.
// Equal to fx_decode for every one of the 65536 codes (checked exhaustively in exact arithmetic,
// tests/test_oracle_cpu.py::test_fx_decode_fast_is_exact; on the device by every bit-exact P2G parity test).
__device__ __forceinline__ float fx_decode_fast(uint32_t u) {
    const float r = 0x1.0001p-16f;   // fl(1/65535)
    const float uf = (float)u;
    const float q = __fmul_rn(uf, r);
    const float rem = __fmaf_rn(-q, 65535.0f, uf);
    return __fsub_rn(__fmaf_rn(rem, r, q), 0.5f);
}
__device__ __forceinline__ uint32_t fx_encode(float p) {
    float s = __fadd_rn(p, 0.5f);
    if (0.0f > s) return 0u;
    else if (1.0f <= s) return 65535u;
    return (uint32_t)(uint16_t)(__fmul_rn(s, 65535.0f));
}
// TruncateCodec on half: IEEE round-to-nearest-even (openvdb/math/Half.h:430-490)
__device__ __forceinline__ float h_decode(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }
__device__ __forceinline__ uint32_t h_encode(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }

// ------------------------------------------------------------------ host side containers
struct Ctx;

template <typename T>
struct DBuf {  // stream-ordered device buffer
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DBuf() {}
    DBuf(size_t count, cudaStream_t st) { alloc(count, st); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { release(); }
    void alloc(size_t count, cudaStream_t st) {
        release();
        s = st; n = count;
        if (count) FB_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), st));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr; n = 0;
    }
    void zero() { if (n) FB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void fill_bytes(int b) { if (n) FB_CUDA(cudaMemsetAsync(p, b, n * sizeof(T), s)); }
};

struct Topo {
    int n = 0;
    int3 dmin{0, 0, 0}, ddim{0, 0, 0};
    DBuf<int> dir;
    DBuf<int3> origin;
    DBuf<int> nbr27;
    uint64_t epoch = 0;
    TopoView view() const { return TopoView{n, dmin, ddim, dir.p, origin.p, nbr27.p}; }
};
using TopoPtr = std::shared_ptr<Topo>;

struct GridF {
    TopoPtr topo;
    DBuf<float> val;      // [n][512]
    DBuf<uint64_t> mask;  // [n][8]
    DBuf<uint8_t> alloc;  // [n] 1 = this leaf exists in the VDB tree the grid stands for
    float bg = 0.f;
    int leaves() const { return topo ? topo->n : 0; }
};
struct GridV {
    TopoPtr topo;
    DBuf<float> val[3];   // channel SoA
    DBuf<uint64_t> mask;  // union mask, the Vec3fGrid's mask
    float bg[3] = {0.f, 0.f, 0.f};
    int leaves() const { return topo ? topo->n : 0; }
};

struct Particles {
    TopoPtr topo;               // same object as the pool the store was binned on
    uint64_t n = 0;
    DBuf<uint32_t> w0, w1, w2;  // packed state
    DBuf<uint32_t> voxelStart;  // [topo->n*512 + 1] global exclusive prefix
};

}  // namespace fb
