// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FastFLIP hot path.
// Nothing under oracle/ is linked, imported or executed by the product
// (zeno_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker.
//
// sgrid.h: a minimal sparse 8^3-leaf grid with OpenVDB value/mask semantics
// (what the reference gets from openvdb::tree::Tree4<T,5,4,3>):
//   * leaf = 512 voxels, offset = x<<6 | y<<3 | z
//     (projects/zenvdb/openvdb/openvdb/openvdb/tree/LeafNode.h:1051-1057)
//   * every voxel of an allocated leaf has a value and an active bit; a voxel
//     in no leaf reads as (background, inactive)
//   * dilateActiveValues(n, NN_FACE | NN_FACE_EDGE_VERTEX) as in
//     openvdb/tools/Morphology.h:58,80,1055-1067 (no tiles exist on this path)
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>
#include <algorithm>

namespace orc {

struct Coord {
    int x, y, z;
    Coord() : x(0), y(0), z(0) {}
    Coord(int a, int b, int c) : x(a), y(b), z(c) {}
    int& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    int operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Coord offsetBy(int a, int b, int c) const { return Coord(x + a, y + b, z + c); }
    bool operator==(const Coord& o) const { return x == o.x && y == o.y && z == o.z; }
};

// 21 bits per axis of the leaf coordinate (voxel >> 3), biased to be non-negative.
inline uint64_t leafKeyOf(int vx, int vy, int vz) {
    const uint64_t B = 1u << 20;
    uint64_t lx = uint64_t(int64_t(vx >> 3) + int64_t(B));
    uint64_t ly = uint64_t(int64_t(vy >> 3) + int64_t(B));
    uint64_t lz = uint64_t(int64_t(vz >> 3) + int64_t(B));
    return (lx << 42) | (ly << 21) | lz;
}
inline int voxelOffset(int vx, int vy, int vz) {
    return ((vx & 7) << 6) | ((vy & 7) << 3) | (vz & 7);
}

using Mask512 = std::array<uint64_t, 8>;
inline bool maskGet(const Mask512& m, int off) { return (m[off >> 6] >> (off & 63)) & 1u; }
inline void maskSet(Mask512& m, int off, bool on) {
    if (on) m[off >> 6] |= (uint64_t(1) << (off & 63));
    else m[off >> 6] &= ~(uint64_t(1) << (off & 63));
}
inline int maskCount(const Mask512& m) {
    int c = 0;
    for (int i = 0; i < 8; i++) c += __builtin_popcountll(m[i]);
    return c;
}

// Topology shared helper: set of leaves keyed by origin.
template <int NC>
struct Grid {
    float bg[NC];
    std::unordered_map<uint64_t, int> dir;
    std::vector<Coord> origins;
    std::vector<Mask512> masks;
    std::vector<float> vals;  // [leaf][channel][512]

    Grid() { for (int c = 0; c < NC; c++) bg[c] = 0.f; }
    explicit Grid(float b) { for (int c = 0; c < NC; c++) bg[c] = b; }

    int leafCount() const { return int(origins.size()); }
    void clear() { dir.clear(); origins.clear(); masks.clear(); vals.clear(); }

    int findLeaf(int vx, int vy, int vz) const {
        auto it = dir.find(leafKeyOf(vx, vy, vz));
        return it == dir.end() ? -1 : it->second;
    }
    int findLeaf(const Coord& c) const { return findLeaf(c.x, c.y, c.z); }

    // touchLeaf: new leaves are filled with background, all inactive.
    int touchLeaf(int vx, int vy, int vz) {
        uint64_t k = leafKeyOf(vx, vy, vz);
        auto it = dir.find(k);
        if (it != dir.end()) return it->second;
        int id = int(origins.size());
        dir.emplace(k, id);
        origins.push_back(Coord(vx & ~7, vy & ~7, vz & ~7));
        Mask512 m; m.fill(0);
        masks.push_back(m);
        size_t base = vals.size();
        vals.resize(base + size_t(NC) * 512);
        for (int c = 0; c < NC; c++)
            std::fill(vals.begin() + base + c * 512, vals.begin() + base + (c + 1) * 512, bg[c]);
        return id;
    }
    int touchLeaf(const Coord& c) { return touchLeaf(c.x, c.y, c.z); }

    float* leafVals(int leaf, int c = 0) { return &vals[(size_t(leaf) * NC + c) * 512]; }
    const float* leafVals(int leaf, int c = 0) const { return &vals[(size_t(leaf) * NC + c) * 512]; }

    float get(int c, int vx, int vy, int vz) const {
        int l = findLeaf(vx, vy, vz);
        if (l < 0) return bg[c];
        return leafVals(l, c)[voxelOffset(vx, vy, vz)];
    }
    float get(int c, const Coord& p) const { return get(c, p.x, p.y, p.z); }
    float get(const Coord& p) const { return get(0, p.x, p.y, p.z); }
    bool isOn(int vx, int vy, int vz) const {
        int l = findLeaf(vx, vy, vz);
        if (l < 0) return false;
        return maskGet(masks[l], voxelOffset(vx, vy, vz));
    }
    bool isOn(const Coord& p) const { return isOn(p.x, p.y, p.z); }
    void setOn(int c, const Coord& p, float v) {
        int l = touchLeaf(p);
        int off = voxelOffset(p.x, p.y, p.z);
        leafVals(l, c)[off] = v;
        maskSet(masks[l], off, true);
    }
    void activate(const Coord& p) {
        int l = touchLeaf(p);
        maskSet(masks[l], voxelOffset(p.x, p.y, p.z), true);
    }
    uint64_t activeCount() const {
        uint64_t n = 0;
        for (auto& m : masks) n += maskCount(m);
        return n;
    }

    // openvdb TopologyCopy constructor: same leaves + active masks, every value = background.
    template <int NC2>
    void topologyCopyFrom(const Grid<NC2>& o) {
        dir = o.dir;
        origins = o.origins;
        masks = o.masks;
        vals.resize(origins.size() * size_t(NC) * 512);
        for (size_t l = 0; l < origins.size(); l++)
            for (int c = 0; c < NC; c++)
                std::fill(vals.begin() + (l * NC + c) * 512, vals.begin() + (l * NC + c + 1) * 512, bg[c]);
    }

    // openvdb Tree::topologyUnion(other): activates voxels active in other, adds missing leaves.
    template <int NC2>
    void topologyUnion(const Grid<NC2>& o) {
        for (int l = 0; l < o.leafCount(); l++) {
            int m = touchLeaf(o.origins[l]);
            for (int w = 0; w < 8; w++) masks[m][w] |= o.masks[l][w];
        }
    }
    // openvdb Tree::topologyDifference(other): deactivates voxels active in other.
    template <int NC2>
    void topologyDifference(const Grid<NC2>& o) {
        for (int l = 0; l < o.leafCount(); l++) {
            int m = findLeaf(o.origins[l]);
            if (m < 0) continue;
            for (int w = 0; w < 8; w++) masks[m][w] &= ~o.masks[l][w];
        }
    }

    // openvdb::tools::dilateActiveValues(tree, iterations, nn, EXPAND_TILES)
    // nn26 = NN_FACE_EDGE_VERTEX, else NN_FACE. Newly allocated leaves hold background.
    void dilate(int iterations, bool nn26) {
        for (int it = 0; it < iterations; it++) {
            std::vector<Mask512> old = masks;
            int nOld = leafCount();
            for (int l = 0; l < nOld; l++) {
                const Mask512& m = old[l];
                bool any = false;
                for (int w = 0; w < 8; w++) any |= (m[w] != 0);
                if (!any) continue;
                Coord o = origins[l];
                for (int off = 0; off < 512; off++) {
                    if (!maskGet(m, off)) continue;
                    int x = off >> 6, y = (off >> 3) & 7, z = off & 7;
                    bool interior = x > 0 && x < 7 && y > 0 && y < 7 && z > 0 && z < 7;
                    if (nn26) {
                        for (int dx = -1; dx <= 1; dx++)
                            for (int dy = -1; dy <= 1; dy++)
                                for (int dz = -1; dz <= 1; dz++) {
                                    if (interior) maskSet(masks[l], ((x + dx) << 6) | ((y + dy) << 3) | (z + dz), true);
                                    else activate(Coord(o.x + x + dx, o.y + y + dy, o.z + z + dz));
                                }
                    } else {
                        static const int d6[6][3] = {{-1,0,0},{1,0,0},{0,-1,0},{0,1,0},{0,0,-1},{0,0,1}};
                        for (int k = 0; k < 6; k++) {
                            if (interior) maskSet(masks[l], ((x + d6[k][0]) << 6) | ((y + d6[k][1]) << 3) | (z + d6[k][2]), true);
                            else activate(Coord(o.x + x + d6[k][0], o.y + y + d6[k][1], o.z + z + d6[k][2]));
                        }
                    }
                }
            }
        }
    }
};

using FloatGrid = Grid<1>;
using Vec3Grid = Grid<3>;

// The particle store: per leaf 512 cumulative end offsets in voxel
// order, attribute arrays P (3 x u16 fixed point) and v (3 x fp16 bits), voxel-sorted.
struct Points {
    std::unordered_map<uint64_t, int> dir;
    std::vector<Coord> origins;
    std::vector<std::array<uint32_t, 512>> voxelEnd;  // cumulative within the leaf
    std::vector<uint64_t> leafBegin;                   // global index of the leaf's first particle (size nLeaves+1)
    std::vector<uint16_t> P;                           // [N][3]
    std::vector<uint16_t> v;                           // [N][3] fp16 bits
    int leafCount() const { return int(origins.size()); }
    size_t size() const { return P.size() / 3; }
    int findLeaf(int vx, int vy, int vz) const {
        auto it = dir.find(leafKeyOf(vx, vy, vz));
        return it == dir.end() ? -1 : it->second;
    }
    void clear() { dir.clear(); origins.clear(); voxelEnd.clear(); leafBegin.clear(); P.clear(); v.clear(); }
};

}  // namespace orc
