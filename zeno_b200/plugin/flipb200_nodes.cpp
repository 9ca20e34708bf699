// flipb200_nodes.cpp -- the Zeno side of the drop-in: the FastFLIP hot-path nodes re-registered under their
// reference names with their reference sockets and params, whose apply() bodies flatten the
// OpenVDB leaves of their socket objects and call libflipb200's C ABI (include/flipb200.h) instead of the
// FLIP_vdb statics. Built INTO the zeno target in place of the listed projects/FastFLIP/nosys/*.cpp files when
// ZENO_WITH_FastFLIP_B200 is ON; it needs only the headers Zeno already has (zeno core,
// zenvdb's VDBGrid.h, OpenVDB) plus flipb200.h. Host language = the reference's (C++17); no CUDA or torch types.
//
// replaces: nosys/P2G.cpp (FLIP_P2G), nosys/SheetG2PAdvector.cpp (G2PAdvectorSheetty),
//           nosys/SolvePoissonPressureEqn.cpp (AssembleSolvePPE), nosys/SubtractPressureGradient.cpp,
//           nosys/EvalFaceWeight.cpp (CutCellWeight), nosys/FixLiquidSDF.cpp (PushOutLiquidSDF),
//           nosys/FieldAddVector.cpp, nosys/CFL.cpp (CFL_dt; SurfaceTension_dt is host arithmetic and is kept).
//
// State model: one device world per FLIP world of the graph. A node finds its world through the socket OBJECTS it is handed
// (the IObject wrappers SetFLIPWorld created: they are stable for the life of the graph, and two FLIP worlds with the same
// voxel size have different objects); the first node that sees a set of objects creates the device world and registers them.
// Marshalling goes through page-locked staging buffers owned by the world (flipb200_host_alloc, grown on demand) and the
// asynchronous download calls (flipb200_*_download_begin / flipb200_download_wait): all outputs of a node cross PCIe
// concurrently. RESIDENT mode is the default: a socket grid is uploaded only if its OpenVDB object is not the one this world
// last wrote or read -- decided by a fingerprint (tree address, leaf count, active-voxel count and a hash of the value
// buffers of up to 64 evenly spaced leaves), so a whole-grid edit by an un-accelerated node in between (VDBRenormalizeSDF,
// CombineVDB, ...) is seen; FLIPB200_RESIDENT=0 uploads every input of every node (for graphs that edit single voxels of a
// world grid in place between two accelerated nodes). Every node still writes its outputs back, so un-accelerated nodes,
// the viewport and IO always see current OpenVDB objects.
// Grids with non-background TILES (a flood-filled level set carries -background tiles inside) are refused with a clear
// error: the device grids are leaf-only, and sampling such a grid as "background outside the leaves" would be wrong.
// FLIPB200_PLUGIN_MARSHAL_ONLY: only the OpenVDB <-> flat-array marshalling below is compiled (no Zeno headers, no
// node registration). oracle/ref/ref_driver.cpp includes this file that way to round-trip REAL reference grids and
// particle trees through upload()/download() against a loopback C ABI (tests/test_plugin_cpu.py).
#ifndef FLIPB200_PLUGIN_MARSHAL_ONLY
#include <zeno/zeno.h>
#include <zeno/VDBGrid.h>
#include <zeno/types/NumericObject.h>
#include <zeno/types/PrimitiveObject.h>
#else
#include <stdexcept>
#include <string>
#include <memory>
#include <tbb/parallel_for.h>
#endif

#include <openvdb/openvdb.h>
#include <openvdb/points/PointDataGrid.h>
#include <openvdb/points/AttributeArray.h>
#ifndef FLIPB200_PLUGIN_MARSHAL_ONLY
#include <openvdb/tools/LevelSetTracker.h>
#endif

#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <initializer_list>
#include <mutex>
#include <vector>

#include "flipb200.h"

// OpenVDB's own test hook (openvdb/points/AttributeArray.h:353,756): raw codec words without a decode/encode trip
#ifndef FLIPB200_HAVE_TEST_ATTRIBUTE_ARRAY
class TestAttributeArray {
public:
    static char* bytes(openvdb::points::AttributeArray& a) { return a.dataAsByteArray(); }
    static const char* bytes(const openvdb::points::AttributeArray& a) { return a.constDataAsByteArray(); }
};
#endif

namespace zeno {
namespace flipb200 {

using Mask = openvdb::util::NodeMask<3>;
using PositionCodec = openvdb::points::FixedPointCodec</*one byte*/ false>;   // FF/FLIP_vdb.h:28
using position_attribute = openvdb::points::TypedAttributeArray<openvdb::Vec3f, PositionCodec>;
using velocity_attribute = openvdb::points::TypedAttributeArray<openvdb::Vec3f, openvdb::points::TruncateCodec>;

inline void check(int rc, const char* what) {
    if (rc == FLIPB200_OK) return;
    const std::string msg = std::string(what) + ": libflipb200 error " + std::to_string(rc) + ": " + flipb200_last_error();
#ifdef FLIPB200_PLUGIN_MARSHAL_ONLY
    throw std::runtime_error(msg);
#else
    throw makeError(msg);   // -> GraphException with the node name (zeno/src/core/Graph.cpp:88-90)
#endif
}

// a page-locked buffer that only grows
struct Pinned {
    void* p = nullptr;
    size_t cap = 0;
    Pinned() {}
    Pinned(const Pinned&) = delete;
    Pinned& operator=(const Pinned&) = delete;
    ~Pinned() { if (p) flipb200_host_free(p); }
    template <typename T> T* need(size_t count) {
        const size_t bytes = count * sizeof(T);
        if (bytes > cap) {
            if (p) flipb200_host_free(p);
            p = nullptr; cap = 0;
            const size_t c = bytes + bytes / 4 + 4096;
            check(flipb200_host_alloc(c, &p), "host_alloc");
            cap = c;
        }
        return static_cast<T*>(p);
    }
};
struct Fingerprint {
    const void* tree = nullptr;
    size_t leaves = 0;
    uint64_t active = 0, hash = 0;
    bool valid = false;
    bool operator==(const Fingerprint& o) const { return valid && o.valid && tree == o.tree && leaves == o.leaves && active == o.active && hash == o.hash; }
};
inline uint64_t fnv(uint64_t h, const void* data, size_t bytes) {
    const uint64_t* q = static_cast<const uint64_t*>(data);
    for (size_t i = 0; i < bytes / 8; i++) { h ^= q[i]; h *= 0x100000001b3ull; }
    return h;
}
struct WorldHolder {
    flipb200_world* w = nullptr;
    float dx = 0.f;
    struct Stage { Pinned o, m, v; int n = 0; float bg[3] = {0.f, 0.f, 0.f}; } stage[FLIPB200_NUM_GRIDS];
    Pinned po, pve, pP, pV;                       // particle staging
    Pinned primPos, primVel;                      // VDBPointsToPrimitive staging
    int pnl = 0; uint64_t pnp = 0;
    uint32_t reseeds = 0;                         // FluidReseed calls so far (seed sequence)
    Fingerprint fp[FLIPB200_NUM_GRIDS + 1];       // what the device copy of each slot corresponds to (last slot: particles)
    ~WorldHolder() { if (w) flipb200_world_destroy(w); }
};
inline bool resident() { static const bool r = !(std::getenv("FLIPB200_RESIDENT") && std::atoi(std::getenv("FLIPB200_RESIDENT")) == 0); return r; }

// the device world of the FLIP world these socket objects belong to
inline WorldHolder& world_for(float dx, std::initializer_list<const void*> objects) {
    static std::mutex mtx;
    static std::map<const void*, WorldHolder*> byObject;
    static std::vector<std::unique_ptr<WorldHolder>> worlds;
    std::lock_guard<std::mutex> lock(mtx);
    WorldHolder* h = nullptr;
    for (const void* o : objects) {
        auto it = o ? byObject.find(o) : byObject.end();
        if (it != byObject.end() && it->second->dx == dx) { h = it->second; break; }
    }
    if (!h) {
        worlds.push_back(std::make_unique<WorldHolder>());
        h = worlds.back().get();
        h->dx = dx;
        check(flipb200_world_create(/*device*/ 0, dx, &h->w), "SetFLIPWorld");
    }
    for (const void* o : objects) if (o) byObject[o] = h;
    return *h;
}

// ---------------------------------------------------------------- leaves <-> flat arrays
template <typename GridT> struct Channels { static constexpr int n = 1; };
template <> struct Channels<openvdb::Vec3fGrid> { static constexpr int n = 3; };

template <typename TreeT>
bool has_foreign_tiles(const TreeT& t) {
    using Root = typename TreeT::RootNodeType;
    const auto bg = t.background();
    for (auto it = t.root().cbeginValueAll(); it; ++it) if (!(*it == bg)) return true;
    for (auto c2 = t.root().cbeginChildOn(); c2; ++c2) {
        for (auto v = c2->cbeginValueAll(); v; ++v) if (!(*v == bg)) return true;
        for (auto c1 = c2->cbeginChildOn(); c1; ++c1)
            for (auto v = c1->cbeginValueAll(); v; ++v) if (!(*v == bg)) return true;
    }
    (void)sizeof(Root);
    return false;
}
template <typename GridT, typename LeafT>
Fingerprint fingerprint(const GridT& g, const std::vector<const LeafT*>& leaves) {
    Fingerprint f;
    f.tree = &g.tree(); f.leaves = leaves.size(); f.active = g.tree().activeVoxelCount();
    uint64_t h = 0xcbf29ce484222325ull;
    const size_t n = leaves.size(), step = n > 64 ? n / 64 : 1;
    for (size_t i = 0; i < n; i += step) h = fnv(h, leaves[i]->buffer().data(), sizeof(typename GridT::ValueType) * 512);
    f.hash = h; f.valid = true;
    return f;
}

template <typename GridT>
void upload(WorldHolder& h, int id, const typename GridT::Ptr& g) {
    if (!g) return;
    using Leaf = typename GridT::TreeType::LeafNodeType;
    constexpr int C = Channels<GridT>::n;
    std::vector<const Leaf*> leaves;
    g->tree().getNodes(leaves);
    const Fingerprint now = fingerprint<GridT, Leaf>(*g, leaves);
    if (resident() && h.fp[id] == now) return;   // the device copy is this object
    if (has_foreign_tiles(g->tree()))
        check(FLIPB200_ERR_ARG, "grid upload: the grid has tiles that differ from its background (e.g. a flood-filled level set); "
                                "voxelize its narrow band / interior first -- the device grids are leaf-only");
    const size_t n = leaves.size();
    WorldHolder::Stage& st = h.stage[id];
    int32_t* o = st.o.need<int32_t>(3 * n + 1);
    uint64_t* m = st.m.need<uint64_t>(8 * n + 1);
    float* v = st.v.need<float>(size_t(512) * C * n + 1);
    tbb::parallel_for(size_t(0), n, [&](size_t i) {
        const auto c = leaves[i]->origin();
        o[3 * i] = c.x(); o[3 * i + 1] = c.y(); o[3 * i + 2] = c.z();
        for (int k = 0; k < 8; k++) m[8 * i + k] = leaves[i]->getValueMask().template getWord<Mask::Word>(k);
        std::memcpy(&v[size_t(512) * C * i], leaves[i]->buffer().data(), sizeof(float) * 512 * C);  // Vec3f buffer = AOS
    });
    float bg[3] = {0.f, 0.f, 0.f};
    if constexpr (C == 3) { bg[0] = g->background()[0]; bg[1] = g->background()[1]; bg[2] = g->background()[2]; }
    else bg[0] = g->background();
    check(flipb200_grid_upload(h.w, id, int(n), o, m, v, C == 3 ? FLIPB200_AOS : FLIPB200_SOA, bg), "grid_upload");
    h.fp[id] = now;
}

// the tree a write-back replaces: when nobody else holds it, Tree::clear() frees its nodes in parallel (the destructor would do it
// one leaf at a time on this thread: milliseconds per grid at a few thousand leaves)
template <typename TreePtrT>
void release_tree(TreePtrT old) { if (old && old.use_count() <= 2) old->clear(); }
// write-back, part 1: the device stages the grid and the copies start (second stream); nothing is valid before download_wait
template <typename GridT>
void download_begin(WorldHolder& h, int id) {
    constexpr int C = Channels<GridT>::n;
    int n = 0;
    check(flipb200_grid_leaf_count(h.w, id, &n), "grid_leaf_count");
    WorldHolder::Stage& st = h.stage[id];
    int32_t* o = st.o.need<int32_t>(3 * size_t(n) + 1);
    uint64_t* m = st.m.need<uint64_t>(8 * size_t(n) + 1);
    float* v = st.v.need<float>(size_t(512) * C * n + 1);
    check(flipb200_grid_download_begin(h.w, id, n, o, m, v, C == 3 ? FLIPB200_AOS : FLIPB200_SOA, st.bg, &st.n), "grid_download_begin");
}
// part 2 (after flipb200_download_wait): the reference nodes replace trees (FF/FLIP_vdb.cpp:1302,3087,3488); so does the write-back
template <typename GridT>
void download_finish(WorldHolder& h, int id, typename GridT::Ptr& g) {
    using Tree = typename GridT::TreeType;
    using Leaf = typename Tree::LeafNodeType;
    constexpr int C = Channels<GridT>::n;
    WorldHolder::Stage& st = h.stage[id];
    const int n = st.n;
    const int32_t* o = static_cast<const int32_t*>(st.o.p);
    const uint64_t* m = static_cast<const uint64_t*>(st.m.p);
    const float* v = static_cast<const float*>(st.v.p);
    typename Tree::Ptr tree;
    if constexpr (C == 3) tree = std::make_shared<Tree>(openvdb::Vec3f(st.bg[0], st.bg[1], st.bg[2]));
    else tree = std::make_shared<Tree>(st.bg[0]);
    // same leaf set as the tree the grid holds now (FieldAddVector, CutCellWeight on a warm world, ...): overwrite its leaves in
    // place -- no allocation, no page faults, nothing to free. The values and masks end up exactly as in a rebuilt tree.
    if (g->tree().leafCount() == openvdb::Index32(n) && n > 0 && !has_foreign_tiles(g->tree())) {
        std::vector<Leaf*> have(n, nullptr);
        std::atomic<bool> all{true};
        auto& cur = g->tree();
        tbb::parallel_for(0, n, [&](int i) {
            have[i] = cur.probeLeaf(openvdb::Coord(o[3 * i], o[3 * i + 1], o[3 * i + 2]));
            if (!have[i]) all = false;
        });
        bool sameBg = true;
        if constexpr (C == 3) sameBg = cur.background() == openvdb::Vec3f(st.bg[0], st.bg[1], st.bg[2]);
        else sameBg = cur.background() == st.bg[0];
        if (all && sameBg) {
            tbb::parallel_for(0, n, [&](int i) {
                Mask mask;
                for (int k = 0; k < 8; k++) mask.template getWord<Mask::Word>(k) = m[8 * size_t(i) + k];
                std::memcpy(have[i]->buffer().data(), &v[size_t(512) * C * i], sizeof(float) * 512 * C);
                have[i]->setValueMask(mask);
            });
            std::vector<const Leaf*> cl;
            g->tree().getNodes(cl);   // tree order: what the next upload's fingerprint walks
            h.fp[id] = fingerprint<GridT, Leaf>(*g, cl);
            return;
        }
    }
    // leaves are allocated and filled in parallel; only the pointer insertion into the tree is serial
    std::vector<Leaf*> leaves(n);
    tbb::parallel_for(0, n, [&](int i) {
        leaves[i] = new Leaf(openvdb::Coord(o[3 * i], o[3 * i + 1], o[3 * i + 2]), tree->background(), /*active=*/false);
        Mask mask;
        for (int k = 0; k < 8; k++) mask.template getWord<Mask::Word>(k) = m[8 * size_t(i) + k];
        std::memcpy(leaves[i]->buffer().data(), &v[size_t(512) * C * i], sizeof(float) * 512 * C);
        leaves[i]->setValueMask(mask);
    });
    for (int i = 0; i < n; i++) tree->addLeaf(leaves[i]);
    release_tree(g->treePtr());
    g->setTree(tree);
    std::vector<const Leaf*> cl;
    g->tree().getNodes(cl);
    h.fp[id] = fingerprint<GridT, Leaf>(*g, cl);
}
inline void download_wait(WorldHolder& h) { check(flipb200_download_wait(h.w), "download_wait"); }
template <typename GridT>
void download(WorldHolder& h, int id, typename GridT::Ptr& g) {
    download_begin<GridT>(h, id);
    download_wait(h);
    download_finish<GridT>(h, id, g);
}

inline Fingerprint fingerprint_particles(const openvdb::points::PointDataGrid& g, const std::vector<const openvdb::points::PointDataTree::LeafNodeType*>& leaves) {
    Fingerprint f;
    f.tree = &g.tree(); f.leaves = leaves.size(); f.active = g.tree().activeVoxelCount();
    uint64_t h = 0xcbf29ce484222325ull;
    const size_t n = leaves.size(), step = n > 64 ? n / 64 : 1;
    for (size_t i = 0; i < n; i += step) {
        const auto cnt = leaves[i]->getLastValue();
        h ^= cnt; h *= 0x100000001b3ull;
        if (!cnt) continue;
        for (const char* a : {"P", "v"}) {
            const auto& arr = leaves[i]->constAttributeArray(a);
            h = fnv(h, TestAttributeArray::bytes(arr), arr.isUniform() ? 0 : size_t(cnt) * 6);
        }
    }
    f.hash = h; f.valid = true;
    return f;
}
inline void upload_particles(WorldHolder& h, const openvdb::points::PointDataGrid::Ptr& g) {
    using Leaf = openvdb::points::PointDataTree::LeafNodeType;
    std::vector<const Leaf*> leaves;
    g->tree().getNodes(leaves);
    const Fingerprint now = fingerprint_particles(*g, leaves);
    if (resident() && h.fp[FLIPB200_NUM_GRIDS] == now) return;
    const size_t n = leaves.size();
    std::vector<uint64_t> begin(n + 1, 0);
    for (size_t i = 0; i < n; i++) begin[i + 1] = begin[i] + leaves[i]->getLastValue();
    int32_t* o = h.po.need<int32_t>(3 * n + 1);
    uint32_t* ve = h.pve.need<uint32_t>(512 * n + 1);
    uint16_t* P = h.pP.need<uint16_t>(3 * begin[n] + 4);
    uint16_t* V = h.pV.need<uint16_t>(3 * begin[n] + 4);
    tbb::parallel_for(size_t(0), n, [&](size_t i) {
        const auto c = leaves[i]->origin();
        o[3 * i] = c.x(); o[3 * i + 1] = c.y(); o[3 * i + 2] = c.z();
        for (int k = 0; k < 512; k++) ve[512 * i + k] = uint32_t(leaves[i]->getValue(openvdb::Index(k)));
        const uint32_t cnt = ve[512 * i + 511];
        if (!cnt) return;
        const auto& pa = leaves[i]->constAttributeArray("P");
        const auto& va = leaves[i]->constAttributeArray("v");
        const uint16_t* ps = reinterpret_cast<const uint16_t*>(TestAttributeArray::bytes(pa));
        const uint16_t* vs = reinterpret_cast<const uint16_t*>(TestAttributeArray::bytes(va));
        if (!pa.isUniform()) std::memcpy(&P[3 * begin[i]], ps, size_t(cnt) * 6);
        else for (uint32_t j = 0; j < cnt; j++) for (int a = 0; a < 3; a++) P[3 * (begin[i] + j) + a] = ps[a];
        if (!va.isUniform()) std::memcpy(&V[3 * begin[i]], vs, size_t(cnt) * 6);
        else for (uint32_t j = 0; j < cnt; j++) for (int a = 0; a < 3; a++) V[3 * (begin[i] + j) + a] = vs[a];
    });
    check(flipb200_particles_upload(h.w, int(n), o, ve, begin[n], P, V), "particles_upload");
    h.fp[FLIPB200_NUM_GRIDS] = now;
}

inline void download_particles_begin(WorldHolder& h) {
    int nl = 0;
    uint64_t np = 0;
    check(flipb200_particles_info(h.w, &nl, &np), "particles_info");
    int32_t* o = h.po.need<int32_t>(3 * size_t(nl) + 1);
    uint32_t* ve = h.pve.need<uint32_t>(512 * size_t(nl) + 1);
    uint16_t* P = h.pP.need<uint16_t>(3 * np + 4);
    uint16_t* V = h.pV.need<uint16_t>(3 * np + 4);
    check(flipb200_particles_download_begin(h.w, nl, np, o, ve, P, V, &h.pnl, &h.pnp), "particles_download_begin");
}
inline void download_particles_finish(WorldHolder& h, openvdb::points::PointDataGrid::Ptr& g) {
    const int nl = h.pnl;
    const int32_t* o = static_cast<const int32_t*>(h.po.p);
    const uint32_t* ve = static_cast<const uint32_t*>(h.pve.p);
    const uint16_t* P = static_cast<const uint16_t*>(h.pP.p);
    const uint16_t* V = static_cast<const uint16_t*>(h.pV.p);
    // the descriptor the reference builds for a new particle tree (FF/FLIP_vdb.cpp:3398-3404)
    // (initializeAttributes accepts only the one-attribute position descriptor; "v" is appended per leaf, as the reference does)
    auto pdescr = openvdb::points::AttributeSet::Descriptor::create(position_attribute::attributeType());
    auto pvdescr = pdescr->duplicateAppend("v", velocity_attribute::attributeType());
    auto tree = std::make_shared<openvdb::points::PointDataTree>();
    using PLeaf = openvdb::points::PointDataTree::LeafNodeType;
    std::vector<PLeaf*> leaves(nl);
    std::vector<uint64_t> begin(size_t(nl) + 1, 0);
    for (int i = 0; i < nl; i++) begin[i + 1] = begin[i] + ve[512 * size_t(i) + 511];
    tbb::parallel_for(0, nl, [&](int i) {
        leaves[i] = new PLeaf(openvdb::Coord(o[3 * i], o[3 * i + 1], o[3 * i + 2]), 0, /*active=*/false);
        const uint32_t cnt = ve[512 * size_t(i) + 511];
        leaves[i]->initializeAttributes(pdescr, cnt);
        leaves[i]->appendAttribute(leaves[i]->attributeSet().descriptor(), pvdescr, 1);
        std::vector<openvdb::PointDataIndex32> offs(512);
        for (int k = 0; k < 512; k++) offs[k] = openvdb::PointDataIndex32(ve[512 * size_t(i) + k]);
        leaves[i]->setOffsets(offs, /*updateValueMask=*/true);
        auto& pa = leaves[i]->attributeArray("P");
        auto& va = leaves[i]->attributeArray("v");
        pa.expand(); va.expand();
        std::memcpy(TestAttributeArray::bytes(pa), &P[3 * begin[i]], size_t(cnt) * 6);
        std::memcpy(TestAttributeArray::bytes(va), &V[3 * begin[i]], size_t(cnt) * 6);
    });
    for (int i = 0; i < nl; i++) tree->addLeaf(leaves[i]);
    release_tree(g->treePtr());
    g->setTree(tree);
    std::vector<const openvdb::points::PointDataTree::LeafNodeType*> cl;
    g->tree().getNodes(cl);
    h.fp[FLIPB200_NUM_GRIDS] = fingerprint_particles(*g, cl);
}
inline void download_particles(WorldHolder& h, openvdb::points::PointDataGrid::Ptr& g) {
    download_particles_begin(h);
    download_wait(h);
    download_particles_finish(h, g);
}

#ifndef FLIPB200_PLUGIN_MARSHAL_ONLY
inline float dx_of(INode* node) {
    float dx = node->get_param<float>("dx");
    if (node->has_input("Dx")) dx = node->get_input("Dx")->as<NumericObject>()->get<float>();
    return dx;
}

#endif  // !FLIPB200_PLUGIN_MARSHAL_ONLY

}  // namespace flipb200

#ifndef FLIPB200_PLUGIN_MARSHAL_ONLY
using namespace flipb200;

// ---- FLIP_P2G (FF/nosys/P2G.cpp:11-62)
struct FLIP_P2G : zeno::INode {
    virtual void apply() override {
        const float dx = dx_of(this);
        const int n = get_param<int>("VelExtraLayer");
        auto Particles = get_input("Particles")->as<VDBPointsGrid>();
        auto VelGrid = get_input("Velocity")->as<VDBFloat3Grid>();
        auto PostP2GVelGrid = get_input("PostP2GVelocity")->as<VDBFloat3Grid>();
        auto LiquidSDFGrid = get_input("LiquidSDF")->as<VDBFloatGrid>();
        WorldHolder& h = world_for(dx, {Particles, VelGrid, PostP2GVelGrid, LiquidSDFGrid});
        upload_particles(h, Particles->m_grid);
        check(flipb200_p2g(h.w, dx, n), "FLIP_P2G");
        download_begin<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY);
        download_begin<openvdb::Vec3fGrid>(h, FLIPB200_POSTADV_VELOCITY);
        download_begin<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF);
        download_wait(h);
        download_finish<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, VelGrid->m_grid);
        download_finish<openvdb::Vec3fGrid>(h, FLIPB200_POSTADV_VELOCITY, PostP2GVelGrid->m_grid);
        download_finish<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, LiquidSDFGrid->m_grid);
    }
};
static int defFLIP_P2G = zeno::defNodeClass<FLIP_P2G>("FLIP_P2G",
    {/* inputs: */ {"Dx", "Particles", "Velocity", "PostP2GVelocity", "LiquidSDF"},
     /* outputs: */ {},
     /* params: */ {{"float", "dx", "0.01 0.0"}, {"int", "VelExtraLayer", "3"}},
     /* category: */ {"FLIPSolver"}});

// ---- G2PAdvectorSheetty (FF/nosys/SheetG2PAdvector.cpp:15-81)
struct G2PAdvectorSheet : zeno::INode {
    virtual void apply() override {
        const float dt = get_input("dt")->as<NumericObject>()->get<float>();
        const float dx = dx_of(this);
        const int surfaceSize = get_param<int>("surface_size");
        const int RK_ORDER = get_param<int>("RK_ORDER");
        const float pic_min = get_input("pic_min")->as<NumericObject>()->get<float>();
        const float pic_max = get_input("pic_max")->as<NumericObject>()->get<float>();
        auto particles = get_input("Particles")->as<VDBPointsGrid>();
        auto velocity = get_input("Velocity")->as<VDBFloat3Grid>();
        auto liquidsdf = get_input("LiquidSDF")->as<VDBFloatGrid>();
        auto velocity_viscous = get_input("ViscousVelocity")->as<VDBFloat3Grid>();
        auto velocity_after_p2g = get_input("PostAdvVelocity")->as<VDBFloat3Grid>();
        WorldHolder& h = world_for(dx, {particles, velocity, liquidsdf, velocity_after_p2g});
        if (has_input("SolidSDF")) upload<openvdb::FloatGrid>(h, FLIPB200_SOLID_SDF, get_input("SolidSDF")->as<VDBFloatGrid>()->m_grid);
        if (has_input("SolidVelocity")) upload<openvdb::Vec3fGrid>(h, FLIPB200_SOLID_VELOCITY, get_input("SolidVelocity")->as<VDBFloat3Grid>()->m_grid);
        upload_particles(h, particles->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_POSTADV_VELOCITY, velocity_after_p2g->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquidsdf->m_grid);
        const bool same = velocity_viscous->m_grid == velocity->m_grid;   // the packaged graphs wire the same object
        if (!same) upload<openvdb::Vec3fGrid>(h, FLIPB200_VISCOUS_VELOCITY, velocity_viscous->m_grid);
        check(flipb200_g2p_advect_sheetty(h.w, dt, dx, surfaceSize, RK_ORDER, pic_min, pic_max, same ? 1 : 0), "G2PAdvectorSheetty");
        download_particles(h, particles->m_grid);
    }
};
static int defG2PAdvectorSheet = zeno::defNodeClass<G2PAdvectorSheet>("G2PAdvectorSheetty",
    {/* inputs: */ {"dt", "Dx", {"float", "pic_min", "0.03"}, {"float", "pic_max", "0.05"}, "Particles", "Velocity", "ViscousVelocity",
                    "LiquidSDF", "PostAdvVelocity", "SolidSDF", "SolidVelocity"},
     /* outputs: */ {},
     /* params: */ {{"float", "dx", "0.01 0.0"}, {"int", "RK_ORDER", "1 1 4"}, {"float", "pic_smoothness", "0.1 0.0 1.0"}, {"int", "surface_size", "4 0 8"}},
     /* category: */ {"FLIPSolver"}});

// ---- VDBRenormalizeSDF (projects/zenvdb/VDBRenormalize.cpp:18-52)
struct VDBRenormalizeSDF : zeno::INode {
    virtual void apply() override {
        auto inoutSDF = get_input("inoutSDF")->as<VDBFloatGrid>();
        const int normIter = get_param<int>("iterations");
        const int dilateIter = get_param<int>("dilateIters");
        WorldHolder& h = world_for(float(inoutSDF->m_grid->voxelSize()[0]), {inoutSDF});
        if (dilateIter != 0) {
            // the tracker's dilate / erode changes the narrow band's TOPOLOGY (openvdb/tools/LevelSetTracker.h); it is not accelerated and
            // runs here exactly as in the reference node (projects/zenvdb/VDBRenormalize.cpp:24-31), the normalisation passes follow on the device
            auto lstracker = openvdb::tools::LevelSetTracker<openvdb::FloatGrid>(*(inoutSDF->m_grid));
            lstracker.setState({openvdb::math::FIRST_BIAS, openvdb::math::TVD_RK3, 1, 1});
            lstracker.setTrimming(openvdb::tools::lstrack::TrimMode::kNone);
            if (dilateIter > 0) lstracker.dilate(dilateIter);
            else lstracker.erode(dilateIter);
        }
        // the socket can carry any level set: it travels in the generic float slot, not in the world's LiquidSDF
        upload<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, inoutSDF->m_grid);
        check(flipb200_renormalize_sdf(h.w, FLIPB200_KILLER_SDF, normIter, 0), "VDBRenormalizeSDF");
        download<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, inoutSDF->m_grid);
        set_output("inoutSDF", get_input("inoutSDF"));
    }
};
static int defVDBRenormalizeSDF = zeno::defNodeClass<VDBRenormalizeSDF>("VDBRenormalizeSDF",
    {/* inputs: */ {"inoutSDF"}, /* outputs: */ {"inoutSDF"},
     /* params: */ {{"enum 1oUpwind", "method", "1oUpwind"}, {"int", "iterations", "4"}, {"int", "dilateIters", "0"}},
     /* category: */ {"openvdb"}});

// ---- VDBSmoothSDF (projects/zenvdb/VDBRenormalize.cpp:108-133; "deprecated" category, still wired in the packaged FLIP template)
struct VDBSmoothSDF : zeno::INode {
    virtual void apply() override {
        auto inoutSDF = get_input("inoutSDF")->as<VDBFloatGrid>();
        const int width = get_param<int>("width");
        const int iterations = get_param<int>("iterations");
        WorldHolder& h = world_for(float(inoutSDF->m_grid->voxelSize()[0]), {inoutSDF});
        upload<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, inoutSDF->m_grid);
        check(flipb200_smooth_sdf(h.w, FLIPB200_KILLER_SDF, width, iterations), "VDBSmoothSDF");
        download<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, inoutSDF->m_grid);
        set_output("inoutSDF", get_input("inoutSDF"));
    }
};
static int defVDBSmoothSDF = zeno::defNodeClass<VDBSmoothSDF>("VDBSmoothSDF",
    {/* inputs: */ {"inoutSDF"}, /* outputs: */ {"inoutSDF"},
     /* params: */ {{"int", "width", "1"}, {"int", "iterations", "1"}, {"string", "DEPRECATED", "Use VDBSmooth Instead"}},
     /* category: */ {"deprecated"}});

// ---- VDBErodeSDF (projects/zenvdb/VDBRenormalize.cpp:155-185)
struct VDBErodeSDF : zeno::INode {
    virtual void apply() override {
        auto inoutSDF = get_input("inoutSDF")->as<VDBFloatGrid>();
        const float depth = get_input("depth")->as<zeno::NumericObject>()->get<float>();
        WorldHolder& h = world_for(float(inoutSDF->m_grid->voxelSize()[0]), {inoutSDF});
        upload<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, inoutSDF->m_grid);
        check(flipb200_erode_sdf(h.w, FLIPB200_KILLER_SDF, depth), "VDBErodeSDF");
        download<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, inoutSDF->m_grid);
        set_output("inoutSDF", get_input("inoutSDF"));
    }
};
static int defVDBErodeSDF = zeno::defNodeClass<VDBErodeSDF>("VDBErodeSDF",
    {/* inputs: */ {"inoutSDF", {"float", "depth"}}, /* outputs: */ {"inoutSDF"}, /* params: */ {}, /* category: */ {"openvdb"}});

// ---- G2P_Advector (FF/nosys/G2P_Advector.cpp:16-69): the plain node
struct G2P_Advector : zeno::INode {
    virtual void apply() override {
        const float dt = get_input("dt")->as<NumericObject>()->get<float>();
        const float dx = dx_of(this);
        const float smoothness = get_param<float>("pic_smoothness");
        const int RK_ORDER = get_param<int>("RK_ORDER");
        auto particles = get_input("Particles")->as<VDBPointsGrid>();
        auto velocity = get_input("Velocity")->as<VDBFloat3Grid>();
        auto velocity_after_p2g = get_input("PostAdvVelocity")->as<VDBFloat3Grid>();
        if (has_input("SolidSDF"))
            throw makeError("G2P_Advector (libflipb200): with a SolidSDF connected the reference node dereferences a null liquid SDF "
                            "(FF/FLIP_vdb.cpp:3251-3278); use G2PAdvectorSheetty");
        WorldHolder& h = world_for(dx, {particles, velocity, velocity_after_p2g});
        upload_particles(h, particles->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_POSTADV_VELOCITY, velocity_after_p2g->m_grid);
        check(flipb200_g2p_advect(h.w, dt, dx, RK_ORDER, smoothness), "G2P_Advector");
        download_particles(h, particles->m_grid);
    }
};
static int defG2P_Advector = zeno::defNodeClass<G2P_Advector>("G2P_Advector",
    {/* inputs: */ {"dt", "Dx", "Particles", "Velocity", "PostAdvVelocity", "SolidSDF", "SolidVelocity"}, /* outputs: */ {},
     /* params: */ {{"float", "dx", "0.01 0.0"}, {"int", "RK_ORDER", "1 1 4"}, {"float", "pic_smoothness", "0.02 0.0 1.0"}},
     /* category: */ {"FLIPSolver"}});

// ---- KillParticlesInSDF (FF/nosys/KillParticles.cpp:150-165)
struct KillParticlesInSDF : zeno::INode {
    virtual void apply() override {
        auto points = get_input("Particles")->as<VDBPointsGrid>();
        auto sdf = get_input("KillerSDF")->as<VDBFloatGrid>();
        const std::string op = has_input("OpType:") ? get_param<std::string>("OpType") : std::string("KEEP");
        WorldHolder& h = world_for(float(points->m_grid->voxelSize()[0]), {points});
        upload_particles(h, points->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, sdf->m_grid);
        check(flipb200_kill_particles_in_sdf(h.w, FLIPB200_KILLER_SDF, op == "KEEP" ? 1 : 0), "KillParticlesInSDF");
        download_particles(h, points->m_grid);
        set_output("Particles", get_input("Particles"));
    }
};
static int defKillParticlesInSDF = zeno::defNodeClass<KillParticlesInSDF>("KillParticlesInSDF",
    {/* inputs: */ {"Particles", "KillerSDF"}, /* outputs: */ {"Particles"}, /* params: */ {{"enum KEEP DEL", "OpType", "KEEP"}},
     /* category: */ {"FLIPSolver"}});

// ---- FluidReseed (FF/nosys/FLIP_Reseed.cpp:8-31). The reference draws where its jitter table starts from std::random_device
// (FF/FLIP_vdb.cpp:2081-2084); the device path is seeded: FLIPB200_SEED (default 0) + the number of reseeds this world has done,
// so successive substeps do not repeat a pattern and a run is reproducible. seed_fixed() pins the seed for the parity tests.
inline uint32_t& seed_base() { static uint32_t s = std::getenv("FLIPB200_SEED") ? uint32_t(std::strtoul(std::getenv("FLIPB200_SEED"), nullptr, 10)) : 0u; return s; }
inline bool& seed_fixed() { static bool f = false; return f; }
struct FluidReseed : zeno::INode {
    virtual void apply() override {
        auto particles = get_input("Particles")->as<VDBPointsGrid>();
        auto liquidSDF = get_input("LiquidSDF")->as<VDBFloatGrid>();
        auto liquidVel = get_input("FluidVel")->as<VDBFloat3Grid>();
        WorldHolder& h = world_for(float(particles->m_grid->voxelSize()[0]), {particles, liquidSDF, liquidVel});
        upload_particles(h, particles->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquidSDF->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, liquidVel->m_grid);
        const uint32_t seed = seed_fixed() ? seed_base() : seed_base() + h.reseeds++;
        check(flipb200_fluid_reseed(h.w, seed), "FluidReseed");
        download_particles(h, particles->m_grid);
    }
};
static int defFluidReseed = zeno::defNodeClass<FluidReseed>("FluidReseed",
    {/* inputs: */ {"Particles", "LiquidSDF", "FluidVel"}, /* outputs: */ {}, /* params: */ {}, /* category: */ {"FLIPSolver"}});

// ---- ParticleEmitter (FF/nosys/ParticleEmitter.cpp:9-62). Accelerated: the constant-velocity branch (vx / vy / vz or VelocityInit)
// with a ShapeSDF on the world's own transform; a VelocityVolume or a shape grid with another transform is refused with a message
// (the device grids carry no transform of their own).
struct ParticleEmitter : zeno::INode {
    virtual void apply() override {
        auto particles = get_input("Particles")->as<VDBPointsGrid>();
        auto shape = get_input("ShapeSDF")->as<VDBFloatGrid>();
        float vx = get_param<float>("vx"), vy = get_param<float>("vy"), vz = get_param<float>("vz");
        if (has_input("VelocityVolume") && std::dynamic_pointer_cast<VDBFloat3Grid>(get_input("VelocityVolume")))
            check(FLIPB200_ERR_ARG, "ParticleEmitter: the VelocityVolume branch is not accelerated (constant emission velocity only)");
        if (has_input("VelocityInit")) {
            auto vel = get_input("VelocityInit")->as<zeno::NumericObject>()->get<zeno::vec3f>();
            vx = vel[0]; vy = vel[1]; vz = vel[2];
        }
        if (!(shape->m_grid->transform() == particles->m_grid->transform()))
            check(FLIPB200_ERR_ARG, "ParticleEmitter: the ShapeSDF must share the particle grid's transform (voxel size dx, cell centred)");
        WorldHolder& h = world_for(float(particles->m_grid->voxelSize()[0]), {particles});
        upload_particles(h, particles->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, shape->m_grid);
        const uint32_t seed = seed_fixed() ? seed_base() : seed_base() + 0x9e3779b9u + h.reseeds++;
        check(flipb200_emit_liquid(h.w, FLIPB200_KILLER_SDF, vx, vy, vz, seed), "ParticleEmitter");
        download_particles(h, particles->m_grid);
        set_output("Particles", get_input("Particles"));
    }
};
static int defParticleEmitter = zeno::defNodeClass<ParticleEmitter>("ParticleEmitter",
    {/* inputs: */ {"Particles", "ShapeSDF", "VelocityVolume", "VelocityInit", "LiquidSDF"}, /* outputs: */ {"Particles"},
     /* params: */ {{"float", "vx", "0.0"}, {"float", "vy", "0.0"}, {"float", "vz", "0.0"}}, /* category: */ {"FLIPSolver"}});

// ---- FLIPApplyBoundary (FF/nosys/Update_Solid_SDF.cpp:9-49). The reference reads the socket "Static_SDF" although its descriptor
// lists "StatSolid_SDF": either name is accepted. Without a DynaSolid_SDF the reference's update is a no-op; so is this.
struct FLIP_Solid_Modifier : zeno::INode {
    virtual void apply() override {
        auto particles = get_input("Particles")->as<VDBPointsGrid>();
        auto stat = get_input(has_input("Static_SDF") ? "Static_SDF" : "StatSolid_SDF")->as<VDBFloatGrid>();
        if (!has_input("DynaSolid_SDF")) return;
        auto dyna = get_input("DynaSolid_SDF")->as<VDBFloatGrid>();
        int vertexCentred = -1;
        if (dyna->m_grid->transform() == particles->m_grid->transform()) vertexCentred = 0;
        else if (dyna->m_grid->transform() == stat->m_grid->transform()) vertexCentred = 1;
        if (vertexCentred < 0)
            check(FLIPB200_ERR_ARG, "FLIPApplyBoundary: the DynaSolid_SDF must be on the particle grid's or on the static SDF's transform");
        WorldHolder& h = world_for(float(particles->m_grid->voxelSize()[0]), {particles, stat});
        upload_particles(h, particles->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_SOLID_SDF, stat->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_KILLER_SDF, dyna->m_grid);
        check(flipb200_apply_boundary(h.w, FLIPB200_KILLER_SDF, vertexCentred), "FLIPApplyBoundary");
        download<openvdb::FloatGrid>(h, FLIPB200_SOLID_SDF, stat->m_grid);
    }
};
static int defFLIP_Solid_Modifier = zeno::defNodeClass<FLIP_Solid_Modifier>("FLIPApplyBoundary",
    {/* inputs: */ {"Particles", "DynaSolid_SDF", "StatSolid_SDF"}, /* outputs: */ {}, /* params: */ {}, /* category: */ {"FLIPSolver"}});

// ---- VDBPointsToPrimitive (projects/zenvdb/GetVDBPoints.cpp:76-266): the particles as a point primitive for the viewport
// or an exporter. With the store resident on the device (the usual case after an accelerated substep) nothing is uploaded and no
// OpenVDB tree is walked: the device writes world positions and velocities straight into pinned staging. A grid without the "v"
// attribute (not a FLIP particle grid) is not handled here.
struct VDBPointsToPrimitive : zeno::INode {
    virtual void apply() override {
        auto grid = get_input("grid")->as<VDBPointsGrid>();
        WorldHolder& h = world_for(float(grid->m_grid->voxelSize()[0]), {grid});
        upload_particles(h, grid->m_grid);
        int nl = 0;
        uint64_t np = 0;
        check(flipb200_particles_info(h.w, &nl, &np), "particles_info");
        float* pos = h.primPos.need<float>(3 * np + 4);
        float* vel = h.primVel.need<float>(3 * np + 4);
        check(flipb200_particles_to_points(h.w, pos, vel), "VDBPointsToPrimitive");
        auto ret = std::make_shared<zeno::PrimitiveObject>();
        ret->resize(np);
        auto& retpos = ret->add_attr<zeno::vec3f>("pos");
        auto& retvel = ret->add_attr<zeno::vec3f>("vel");
        tbb::parallel_for(uint64_t(0), np, [&](uint64_t i) {
            retpos[i] = zeno::vec3f(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
            retvel[i] = zeno::vec3f(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        });
        set_output("prim", ret);
    }
};
static int defVDBPointsToPrimitive = zeno::defNodeClass<VDBPointsToPrimitive>("VDBPointsToPrimitive",
    {/* inputs: */ {"grid"}, /* outputs: */ {"prim"}, /* params: */ {}, /* category: */ {"openvdb"}});

// ---- ParticleAddDV (FF/nosys/ParticleAddGravity.cpp:9-41)
struct ParticleAddDV : zeno::INode {
    virtual void apply() override {
        auto particles = get_input("Particles")->as<VDBPointsGrid>();
        auto dv = get_input("dv")->as<zeno::NumericObject>()->get<zeno::vec3f>();
        WorldHolder& h = world_for(float(particles->m_grid->voxelSize()[0]), {particles});
        upload_particles(h, particles->m_grid);
        check(flipb200_particles_add_dv(h.w, dv[0], dv[1], dv[2]), "ParticleAddDV");   // the node's channel is always "vel"
        download_particles(h, particles->m_grid);
    }
};
static int defParticleAddDV = zeno::defNodeClass<ParticleAddDV>("ParticleAddDV",
    {/* inputs: */ {"Particles", "dv"}, /* outputs: */ {},
     /* params: */ {{"string", "channel", "vel"}, {"float", "vx", "0.0"}, {"float", "vy", "0.0"}, {"float", "vz", "0.0"}},
     /* category: */ {"FLIPSolver"}});

// ---- CutCellWeight (FF/nosys/EvalFaceWeight.cpp:17-41)
struct CutCellWeightEval : zeno::INode {
    virtual void apply() override {
        auto face_weight = get_input("FaceWeight")->as<VDBFloat3Grid>();
        auto liquid_sdf = get_input("LiquidSDF")->as<VDBFloatGrid>();
        auto solid_sdf = get_input("SolidSDF")->as<VDBFloatGrid>();
        WorldHolder& h = world_for(float(liquid_sdf->m_grid->voxelSize()[0]), {liquid_sdf, face_weight});
        upload<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquid_sdf->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_SOLID_SDF, solid_sdf->m_grid);
        check(flipb200_face_weights(h.w), "CutCellWeight");
        download<openvdb::Vec3fGrid>(h, FLIPB200_FACE_WEIGHT, face_weight->m_grid);
    }
};
static int defCutCellWeightEval = zeno::defNodeClass<CutCellWeightEval>("CutCellWeight",
    {/* inputs: */ {"LiquidSDF", "SolidSDF", "FaceWeight"}, /* outputs: */ {}, /* params: */ {}, /* category: */ {"FLIPSolver"}});

// ---- PushOutLiquidSDF (FF/nosys/FixLiquidSDF.cpp:16-45)
struct PushOutLiquidSDF : zeno::INode {
    virtual void apply() override {
        const float dx = dx_of(this);
        auto liquid_sdf = get_input("LiquidSDF")->as<VDBFloatGrid>();
        auto solid_sdf = get_input("SolidSDF")->as<VDBFloatGrid>();
        WorldHolder& h = world_for(dx, {liquid_sdf});
        upload<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquid_sdf->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_SOLID_SDF, solid_sdf->m_grid);
        check(flipb200_pushout_sdf(h.w, dx), "PushOutLiquidSDF");
        download<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquid_sdf->m_grid);
    }
};
static int defPushOutLiquidSDF = zeno::defNodeClass<PushOutLiquidSDF>("PushOutLiquidSDF",
    {/* inputs: */ {"Dx", "LiquidSDF", "SolidSDF"}, /* outputs: */ {}, /* params: */ {{"float", "dx", "0.0"}}, /* category: */ {"FLIPSolver"}});

// ---- FieldAddVector (FF/nosys/FieldAddVector.cpp:16-50)
struct FieldAddVector : zeno::INode {
    virtual void apply() override {
        auto ivec3 = get_input("invec3")->as<NumericObject>()->get<zeno::vec3f>();
        auto velocity = get_input("Velocity")->as<VDBFloat3Grid>();
        WorldHolder& h = world_for(float(velocity->m_grid->voxelSize()[0]), {velocity});
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
        check(flipb200_add_vector(h.w, ivec3[0], ivec3[1], ivec3[2]), "FieldAddVector");
        download<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
    }
};
static int defFieldAddVector = zeno::defNodeClass<FieldAddVector>("FieldAddVector",
    {/* inputs: */ {"invec3", "Velocity", "FieldWeight"}, /* outputs: */ {}, /* params: */ {}, /* category: */ {"FLIPSolver"}});

// ---- CFL_dt (FF/nosys/CFL.cpp:13-27,53-70)
struct CFL : zeno::INode {
    virtual void apply() override {
        auto velocity = get_input("Velocity")->as<VDBFloat3Grid>();
        const float vdx = float(velocity->m_grid->voxelSize()[0]);
        float dx = get_param<float>("dx");
        if (has_input("Dx")) dx = get_input("Dx")->as<NumericObject>()->get<float>();
        WorldHolder& h = world_for(vdx, {velocity});
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
        float dt = 0.f;
        check(flipb200_cfl(h.w, &dt), "CFL_dt");
        printf("CFL dt: %f\n", dt);
        auto out_dt = zeno::IObject::make<zeno::NumericObject>();
        out_dt->set<float>(dx / vdx * dt);
        set_output("cfl_dt", out_dt);
    }
};
static int defCFL = zeno::defNodeClass<CFL>("CFL_dt",
    {/* inputs: */ {"Velocity", "Dx"}, /* outputs: */ {"cfl_dt"}, /* params: */ {{"float", "dx", "0.0"}}, /* category: */ {"FLIPSolver"}});

// ---- AssembleSolvePPE (FF/nosys/SolvePoissonPressureEqn.cpp:23-89)
struct AssembleSolvePPE : zeno::INode {
    virtual void apply() override {
        const float dt = get_input("dt")->as<NumericObject>()->get<float>();
        const float dx = dx_of(this);
        auto liquid_sdf = get_input("LiquidSDF")->as<VDBFloatGrid>();
        auto rhsgrid = get_input("Divergence")->as<VDBFloatGrid>();
        auto curr_pressure = get_input("Pressure")->as<VDBFloatGrid>();
        auto face_weight = get_input("CellFWeight")->as<VDBFloat3Grid>();
        auto velocity = get_input("Velocity")->as<VDBFloat3Grid>();
        auto solid_velocity = get_input("SolidVelocity")->as<VDBFloat3Grid>();
        const float tension_coef = get_input("SurfaceTension")->as<NumericObject>()->get<float>();
        const float density = get_input("Density")->as<NumericObject>()->get<float>();
        WorldHolder& h = world_for(dx, {liquid_sdf, velocity, curr_pressure, face_weight, rhsgrid});
        if (tension_coef > 0 && has_input("Curvature")) upload<openvdb::FloatGrid>(h, FLIPB200_CURVATURE, get_input("Curvature")->as<VDBFloatGrid>()->m_grid);
        check(flipb200_set_surface_tension(h.w, density, tension_coef > 0 ? tension_coef : 0.f), "AssembleSolvePPE");
        upload<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquid_sdf->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_FACE_WEIGHT, face_weight->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_SOLID_VELOCITY, solid_velocity->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_PRESSURE, curr_pressure->m_grid);   // warm start of the fallback path
        int iters = 0, status = 0;
        float res = 0.f;
        check(flipb200_solve_ppe(h.w, dt, dx, &iters, &res, &status), "AssembleSolvePPE");
        printf("iter:%d err:%e%s\n", iters + 1, res, status ? " (pure multigrid fallback)" : "");
        download_begin<openvdb::FloatGrid>(h, FLIPB200_PRESSURE);
        download_begin<openvdb::FloatGrid>(h, FLIPB200_DIVERGENCE);
        download_wait(h);
        download_finish<openvdb::FloatGrid>(h, FLIPB200_PRESSURE, curr_pressure->m_grid);
        download_finish<openvdb::FloatGrid>(h, FLIPB200_DIVERGENCE, rhsgrid->m_grid);
    }
};
static int defAssembleSolvePPE = zeno::defNodeClass<AssembleSolvePPE>("AssembleSolvePPE",
    {/* inputs: */ {"dt", "Dx", {"float", "Density", "1000.0"}, {"float", "SurfaceTension", "0.0"}, "LiquidSDF", "Divergence", "Pressure",
                    "CellFWeight", "Velocity", "SolidVelocity", "Curvature"},
     /* outputs: */ {}, /* params: */ {{"float", "dx", "0.0"}}, /* category: */ {"FLIPSolver"}});

// ---- SubtractPressureGradient (FF/nosys/SubtractPressureGradient.cpp:25-94)
struct SubtractPressureGradient : zeno::INode {
    virtual void apply() override {
        const float dx = dx_of(this);
        const int n = get_param<int>("VelExtraLayer");
        const float dt = get_input("dt")->as<NumericObject>()->get<float>();
        auto liquid_sdf = get_input("LiquidSDF")->as<VDBFloatGrid>();
        auto solid_sdf = get_input("SolidSDF")->as<VDBFloatGrid>();
        auto curr_pressure = get_input("Pressure")->as<VDBFloatGrid>();
        auto face_weight = get_input("CellFWeight")->as<VDBFloat3Grid>();
        auto velocity = get_input("Velocity")->as<VDBFloat3Grid>();
        auto solid_velocity = get_input("SolidVelocity")->as<VDBFloat3Grid>();
        const float tension_coef = get_input("SurfaceTension")->as<NumericObject>()->get<float>();
        const float density = get_input("Density")->as<NumericObject>()->get<float>();
        WorldHolder& h = world_for(dx, {liquid_sdf, velocity, curr_pressure, face_weight});
        if (tension_coef > 0 && has_input("Curvature")) upload<openvdb::FloatGrid>(h, FLIPB200_CURVATURE, get_input("Curvature")->as<VDBFloatGrid>()->m_grid);
        check(flipb200_set_surface_tension(h.w, density, tension_coef > 0 ? tension_coef : 0.f), "SubtractPressureGradient");
        upload<openvdb::FloatGrid>(h, FLIPB200_LIQUID_SDF, liquid_sdf->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_SOLID_SDF, solid_sdf->m_grid);
        upload<openvdb::FloatGrid>(h, FLIPB200_PRESSURE, curr_pressure->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_FACE_WEIGHT, face_weight->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
        upload<openvdb::Vec3fGrid>(h, FLIPB200_SOLID_VELOCITY, solid_velocity->m_grid);
        check(flipb200_subtract_grad(h.w, dt, dx, n), "SubtractPressureGradient");
        download<openvdb::Vec3fGrid>(h, FLIPB200_VELOCITY, velocity->m_grid);
    }
};
static int defSubtractPressureGradient = zeno::defNodeClass<SubtractPressureGradient>("SubtractPressureGradient",
    {/* inputs: */ {"dt", "Dx", {"float", "Density", "1000.0"}, {"float", "SurfaceTension", "0.0"}, "LiquidSDF", "SolidSDF", "Pressure",
                    "CellFWeight", "Velocity", "SolidVelocity", "Curvature"},
     /* outputs: */ {}, /* params: */ {{"float", "dx", "0.0"}, {"int", "VelExtraLayer", "3"}}, /* category: */ {"FLIPSolver"}});

#endif  // !FLIPB200_PLUGIN_MARSHAL_ONLY
}  // namespace zeno
