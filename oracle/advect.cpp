// TEST INFRASTRUCTURE ONLY -- see flip_oracle.h.
// G2P + advection + re-binning (K2, K7, K8).
#include <map>
#include "flip_oracle.h"
#include <cmath>
#include <numeric>

namespace orc {
namespace {

// local fp32 sampler (FF/FLIP_vdb.cpp:25-110)
inline float mixf(float a, float b, float w) { return a + (b - a) * w; }
float samplec(const Vec3Grid& g, int c, float x, float y, float z) {
    int bx = int(std::floor(double(x))), by = int(std::floor(double(y))), bz = int(std::floor(double(z)));
    float d[8];
    // index = i*4 + j*2 + k (:27-54)
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) d[i * 4 + j * 2 + k] = g.get(c, bx + i, by + j, bz + k);
    float wx = x - float(bx), wy = y - float(by), wz = z - float(bz);
    return mixf(mixf(mixf(d[0], d[1], wz), mixf(d[2], d[3], wz), wy),
                mixf(mixf(d[4], d[5], wz), mixf(d[6], d[7], wz), wy), wx);
}
void staggered_sample_f32(const Vec3Grid& g, const float p[3], float out[3]) {
    out[0] = samplec(g, 0, p[0] + 0.5f, p[1], p[2]);
    out[1] = samplec(g, 1, p[0], p[1] + 0.5f, p[2]);
    out[2] = samplec(g, 2, p[0], p[1], p[2] + 0.5f);
}

// openvdb::tools::BoxSampler::sample with double weights
// (openvdb/tools/Interpolation.h:712-737,763-778): a + float((b-a)*w)
template <int NC>
float box_sample_f64(const Grid<NC>& g, int c, double x, double y, double z) {
    int bx = int(std::floor(x)), by = int(std::floor(y)), bz = int(std::floor(z));
    double u = x - bx, v = y - by, w = z - bz;
    float d[2][2][2];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) d[i][j][k] = g.get(c, bx + i, by + j, bz + k);
    auto ip = [](float a, float b, double wt) { return a + float(double(b - a) * wt); };
    return ip(ip(ip(d[0][0][0], d[0][0][1], w), ip(d[0][1][0], d[0][1][1], w), v),
              ip(ip(d[1][0][0], d[1][0][1], w), ip(d[1][1][0], d[1][1][1], w), v), u);
}
// openvdb::tools::StaggeredBoxSampler::sample (Interpolation.h:944-953); input Vec3f -> Vec3R
void staggered_sample_f64(const Vec3Grid& g, const float p[3], float out[3]) {
    out[0] = box_sample_f64(g, 0, double(p[0]) + 0.5, double(p[1]), double(p[2]));
    out[1] = box_sample_f64(g, 1, double(p[0]), double(p[1]) + 0.5, double(p[2]));
    out[2] = box_sample_f64(g, 2, double(p[0]), double(p[1]), double(p[2]) + 0.5);
}

// custom_integrator (FF/FLIP_vdb.cpp:169-214)
void integrate(int order, const Vec3Grid& vel, float dtinvx, float ipos[3], const float V0[3]) {
    float q[3], V1[3], V2[3], V3[3];
    switch (order) {
    case 2:
        for (int a = 0; a < 3; a++) q[a] = ipos[a] + 0.5f * V0[a] * dtinvx;
        staggered_sample_f64(vel, q, V1);
        for (int a = 0; a < 3; a++) ipos[a] += V1[a] * dtinvx;
        break;
    case 3:
        for (int a = 0; a < 3; a++) q[a] = ipos[a] + 0.5f * V0[a] * dtinvx;
        staggered_sample_f64(vel, q, V1);
        for (int a = 0; a < 3; a++) q[a] = ipos[a] + dtinvx * (2.0f * V1[a] - V0[a]);
        staggered_sample_f64(vel, q, V2);
        for (int a = 0; a < 3; a++) ipos[a] += dtinvx * (V0[a] + 4.0f * V1[a] + V2[a]) * (1.0f / 6.0f);
        break;
    case 4:
        for (int a = 0; a < 3; a++) q[a] = ipos[a] + 0.5f * V0[a] * dtinvx;
        staggered_sample_f64(vel, q, V1);
        for (int a = 0; a < 3; a++) q[a] = ipos[a] + 0.5f * V1[a] * dtinvx;
        staggered_sample_f64(vel, q, V2);
        for (int a = 0; a < 3; a++) q[a] = ipos[a] + V2[a] * dtinvx;
        staggered_sample_f64(vel, q, V3);
        for (int a = 0; a < 3; a++)
            ipos[a] += dtinvx * (V0[a] + 2.0f * (V1[a] + V2[a]) + V3[a]) * (1.0f / 6.0f);
        break;
    case 1:
    default:
        for (int a = 0; a < 3; a++) ipos[a] += V0[a] * dtinvx;
    }
}

// K8: voxel-centre solid normal and normal velocity (FF/FLIP_vdb.cpp:3270-3367)
void build_solid_normals(const World& w, float dx, Vec3Grid& normal, FloatGrid& vn) {
    normal = Vec3Grid(0.f);
    normal.topologyCopyFrom(w.liquidSDF);
    normal.dilate(5, true);
    vn = FloatGrid(0.f);
    vn.topologyCopyFrom(normal);
    // AT(:,i) = {a-.5, b-.5, c-.5, 1} * dx ; invATA = {.5,.5,.5,.125}/dx^2 (:3292-3308)
    float AT[4][8];
    for (int i = 0; i < 8; i++) {
        int a = i / 4, b = (i - a * 4) / 2, c = i - a * 4 - b * 2;
        AT[0][i] = (a - 0.5f) * dx;
        AT[1][i] = (b - 0.5f) * dx;
        AT[2][i] = (c - 0.5f) * dx;
        AT[3][i] = 1.0f * dx;
    }
    float s = 1.0f / (dx * dx);
    float invATA[4] = {0.5f * s, 0.5f * s, 0.5f * s, 0.125f * s};
    for (int l = 0; l < normal.leafCount(); l++) {
        Coord o = normal.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(normal.masks[l], off)) continue;
            int i = o.x + (off >> 6), j = o.y + ((off >> 3) & 7), k = o.z + (off & 7);
            float data[8];
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    for (int c = 0; c < 2; c++) data[a * 4 + b * 2 + c] = w.solidSDF.get(0, i + a, j + b, k + c);
            float abcd[4];
            for (int r = 0; r < 4; r++) {
                float acc = 0.f;  // Eigen's dense product order is not pinned (third-party, absent); sequential here
                for (int q = 0; q < 8; q++) acc += AT[r][q] * data[q];
                abcd[r] = invATA[r] * acc;
            }
            if (abcd[3] < 0.5f) {
                float n[3] = {abcd[0], abcd[1], abcd[2]};
                // Vec3::normalize (openvdb/math/Vec3.h:204-210,366-374)
                float d = float(std::sqrt(double(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])));
                if (!(std::fabs(d - 0.f) <= 1.0e-7f)) {
                    float inv = 1.0f / d;
                    n[0] *= inv; n[1] *= inv; n[2] *= inv;
                }
                for (int c = 0; c < 3; c++) normal.leafVals(l, c)[off] = n[c];
                float sv[3] = {w.solidVelocity.get(0, i, j, k), w.solidVelocity.get(1, i, j, k),
                               w.solidVelocity.get(2, i, j, k)};
                // Vec3::dot: x*x' + y*y' + z*z'
                vn.leafVals(l)[off] = sv[0] * n[0] + sv[1] * n[1] + sv[2] * n[2];
                maskSet(vn.masks[l], off, true);
            } else {
                maskSet(normal.masks[l], off, false);
                maskSet(vn.masks[l], off, false);
            }
        }
    }
}

}  // namespace

// FLIP_vdb::custom_move_points_and_set_flip_vel (FF/FLIP_vdb.cpp:3237-3490) with
// point_to_counter_reducer2::operator() (:513-754) and set_new_attribute_list (:3406-3476).
static void custom_move_points_and_set_flip_vel(World& w, const FloatGrid* liquidSdfIn,
                                                const Vec3Grid& velocity, const Vec3Grid& velToAdvect,
                                                bool advSameField, const Vec3Grid& oldVelocity,
                                                bool hasSolid, float picMin, float picMax, float dt,
                                                float surfacedist, int rkOrder) {
    Points& pts = w.particles;
    const float dx = w.dx;
    FloatGrid solidBg(3 * dx);
    Vec3Grid solidVelBg(0.f);
    const FloatGrid& solidSdf = hasSolid ? w.solidSDF : solidBg;
    Vec3Grid normal(0.f);
    FloatGrid vn(0.f);
    if (hasSolid) build_solid_normals(w, dx, normal, vn);
    FloatGrid liquidBg(dx);
    const FloatGrid& liquidSdf = liquidSdfIn ? *liquidSdfIn : liquidBg;

    const size_t n = pts.size();
    std::vector<uint8_t> alive(n, 0);
    std::vector<uint64_t> tkey(n);
    std::vector<uint16_t> toff(n);
    std::vector<std::array<int, 3>> tijk(n);
    if (w.capturePreCodec) {
        w.preCodecPos.assign(3 * n, 0.f);
        w.preCodecVel.assign(3 * n, 0.f);
        w.preCodecAlive.assign(n, 0);
    }
    const float deep_threshold = float(-4.0 * dx);
    const float invdx = 1.0f / dx;
    const float dtinvx = dt / dx;

#pragma omp parallel for schedule(dynamic, 4)
    for (int l = 0; l < pts.leafCount(); l++) {
        Coord o = pts.origins[l];
        for (int off = 0; off < 512; off++) {
            uint32_t end = pts.voxelEnd[l][off];
            uint32_t beg = off == 0 ? 0u : pts.voxelEnd[l][off - 1];
            for (uint32_t it = beg; it < end; it++) {
                size_t gi = pts.leafBegin[l] + it;
                float pIs[3] = {float(o.x + (off >> 6)) + fxpt16_decode(pts.P[3 * gi + 0]),
                                float(o.y + ((off >> 3) & 7)) + fxpt16_decode(pts.P[3 * gi + 1]),
                                float(o.z + (off & 7)) + fxpt16_decode(pts.P[3 * gi + 2])};
                float pvel[3] = {half_decode(pts.v[3 * gi + 0]), half_decode(pts.v[3 * gi + 1]),
                                 half_decode(pts.v[3 * gi + 2])};
                float adv[3], old[3], carried[3];
                staggered_sample_f32(velocity, pIs, adv);
                staggered_sample_f32(oldVelocity, pIs, old);
                // FLIP/PIC blend factor (:628-648)
                float flip = 1.0f - picMin;
                float pls = box_sample_f64(liquidSdf, 0, double(pIs[0]), double(pIs[1]), double(pIs[2]));
                float t_coef = 1;
                if (pls < 0 && pls >= -surfacedist) {
                    t_coef = pls / -surfacedist;
                    t_coef = std::min(std::max(t_coef, 0.0f), 1.0f);
                }
                if (pls >= 0) t_coef = 0;
                if (surfacedist > 0)
                    flip = t_coef * flip + (1.0f - t_coef) * std::min(1.0f - picMax, flip);
                // pIspos + Vec3f{0.5f} is a float add, then promoted to Vec3R (:642-643)
                float pss = box_sample_f64(solidSdf, 0, double(pIs[0] + 0.5f), double(pIs[1] + 0.5f),
                                           double(pIs[2] + 0.5f));
                if (pss >= 0 && pss <= 2.0 * dx) {  // double comparison in the reference (2.0*m_dx)
                    float scoef = pss / (2.0f * dx);
                    flip = scoef * flip + (1.0f - scoef) * 1.0f;
                }
                if (advSameField) { carried[0] = adv[0]; carried[1] = adv[1]; carried[2] = adv[2]; }
                else staggered_sample_f32(velToAdvect, pIs, carried);
                for (int a = 0; a < 3; a++) pvel[a] = carried[a] + flip * (-old[a] + pvel[a]);  // (:668)

                float pIt[3] = {pIs[0], pIs[1], pIs[2]};
                if (pls >= -surfacedist) integrate(1, velocity, dtinvx, pIt, adv);
                else integrate(rkOrder, velocity, dtinvx, pIt, adv);
                int pt[3];
                for (int a = 0; a < 3; a++) pt[a] = int(std::floor(double(pIt[a] + 0.5f)));
                // pItpos + Vec3f{0.5}: Vec3f(double 0.5) -> float add (:677-678)
                float nps = box_sample_f64(solidSdf, 0, double(pIt[0] + 0.5f), double(pIt[1] + 0.5f),
                                           double(pIt[2] + 0.5f));
                if (nps < 0) {
                    if (nps < deep_threshold) continue;  // dropped (:682-685)
                    float sn[3] = {normal.get(0, pt[0], pt[1], pt[2]), normal.get(1, pt[0], pt[1], pt[2]),
                                   normal.get(2, pt[0], pt[1], pt[2])};
                    // pItpos -= new_pos_solid_sdf * snormal * invdx * 1.0f  (:690)
                    for (int a = 0; a < 3; a++) pIt[a] -= ((nps * sn[a]) * invdx) * 1.0f;
                    for (int a = 0; a < 3; a++) pt[a] = int(std::floor(double(pIt[a] + 0.5f)));
                    float vnv = vn.get(0, pt[0], pt[1], pt[2]);
                    float dot = sn[0] * pvel[0] + sn[1] * pvel[1] + sn[2] * pvel[2];
                    float coef = vnv - dot;
                    for (int a = 0; a < 3; a++) pvel[a] += coef * sn[a];  // (:693-694)
                }
                // codec write-back in place (:702-703)
                for (int a = 0; a < 3; a++) {
                    float local = float(double(pIt[a]) - double(pt[a]));
                    pts.P[3 * gi + a] = fxpt16_encode(local);
                    pts.v[3 * gi + a] = half_encode(pvel[a]);
                }
                if (w.capturePreCodec) {
                    for (int a = 0; a < 3; a++) { w.preCodecPos[3 * gi + a] = pIt[a]; w.preCodecVel[3 * gi + a] = pvel[a]; }
                    w.preCodecAlive[gi] = 1;
                }
                alive[gi] = 1;
                tijk[gi] = {pt[0], pt[1], pt[2]};
                tkey[gi] = leafKeyOf(pt[0], pt[1], pt[2]);
                toff[gi] = uint16_t(voxelOffset(pt[0], pt[1], pt[2]));
            }
        }
    }

    // K2: per-target-voxel counters with the >27 cap (:706-751), applied in source order.
    // (The reference applies the cap per TBB sub-range before the join, so which particle
    // is dropped is not deterministic there)
    std::unordered_map<uint64_t, uint32_t> counter;
    counter.reserve(n / 4 + 16);
    uint64_t dropped = 0;
    std::vector<uint32_t> order;
    order.reserve(n);
    for (size_t gi = 0; gi < n; gi++) {
        if (!alive[gi]) { dropped++; continue; }
        uint64_t vk = (tkey[gi] << 9) | toff[gi];
        uint32_t& c = counter[vk];
        if (c > 27) { dropped++; continue; }
        c++;
        order.push_back(uint32_t(gi));
    }
    w.droppedParticles = dropped;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (tkey[a] != tkey[b]) return tkey[a] < tkey[b];
        return toff[a] < toff[b];
    });
    Points np;
    const size_t m = order.size();
    np.P.resize(3 * m);
    np.v.resize(3 * m);
    for (size_t s = 0; s < m; s++) {
        uint32_t gi = order[s];
        int l;
        auto it = np.dir.find(tkey[gi]);
        if (it == np.dir.end()) {
            l = np.leafCount();
            np.dir.emplace(tkey[gi], l);
            np.origins.push_back(Coord(tijk[gi][0] & ~7, tijk[gi][1] & ~7, tijk[gi][2] & ~7));
            std::array<uint32_t, 512> z; z.fill(0);
            np.voxelEnd.push_back(z);
            np.leafBegin.push_back(s);
        } else l = it->second;
        np.voxelEnd[l][toff[gi]]++;
        for (int a = 0; a < 3; a++) {
            np.P[3 * s + a] = pts.P[3 * gi + a];
            np.v[3 * s + a] = pts.v[3 * gi + a];
        }
    }
    np.leafBegin.push_back(m);
    for (auto& ve : np.voxelEnd)
        for (int o = 1; o < 512; o++) ve[o] += ve[o - 1];
    pts = std::move(np);
}

// G2PAdvectorSheet::apply (FF/nosys/SheetG2PAdvector.cpp:15-54) -> FLIP_vdb::AdvectSheetty (:3221-3235)
void node_G2PAdvectorSheetty(World& w, float dt, float dx, int surfaceSize, int rkOrder, float picMin,
                             float picMax) {
    picMin = picMin > picMax ? picMax : picMin;
    // velocity and ViscousVelocity are distinct grid objects in the graph, so
    // adv_same_field (tree pointer equality, :617-618) is false.
    custom_move_points_and_set_flip_vel(w, &w.liquidSDF, w.velocity, w.viscousVelocity, false,
                                        w.postAdvVelocity, w.hasSolidSDF, picMin, picMax, dt,
                                        float(surfaceSize) * dx, rkOrder);
}

// G2P_Advector::apply (FF/nosys/G2P_Advector.cpp:16-47) -> FLIP_vdb::Advect (:3209-3219): liquid sdf = nullptr, the advected and
// the carried field are both `velocity`, pic_min = pic_smoothness, pic_max = 0.05, surfacedist = 0; no solid (with one connected the
// reference dereferences the null liquid sdf, :3251-3278)
void node_G2P_Advector(World& w, float dt, float dx, int rkOrder, float picSmoothness) {
    (void)dx;
    custom_move_points_and_set_flip_vel(w, nullptr, w.velocity, w.velocity, true, w.postAdvVelocity, false, picSmoothness, 0.05f, dt, 0.f, rkOrder);
}

// kill_particles_inside (FF/nosys/KillParticles.cpp:13-149): per leaf, per voxel, per particle in store order: the killer SDF is
// sampled with openvdb's BoxSampler at voxel + decoded position -- a float sum (Coord + Vec3f, math/Coord.h) handed to the
// sampler as doubles, in the SDF grid's OWN index space (no transform is applied) -- and the particle survives when the sample
// is <= 0 (keep) / >= 0 (delete). Survivors are written back through the attribute write handles (:138-141), i.e. the
// position goes decode -> encode once more (not the identity for 5461 of the 65536 fixed-point codes); the half velocity
// round-trips exactly. Leaves stay in the tree even when they end up empty (clearAttributes, :117-118).
void node_KillParticlesInSDF(World& w, const FloatGrid& sdf, bool keep) {
    const Points& in = w.particles;
    Points out;
    out.dir = in.dir;
    out.origins = in.origins;
    out.voxelEnd.resize(in.leafCount());
    out.P.reserve(in.P.size());
    out.v.reserve(in.v.size());
    for (int l = 0; l < in.leafCount(); l++) {
        out.leafBegin.push_back(out.P.size() / 3);
        const Coord o = in.origins[l];
        uint32_t count = 0;
        for (int off = 0; off < 512; off++) {
            const uint32_t b = off ? in.voxelEnd[l][off - 1] : 0u, e = in.voxelEnd[l][off];
            const int vx = o.x + (off >> 6), vy = o.y + ((off >> 3) & 7), vz = o.z + (off & 7);
            for (uint32_t i = b; i < e; i++) {
                const size_t gi = in.leafBegin[l] + i;
                const float px = fxpt16_decode(in.P[3 * gi]), py = fxpt16_decode(in.P[3 * gi + 1]), pz = fxpt16_decode(in.P[3 * gi + 2]);
                const float x = float(vx) + px, y = float(vy) + py, z = float(vz) + pz;
                const float s = box_sample_f64(sdf, 0, double(x), double(y), double(z));
                if (keep ? (s <= 0.f) : (s >= 0.f)) {
                    out.P.push_back(fxpt16_encode(px)); out.P.push_back(fxpt16_encode(py)); out.P.push_back(fxpt16_encode(pz));
                    for (int a = 0; a < 3; a++) out.v.push_back(half_encode(half_decode(in.v[3 * gi + a])));
                    count++;
                }
            }
            out.voxelEnd[l][off] = count;
        }
    }
    out.leafBegin.push_back(out.P.size() / 3);
    w.particles = std::move(out);
}

// ---- FluidReseed (FF/nosys/FLIP_Reseed.cpp:8-16 -> FLIP_vdb::reseed_fluid, FF/FLIP_vdb.cpp:2047-2220)
// The jitter table (FF/FLIP_vdb.h:10-20): randomTable[i] = float(double(frand(i)) - 0.5), frand a 32-bit integer hash.
static inline float reseed_frand(unsigned int i) {
    unsigned int value = (i ^ 61) ^ (i >> 16);
    value *= 9;
    value ^= value << 4;
    value *= 0x27d4eb2d;
    value ^= value >> 15;
    return (float)value / (float)4294967296;
}
static inline float reseed_table(uint64_t index) { return float(double(reseed_frand((unsigned int)(index % 21474836ull))) - 0.5); }
// Where a leaf's draw sequence starts when the caller gives no explicit start. The reference takes std::random_device per TBB
// chunk (:2081-2084) and runs on through the chunk's leaves, so ANY start is one of its executions for a one-leaf chunk; the
// seeded variant (also what the CUDA path does) derives it from (seed, leaf origin), which makes leaves independent.
uint64_t reseed_leaf_start(uint32_t seed, int ox, int oy, int oz) {
    uint32_t h = seed ^ ((uint32_t)ox * 73856093u) ^ ((uint32_t)oy * 19349663u) ^ ((uint32_t)oz * 83492791u);
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return (uint64_t)(h % 21474836u);
}
// One leaf of seed_leaf2 (:2087-2209). Per voxel in offset order: the voxel's particles are re-emitted through the write handles
// (position decode -> encode once more, as in KillParticlesInSDF above), their octants (p > 0 per axis, :2128-2129) marked; then,
// if the liquid SDF at the voxel centre is < dx and the voxel holds <= 4 particles (:2146), up to 16 trials while it holds < 8:
// a jitter from three consecutive table entries, rejected when the SDF at that point is > -dx (:2157) or its octant is taken
// (:2164), otherwise appended with the StaggeredBoxSampler velocity at that point (:2173). Every trial consumes three entries.
// Both grids share the particles' cell-centred transform: indexToWorld = ijk * s, worldToIndex = xyz * (1 / s) in double
// (math/Maps.h ScaleMap::applyMap / applyInverseMap), which does NOT round-trip to the integer for every ijk -- restated literally.
static uint64_t reseed_leaf(const World& w, int l, uint64_t index, std::vector<uint16_t>& P, std::vector<uint16_t>& V,
                            std::array<uint32_t, 512>& ends) {
    const Points& in = w.particles;
    const float dx = w.dx;
    const double s = double(dx), inv = 1.0 / s;
    const Coord o = in.origins[l];
    uint32_t emitted = 0;
    for (int off = 0; off < 512; off++) {
        const uint32_t b = off ? in.voxelEnd[l][off - 1] : 0u, e = in.voxelEnd[l][off];
        unsigned occ = 0;
        uint32_t here = 0;
        for (uint32_t i = b; i < e; i++) {
            const size_t gi = in.leafBegin[l] + i;
            const float px = fxpt16_decode(in.P[3 * gi]), py = fxpt16_decode(in.P[3 * gi + 1]), pz = fxpt16_decode(in.P[3 * gi + 2]);
            occ |= 1u << (((pz > 0.f) << 2) | ((py > 0.f) << 1) | (px > 0.f));
            P.push_back(fxpt16_encode(px)); P.push_back(fxpt16_encode(py)); P.push_back(fxpt16_encode(pz));
            for (int a = 0; a < 3; a++) V.push_back(half_encode(half_decode(in.v[3 * gi + a])));
            emitted++; here++;
        }
        const int vx = o.x + (off >> 6), vy = o.y + ((off >> 3) & 7), vz = o.z + (off & 7);
        const double wx = double(vx) * s, wy = double(vy) * s, wz = double(vz) * s;
        const float phi = box_sample_f64(w.liquidSDF, 0, wx * inv, wy * inv, wz * inv);
        if (phi < dx && here <= 4) {
            for (int trial = 0; here < 8 && trial < 16; trial++) {
                const float jx = reseed_table(index++), jy = reseed_table(index++), jz = reseed_table(index++);
                const double qx = double(jx) * s + wx, qy = double(jy) * s + wy, qz = double(jz) * s + wz;
                const double ix = qx * inv, iy = qy * inv, iz = qz * inv;
                const float phi2 = box_sample_f64(w.liquidSDF, 0, ix, iy, iz);
                if (phi2 > -dx) continue;
                const unsigned sv = ((double(jz) > 0) << 2) | ((double(jy) > 0) << 1) | (double(jx) > 0);
                if (occ & (1u << sv)) continue;
                occ |= 1u << sv;
                const float vel[3] = {box_sample_f64(w.velocity, 0, ix + 0.5, iy, iz), box_sample_f64(w.velocity, 1, ix, iy + 0.5, iz),
                                      box_sample_f64(w.velocity, 2, ix, iy, iz + 0.5)};
                P.push_back(fxpt16_encode(jx)); P.push_back(fxpt16_encode(jy)); P.push_back(fxpt16_encode(jz));
                for (int a = 0; a < 3; a++) V.push_back(half_encode(vel[a]));
                emitted++; here++;
            }
        }
        ends[off] = emitted;
    }
    return index;
}
// leafStart: one draw-sequence start per particle leaf (store order), or nullptr for reseed_leaf_start(seed, origin);
// leafEnd (optional) receives where each leaf's sequence ended -- what a following leaf of the same TBB chunk starts from.
void node_FluidReseed(World& w, uint32_t seed, const uint64_t* leafStart, uint64_t* leafEnd) {
    const Points& in = w.particles;
    Points out;
    out.dir = in.dir;
    out.origins = in.origins;
    out.voxelEnd.resize(in.leafCount());
    for (int l = 0; l < in.leafCount(); l++) {
        out.leafBegin.push_back(out.P.size() / 3);
        const Coord o = in.origins[l];
        const uint64_t start = leafStart ? leafStart[l] : reseed_leaf_start(seed, o.x, o.y, o.z);
        const uint64_t end = reseed_leaf(w, l, start, out.P, out.v, out.voxelEnd[l]);
        if (leafEnd) leafEnd[l] = end;
    }
    out.leafBegin.push_back(out.P.size() / 3);
    w.particles = std::move(out);
}

// ---- ParticleEmitter (FF/nosys/ParticleEmitter.cpp:9-40 -> FLIP_vdb::emit_liquid, FF/FLIP_vdb.cpp:2222-2642), the branch
// without a velocity volume (:2488-2624): new particles carry the constant (vx, vy, vz).
// Leaves: every particle-index leaf box one of whose 9^3 lattice corners samples the shape SDF < 0 (:2270-2312; the loop range,
// the shape's leaf bounding box +- 8 voxels, only has to be a superset: outside it the shape reads its positive background) --
// existing leaves are reused, missing ones created. Per voxel of such a leaf, in offset order: the particles stay (re-emitted
// through the write handles); if the shape at the voxel centre is < dx, up to 16 trials while the voxel holds < 8: three table
// entries per trial, skipped when its octant is taken (:2577-2581), taken when the shape at the candidate is < -0.1 dx (:2585).
// Leaves the shape does not touch are not visited at all (their positions are NOT re-encoded). The shape is sampled through
// shape.worldToIndex(particles.indexToWorld(.)); here both share the cell-centred transform of voxel size dx (the restriction the
// device path has; ScaleMap arithmetic as in reseed_leaf).
static bool emitter_touches(const FloatGrid& shape, double s, double inv, int ox, int oy, int oz) {
    for (int ii = 0; ii <= 8; ii++)
        for (int jj = 0; jj <= 8; jj++)
            for (int kk = 0; kk <= 8; kk++) {
                const double wx = double(ox + ii) * s, wy = double(oy + jj) * s, wz = double(oz + kk) * s;
                if (box_sample_f64(shape, 0, wx * inv, wy * inv, wz * inv) < 0) return true;
            }
    return false;
}
void node_ParticleEmitter(World& w, const FloatGrid& shape, float vx, float vy, float vz, uint32_t seed, const uint64_t* leafStart,
                          uint64_t* leafEnd) {
    const Points& in = w.particles;
    const float dx = w.dx;
    const double s = double(dx), inv = 1.0 / s;
    const float thr = float(double(-dx) * 0.1);
    // candidate leaves: the shape's leaves and their 26 neighbours
    std::map<uint64_t, Coord> cand;
    for (const Coord& o : shape.origins)
        for (int a = -8; a <= 8; a += 8)
            for (int b = -8; b <= 8; b += 8)
                for (int c = -8; c <= 8; c += 8) cand.emplace(leafKeyOf(o.x + a, o.y + b, o.z + c), Coord(o.x + a, o.y + b, o.z + c));
    std::map<uint64_t, std::pair<Coord, bool>> leaves;   // key -> (origin, touched by the shape); std::map = the store's leaf order
    for (int l = 0; l < in.leafCount(); l++) leaves[leafKeyOf(in.origins[l].x, in.origins[l].y, in.origins[l].z)] = {in.origins[l], false};
    for (auto& kv : cand)
        if (emitter_touches(shape, s, inv, kv.second.x, kv.second.y, kv.second.z)) {
            auto it = leaves.find(kv.first);
            if (it == leaves.end()) leaves[kv.first] = {kv.second, true};
            else it->second.second = true;
        }
    Points out;
    int li = 0;
    for (auto& kv : leaves) {
        const Coord o = kv.second.first;
        const int l = in.findLeaf(o.x, o.y, o.z);
        out.dir.emplace(kv.first, li);
        out.origins.push_back(o);
        out.voxelEnd.emplace_back();
        out.leafBegin.push_back(out.P.size() / 3);
        std::array<uint32_t, 512>& ends = out.voxelEnd.back();
        if (!kv.second.second) {   // untouched: copied as it is
            const size_t b = in.leafBegin[l], e = in.leafBegin[l + 1];
            out.P.insert(out.P.end(), in.P.begin() + 3 * b, in.P.begin() + 3 * e);
            out.v.insert(out.v.end(), in.v.begin() + 3 * b, in.v.begin() + 3 * e);
            ends = in.voxelEnd[l];
            if (leafEnd) leafEnd[li] = ~0ull;
            li++;
            continue;
        }
        uint64_t index = leafStart ? leafStart[li] : reseed_leaf_start(seed, o.x, o.y, o.z);
        uint32_t emitted = 0;
        for (int off = 0; off < 512; off++) {
            unsigned occ = 0;
            uint32_t here = 0;
            if (l >= 0) {
                const uint32_t b = off ? in.voxelEnd[l][off - 1] : 0u, e = in.voxelEnd[l][off];
                for (uint32_t i = b; i < e; i++) {
                    const size_t gi = in.leafBegin[l] + i;
                    const float px = fxpt16_decode(in.P[3 * gi]), py = fxpt16_decode(in.P[3 * gi + 1]), pz = fxpt16_decode(in.P[3 * gi + 2]);
                    occ |= 1u << (((pz > 0.f) << 2) | ((py > 0.f) << 1) | (px > 0.f));
                    out.P.push_back(fxpt16_encode(px)); out.P.push_back(fxpt16_encode(py)); out.P.push_back(fxpt16_encode(pz));
                    for (int a = 0; a < 3; a++) out.v.push_back(half_encode(half_decode(in.v[3 * gi + a])));
                    emitted++; here++;
                }
            }
            const int ix = o.x + (off >> 6), iy = o.y + ((off >> 3) & 7), iz = o.z + (off & 7);
            const double wx = double(ix) * s, wy = double(iy) * s, wz = double(iz) * s;
            if (box_sample_f64(shape, 0, wx * inv, wy * inv, wz * inv) < dx) {
                for (int trial = 0; here < 8 && trial < 16; trial++) {
                    const float jx = reseed_table(index++), jy = reseed_table(index++), jz = reseed_table(index++);
                    const unsigned sv = ((double(jz) > 0) << 2) | ((double(jy) > 0) << 1) | (double(jx) > 0);
                    if (occ & (1u << sv)) continue;
                    const double qx = double(jx) * s + wx, qy = double(jy) * s + wy, qz = double(jz) * s + wz;
                    if (box_sample_f64(shape, 0, qx * inv, qy * inv, qz * inv) < thr) {
                        occ |= 1u << sv;
                        out.P.push_back(fxpt16_encode(jx)); out.P.push_back(fxpt16_encode(jy)); out.P.push_back(fxpt16_encode(jz));
                        out.v.push_back(half_encode(vx)); out.v.push_back(half_encode(vy)); out.v.push_back(half_encode(vz));
                        emitted++; here++;
                    }
                }
            }
            ends[off] = emitted;
        }
        if (leafEnd) leafEnd[li] = index;
        li++;
    }
    out.leafBegin.push_back(out.P.size() / 3);
    w.particles = std::move(out);
}

// VDBPointsToPrimitive (projects/zenvdb/GetVDBPoints.cpp:76-258): per particle, in store order, the world position
// Vec3f(indexToWorld(Vec3d(decoded P) + Vec3d(voxel))) -- the sum and the scaling in double, one rounding to float -- and the
// decoded velocity. (The reference walks leaves in tree order; a primitive is an unordered point set.)
void node_VDBPointsToPrimitive(const World& w, float* pos, float* vel) {
    const Points& in = w.particles;
    const double s = double(w.dx);
    size_t k = 0;
    for (int l = 0; l < in.leafCount(); l++) {
        const Coord o = in.origins[l];
        for (int off = 0; off < 512; off++) {
            const uint32_t b = off ? in.voxelEnd[l][off - 1] : 0u, e = in.voxelEnd[l][off];
            const int c[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
            for (uint32_t i = b; i < e; i++, k++) {
                const size_t gi = in.leafBegin[l] + i;
                for (int a = 0; a < 3; a++) {
                    pos[3 * k + a] = float((double(fxpt16_decode(in.P[3 * gi + a])) + double(c[a])) * s);
                    if (vel) vel[3 * k + a] = half_decode(in.v[3 * gi + a]);
                }
            }
        }
    }
}

// FLIP_vdb::point_integrate_vector, channel "vel" (FF/FLIP_vdb.cpp:3492-3535): v is read as Vec3R (half -> float -> double),
// dv (a Vec3R built from the node's float vec3) is added in double, and the sum goes back through the Vec3f write handle:
// double -> float (round to nearest) -> half (TruncateCodec, round to nearest even). Positions are untouched.
void node_ParticleAddDV(World& w, float dvx, float dvy, float dvz) {
    const double dv[3] = {double(dvx), double(dvy), double(dvz)};
    Points& p = w.particles;
    for (size_t i = 0; i < p.size(); i++)
        for (int a = 0; a < 3; a++) {
            const double v = double(half_decode(p.v[3 * i + a])) + dv[a];
            p.v[3 * i + a] = half_encode(float(v));
        }
}

}  // namespace orc
