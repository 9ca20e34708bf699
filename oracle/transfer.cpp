// TEST INFRASTRUCTURE ONLY -- see flip_oracle.h.
// Codecs, binning (K1), P2G (K3/K4), extrapolation (K5), vec3 packing (K6).
#include "flip_oracle.h"
#include <cmath>
#include <numeric>

namespace orc {

World::World(float dx_) : dx(dx_) {
    // FF/nosys/FLIP_Creator.cpp:37-116 backgrounds
    liquidSDF = FloatGrid(1.0f * dx_);
    solidSDF = FloatGrid(3.0f * dx_);
    pressure = FloatGrid(0.f);
    divergence = FloatGrid(0.f);
    curvature = FloatGrid(0.f);
}

// ---------------------------------------------------------------- codecs
// FixedPointCodec<false, PositionRange>::encode: value + 0.5 then
// floatingPointToFixedPoint<uint16_t> (AttributeArray.h:47-55,966-976,478-483)
uint16_t fxpt16_encode(float p) {
    float s = p + 0.5f;
    if (0.0f > s) return 0;
    else if (1.0f <= s) return 65535;
    return uint16_t(s * 65535.0f);
}
// decode: float(u)/float(65535) - 0.5 (AttributeArray.h:58-65,953-962)
float fxpt16_decode(uint16_t u) { return float(u) / 65535.0f - 0.5f; }

// TruncateCodec on Vec3<half>: float -> half round-to-nearest-even with
// denormals and overflow to inf (math/Half.h:430-490, Half.cc:87-215).
uint16_t half_encode(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t e = int32_t((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x007fffffu;
    if (((x >> 23) & 0xff) == 0xff) {  // inf / nan
        if (m == 0) return uint16_t(sign | 0x7c00u);
        m >>= 13;
        return uint16_t(sign | 0x7c00u | m | (m == 0));
    }
    if (e <= 0) {
        if (e < -10) return uint16_t(sign);  // underflow to signed zero
        m = m | 0x00800000u;
        int t = 14 - e;
        uint32_t a = (1u << (t - 1)) - 1;
        uint32_t b = (m >> t) & 1;
        m = (m + a + b) >> t;
        return uint16_t(sign | m);
    }
    m = m + 0x00000fffu + ((m >> 13) & 1);
    if (m & 0x00800000u) { m = 0; e += 1; }
    if (e > 30) return uint16_t(sign | 0x7c00u);  // overflow -> inf
    return uint16_t(sign | (uint32_t(e) << 10) | (m >> 13));
}
float half_decode(uint16_t h) {
    uint32_t sign = uint32_t(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1f;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            int ee = -1;
            do { ee++; m <<= 1; } while ((m & 0x400u) == 0);
            m &= 0x3ffu;
            x = sign | (uint32_t(127 - 15 - ee) << 23) | (m << 13);
        }
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 127 - 15) << 23) | (m << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

// ---------------------------------------------------------------- K6
// packed_FloatGrid3::from_vec3 (projects/zenvdb/include/zeno/packed3grids.cpp:11-46)
void from_vec3(Packed3& out, const Vec3Grid& in, bool topologyOnly) {
    for (int i = 0; i < 3; i++) {
        out.v[i] = FloatGrid(in.bg[i]);
        out.v[i].topologyCopyFrom(in);
    }
    if (topologyOnly) return;
    for (int l = 0; l < in.leafCount(); l++)
        for (int i = 0; i < 3; i++) {
            const float* src = in.leafVals(l, i);
            float* dst = out.v[i].leafVals(l, 0);
            for (int off = 0; off < 512; off++)
                if (maskGet(in.masks[l], off)) dst[off] = src[off];
        }
}
// packed_FloatGrid3::to_vec3 (packed3grids.cpp:49-83): new tree bg 0, union topology,
// each channel's ON voxels copy their value.
void to_vec3(Vec3Grid& out, const Packed3& in) {
    out = Vec3Grid(0.f);
    for (int i = 0; i < 3; i++) out.topologyUnion(in.v[i]);
    for (int i = 0; i < 3; i++)
        for (int l = 0; l < in.v[i].leafCount(); l++) {
            int m = out.findLeaf(in.v[i].origins[l]);
            const float* src = in.v[i].leafVals(l, 0);
            float* dst = out.leafVals(m, i);
            for (int off = 0; off < 512; off++)
                if (maskGet(in.v[i].masks[l], off)) dst[off] = src[off];
        }
}

// ---------------------------------------------------------------- K1
// particleArrayToGrid (projects/zenvdb/SetVDBPointDataGrid.cpp:17-72):
// worldToIndex = double(pos) * (1.0/double(dx)) (math/Maps.h:688,751-753),
// ijk = Coord::round = floor(x+0.5) (math/Coord.h:50-53, math/Math.h:822-823),
// P_local = float(idx - ijk) through the fxpt16 codec (points/PointConversion.h:700-718),
// v through the half codec; stable voxel bucketing (tools/PointPartitioner.h:9).
void bin_from_points(World& w, const float* pos, const float* vel, size_t n) {
    Points& P = w.particles;
    P.clear();
    const double inv = 1.0 / double(w.dx);
    std::vector<uint64_t> key(n);
    std::vector<uint16_t> off(n);
    std::vector<std::array<int, 3>> ijk(n);
    for (size_t i = 0; i < n; i++) {
        int c[3];
        for (int a = 0; a < 3; a++) {
            double idx = double(pos[3 * i + a]) * inv;
            c[a] = int(std::floor(idx + 0.5));
        }
        ijk[i] = {c[0], c[1], c[2]};
        key[i] = leafKeyOf(c[0], c[1], c[2]);
        off[i] = uint16_t(voxelOffset(c[0], c[1], c[2]));
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (key[a] != key[b]) return key[a] < key[b];
        return off[a] < off[b];
    });
    P.P.resize(3 * n);
    P.v.resize(3 * n);
    for (size_t s = 0; s < n; s++) {
        uint32_t i = order[s];
        int l;
        auto it = P.dir.find(key[i]);
        if (it == P.dir.end()) {
            l = P.leafCount();
            P.dir.emplace(key[i], l);
            P.origins.push_back(Coord(ijk[i][0] & ~7, ijk[i][1] & ~7, ijk[i][2] & ~7));
            std::array<uint32_t, 512> z; z.fill(0);
            P.voxelEnd.push_back(z);
            P.leafBegin.push_back(s);
        } else l = it->second;
        P.voxelEnd[l][off[i]]++;
        for (int a = 0; a < 3; a++) {
            double idx = double(pos[3 * i + a]) * inv;
            float local = float(idx - double(ijk[i][a]));
            P.P[3 * s + a] = fxpt16_encode(local);
            P.v[3 * s + a] = half_encode(vel ? vel[3 * i + a] : 0.f);
        }
    }
    P.leafBegin.push_back(n);
    for (auto& ve : P.voxelEnd)
        for (int o = 1; o < 512; o++) ve[o] += ve[o - 1];
}

// ---------------------------------------------------------------- K5
// vdb_velocity_extrapolator::union_extrapolate (FF/vdb_velocity_extrapolator.cpp:584-661)
void union_extrapolate(int nLayer, Packed3& v, const FloatGrid* targetTopo) {
    Grid<1> unionTopo(0.f);
    if (!targetTopo) {
        for (int i = 0; i < 3; i++) unionTopo.topologyUnion(v.v[i]);
        unionTopo.dilate(nLayer, /*nn26=*/false);
    } else {
        unionTopo.topologyCopyFrom(*targetTopo);
    }
    for (int ch = 0; ch < 3; ch++) {
        FloatGrid& g = v.v[ch];
        Grid<1> extra = unionTopo;  // deepCopy
        extra.topologyDifference(g);
        std::vector<int> chLeaf(extra.leafCount());
        for (int l = 0; l < extra.leafCount(); l++) chLeaf[l] = g.touchLeaf(extra.origins[l]);
        for (int layer = 0; layer < nLayer; layer++) {
            // valid_vel_topo: snapshot of the channel topology at the start of the layer
            std::vector<Mask512> valid = g.masks;
            auto validOn = [&](const Coord& c) {
                int l = g.findLeaf(c);
                if (l < 0 || l >= int(valid.size())) return false;
                return maskGet(valid[l], voxelOffset(c.x, c.y, c.z));
            };
            for (int l = 0; l < extra.leafCount(); l++) {
                Coord o = extra.origins[l];
                for (int off = 0; off < 512; off++) {
                    if (!maskGet(extra.masks[l], off)) continue;
                    Coord c(o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
                    int tw = 0;
                    float sum = 0;
                    for (int d = 0; d < 6; d++) {
                        int dir = d / 2, positive = d % 2;
                        Coord nb = c;
                        if (positive) nb[dir]++; else nb[dir]--;
                        if (validOn(nb)) { tw++; sum += g.get(nb); }
                    }
                    if (tw != 0) {
                        g.leafVals(chLeaf[l])[off] = sum / tw;
                        maskSet(g.masks[chLeaf[l]], off, true);
                        maskSet(extra.masks[l], off, false);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- K3/K4
namespace {
// p2g_collector::operator() (FF/FLIP_vdb.cpp:1137-1263) for one velocity leaf.
void p2g_collect_leaf(const Points& pts, const Coord& origin, float dx, float radius,
                      float* vel[3], float* wgt[3], float* sdf, Mask512& sdfMask) {
    // fill_particle_leafs (FF/FLIP_vdb.cpp:1004-1015)
    int nb[27];
    {
        int c = 0;
        for (int ii = -8; ii <= 8; ii += 8)
            for (int jj = -8; jj <= 8; jj += 8)
                for (int kk = -8; kk <= 8; kk += 8)
                    nb[c++] = pts.findLeaf(origin.x + ii, origin.y + jj, origin.z + kk);
    }
    // all_particle_iterator (FF/FLIP_vdb.cpp:1021-1132): centre voxel runs over [-1,8]^3, z fastest
    for (int cx = -1; cx <= 8; cx++)
        for (int cy = -1; cy <= 8; cy++)
            for (int cz = -1; cz <= 8; cz++) {
                int lp = (cx < 0 ? 0 : (cx < 8 ? 1 : 2)) * 9 + (cy < 0 ? 0 : (cy < 8 ? 1 : 2)) * 3 +
                         (cz < 0 ? 0 : (cz < 8 ? 1 : 2));
                int pl = nb[lp];
                if (pl < 0) continue;
                int off = ((cx & 7) << 6) | ((cy & 7) << 3) | (cz & 7);
                uint32_t end = pts.voxelEnd[pl][off];
                uint32_t beg = off == 0 ? 0u : pts.voxelEnd[pl][off - 1];
                for (uint32_t it = beg; it < end; it++) {
                    size_t gi = pts.leafBegin[pl] + it;
                    float px = fxpt16_decode(pts.P[3 * gi + 0]);
                    float py = fxpt16_decode(pts.P[3 * gi + 1]);
                    float pz = fxpt16_decode(pts.P[3 * gi + 2]);
                    float pv[3] = {half_decode(pts.v[3 * gi + 0]), half_decode(pts.v[3 * gi + 1]),
                                   half_decode(pts.v[3 * gi + 2])};
                    for (int iv = 0; iv < 27; iv++) {
                        int bx = iv / 9 - 1, by = (iv / 3) % 3 - 1, bz = iv % 3 - 1;
                        int hx = cx + bx, hy = cy + by, hz = cz + bz;
                        if (hx < 0 || hy < 0 || hz < 0 || hx > 7 || hy > 7 || hz > 7) continue;
                        int wo = (hx << 6) | (hy << 3) | hz;
                        // lane order of the SSE pack: (phi, w, v, u); offsets set at :851-876
                        float fx = float(bx), fy = float(by), fz = float(bz);
                        float tx = std::fabs(fx - px), ty = std::fabs(fy - py), tz = std::fabs(fz - pz);
                        float dist = dx * std::sqrt(tx * tx + ty * ty + tz * tz);
                        sdf[wo] = std::min(sdf[wo], dist - radius);
                        maskSet(sdfMask, wo, true);
                        float xs = std::fabs((fx + -0.5f) - px);  // staggered distances
                        float ys = std::fabs((fy + -0.5f) - py);
                        float zs = std::fabs((fz + -0.5f) - pz);
                        float dxc[3] = {xs, tx, tx}, dyc[3] = {ty, ys, ty}, dzc[3] = {tz, tz, zs};
                        // the three early-outs (:1206-1217) only skip all-zero updates
                        for (int c = 0; c < 3; c++) {
                            float wx = std::max(0.f, 1.0f - dxc[c]);
                            float wy = std::max(0.f, 1.0f - dyc[c]);
                            float wz = std::max(0.f, 1.0f - dzc[c]);
                            float wgtc = (wx * wy) * wz;
                            vel[c][wo] = pv[c] * wgtc + vel[c][wo];
                            wgt[c][wo] = wgtc + wgt[c][wo];
                        }
                    }
                }
            }
}
}  // namespace

// FLIP_vdb::particle_to_grid_collect_style (FF/FLIP_vdb.cpp:1282-1388)
static void particle_to_grid_collect_style(World& w, Packed3& outVel, Packed3& outVelAfter,
                                           FloatGrid& outSdf, float dx) {
    const Points& pts = w.particles;
    float radius = dx * 0.8f * 1.01f;
    // velocity topology = occupied voxels dilated once, 26-neighbourhood (:1290-1296)
    Vec3Grid unweighted(0.f);
    for (int l = 0; l < pts.leafCount(); l++) {
        int m = unweighted.touchLeaf(pts.origins[l]);
        for (int off = 0; off < 512; off++) {
            uint32_t end = pts.voxelEnd[l][off];
            uint32_t beg = off == 0 ? 0u : pts.voxelEnd[l][off - 1];
            if (end > beg) maskSet(unweighted.masks[m], off, true);
        }
    }
    unweighted.dilate(1, true);
    Vec3Grid weights = unweighted;  // deepCopy (:1299)
    float sdfBg = outSdf.bg[0];
    outSdf = FloatGrid(sdfBg);
    outSdf.topologyCopyFrom(unweighted);  // (:1301-1303)

#pragma omp parallel for schedule(dynamic, 4)
    for (int l = 0; l < unweighted.leafCount(); l++) {
        float* vel[3] = {unweighted.leafVals(l, 0), unweighted.leafVals(l, 1), unweighted.leafVals(l, 2)};
        float* wg[3] = {weights.leafVals(l, 0), weights.leafVals(l, 1), weights.leafVals(l, 2)};
        p2g_collect_leaf(pts, unweighted.origins[l], dx, radius, vel, wg, outSdf.leafVals(l), outSdf.masks[l]);
    }

    // from_vec3(topology only) + normalize_p2g_velocity (:1318-1322, :120-165)
    from_vec3(outVel, unweighted, true);
    for (int l = 0; l < unweighted.leafCount(); l++)
        for (int c = 0; c < 3; c++) {
            float* dst = outVel.v[c].leafVals(l);
            for (int off = 0; off < 512; off++) {
                if (!maskGet(unweighted.masks[l], off)) continue;
                float weight = weights.leafVals(l, c)[off];
                if (weight == 0) {
                    dst[off] = 0.f;
                    maskSet(outVel.v[c].masks[l], off, false);
                } else {
                    dst[off] = unweighted.leafVals(l, c)[off] / (weight + 0.001f);
                    maskSet(outVel.v[c].masks[l], off, true);
                }
            }
        }

    // air one-ring (:1325-1382)
    Grid<1> airmask(0.f);
    airmask.topologyCopyFrom(outSdf);
    airmask.dilate(1, true);
    airmask.topologyDifference(outSdf);
    outSdf.dilate(1, true);
    // values read through the accessor are the pre-deduction ones wherever they can be
    // negative (air voxels only ever receive positive values), so a snapshot is equivalent.
    for (int l = 0; l < outSdf.leafCount(); l++) {
        int al = airmask.findLeaf(outSdf.origins[l]);
        if (al < 0) continue;
        Coord o = outSdf.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(airmask.masks[al], off)) continue;
            Coord c(o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
            bool hasLiquid = false;
            float newSdf = outSdf.leafVals(l)[off];
            for (int i = 0; i < 6; i++) {
                int comp = i / 2;
                bool pos = (i % 2 == 0);
                Coord at = c;
                if (pos) at[comp]++; else at[comp]--;
                float nsdf = outSdf.get(at);
                if (nsdf < 0) {
                    hasLiquid = true;
                    newSdf = std::min(newSdf, dx + nsdf);
                }
            }
            if (!hasLiquid) maskSet(outSdf.masks[l], off, false);
            else {
                outSdf.leafVals(l)[off] = newSdf;
                maskSet(outSdf.masks[l], off, true);
            }
        }
    }
    outVelAfter = outVel;  // deepCopy (:1386)
}

// FLIP_P2G::apply (FF/nosys/P2G.cpp:11-42)
void node_FLIP_P2G(World& w, float dx, int velExtraLayer) {
    Packed3 vel, post;
    from_vec3(vel, w.velocity, false);
    from_vec3(post, w.postAdvVelocity, false);
    particle_to_grid_collect_style(w, vel, post, w.liquidSDF, dx);
    union_extrapolate(velExtraLayer, vel, &w.liquidSDF);
    to_vec3(w.velocity, vel);
    to_vec3(w.postAdvVelocity, post);
}

}  // namespace orc
