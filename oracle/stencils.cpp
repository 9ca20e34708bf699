// TEST INFRASTRUCTURE ONLY -- see flip_oracle.h.
// Small single-pass stencil stages (K14) and level-set fractions.
#include "flip_oracle.h"
#include <cmath>

namespace orc {

// FF/levelset_util.cpp:5-15
float fraction_inside(float phi_left, float phi_right) {
    if (phi_left < 0 && phi_right < 0) return 1;
    if (phi_left < 0 && phi_right >= 0) return phi_left / (phi_left - phi_right);
    if (phi_left >= 0 && phi_right < 0) return phi_right / (phi_right - phi_left);
    else return 0;
}
static void cycle_array(float* arr, int size) {
    float t = arr[0];
    for (int i = 0; i < size - 1; ++i) arr[i] = arr[i + 1];
    arr[size - 1] = t;
}
// FF/levelset_util.cpp:26-99
float fraction_inside(float phi_bl, float phi_br, float phi_tl, float phi_tr) {
    int inside_count = (phi_bl < 0 ? 1 : 0) + (phi_tl < 0 ? 1 : 0) + (phi_br < 0 ? 1 : 0) + (phi_tr < 0 ? 1 : 0);
    float list[] = {phi_bl, phi_br, phi_tr, phi_tl};
    if (inside_count == 4) return 1;
    else if (inside_count == 3) {
        while (list[0] < 0) cycle_array(list, 4);
        float side0 = 1 - fraction_inside(list[0], list[3]);
        float side1 = 1 - fraction_inside(list[0], list[1]);
        return 1 - 0.5f * side0 * side1;
    } else if (inside_count == 2) {
        while (list[0] >= 0 || !(list[1] < 0 || list[2] < 0)) cycle_array(list, 4);
        if (list[1] < 0) {
            float side_left = fraction_inside(list[0], list[3]);
            float side_right = fraction_inside(list[1], list[2]);
            return 0.5f * (side_left + side_right);
        } else {
            float middle_point = 0.25f * (list[0] + list[1] + list[2] + list[3]);
            if (middle_point < 0) {
                float area = 0;
                float side1 = 1 - fraction_inside(list[0], list[3]);
                float side3 = 1 - fraction_inside(list[2], list[3]);
                area += 0.5f * side1 * side3;
                float side2 = 1 - fraction_inside(list[2], list[1]);
                float side0 = 1 - fraction_inside(list[0], list[1]);
                area += 0.5f * side0 * side2;
                return 1 - area;
            } else {
                float area = 0;
                float side0 = fraction_inside(list[0], list[1]);
                float side1 = fraction_inside(list[0], list[3]);
                area += 0.5f * side0 * side1;
                float side2 = fraction_inside(list[2], list[1]);
                float side3 = fraction_inside(list[2], list[3]);
                area += 0.5f * side2 * side3;
                return area;
            }
        }
    } else if (inside_count == 1) {
        while (list[0] >= 0) cycle_array(list, 4);
        float side0 = fraction_inside(list[0], list[3]);
        float side1 = fraction_inside(list[0], list[1]);
        return 0.5f * side0 * side1;
    } else return 0;
}

// FLIP_vdb::calculate_face_weights (FF/FLIP_vdb.cpp:2644-2718)
void node_CutCellWeight(World& w) {
    Vec3Grid fw(1.0f);
    fw.topologyCopyFrom(w.liquidSDF);
    fw.dilate(1, true);
    for (int l = 0; l < fw.leafCount(); l++) {
        Coord o = fw.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(fw.masks[l], off)) continue;
            int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
            float s[2][2][2];
            for (int ii = 0; ii < 2; ii++)
                for (int jj = 0; jj < 2; jj++)
                    for (int kk = 0; kk < 2; kk++) s[ii][jj][kk] = w.solidSDF.get(0, x + ii, y + jj, z + kk);
            float u = 1.0f - fraction_inside(s[0][0][0], s[0][1][0], s[0][0][1], s[0][1][1]);
            u = std::max(0.f, std::min(u, 1.f));
            float v = 1.0f - fraction_inside(s[0][0][0], s[0][0][1], s[1][0][0], s[1][0][1]);
            v = std::max(0.f, std::min(v, 1.f));
            float ww = 1.0f - fraction_inside(s[0][0][0], s[1][0][0], s[0][1][0], s[1][1][0]);
            ww = std::max(0.f, std::min(ww, 1.f));
            fw.leafVals(l, 0)[off] = u;
            fw.leafVals(l, 1)[off] = v;
            fw.leafVals(l, 2)[off] = ww;
        }
    }
    w.faceWeight = std::move(fw);
}

namespace {
// openvdb BoxSampler with double weights on a float grid (Interpolation.h:712-737,763-778)
float box_sample_f64(const FloatGrid& g, double x, double y, double z) {
    int bx = int(std::floor(x)), by = int(std::floor(y)), bz = int(std::floor(z));
    double u = x - bx, v = y - by, wz = z - bz;
    float d[2][2][2];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) d[i][j][k] = g.get(0, bx + i, by + j, bz + k);
    auto ip = [](float a, float b, double wt) { return a + float(double(b - a) * wt); };
    return ip(ip(ip(d[0][0][0], d[0][0][1], wz), ip(d[0][1][0], d[0][1][1], wz), v),
              ip(ip(d[1][0][0], d[1][0][1], wz), ip(d[1][1][0], d[1][1][1], wz), v), u);
}
}  // namespace

// FLIP_vdb::immerse_liquid_phi_in_solids (FF/FLIP_vdb.cpp:2720-2803)
void node_PushOutLiquidSDF(World& w, float dx) {
    FloatGrid& phi = w.liquidSDF;
    const FloatGrid& solid = w.solidSDF;
    const int nLeafBefore = phi.leafCount();  // phi_manager is built once, before the dilation (:2738)
    for (int l = 0; l < nLeafBefore; l++) {
        Coord o = phi.origins[l];
        if (solid.findLeaf(o) < 0) continue;
        for (int off = 0; off < 512; off++) {
            if (!maskGet(phi.masks[l], off)) continue;
            int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
            float vs = box_sample_f64(solid, x + 0.5, y + 0.5, z + 0.5);
            if (vs < 0) phi.leafVals(l)[off] = phi.leafVals(l)[off] - 0.5f * dx;
        }
    }
    FloatGrid ref = phi;  // deepCopy (:2741)
    phi.dilate(1, true);
    for (int l = 0; l < nLeafBefore; l++) {
        Coord o = phi.origins[l];
        if (solid.findLeaf(o) < 0) continue;
        for (int off = 0; off < 512; off++) {
            if (!maskGet(phi.masks[l], off)) continue;
            Coord c(o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
            float vs = box_sample_f64(solid, c.x + 0.5, c.y + 0.5, c.z + 0.5);
            if (vs < 0) {
                bool found = false;
                float minFluid = dx * 3.0f;
                for (int i = 0; i < 6 && !found; i++) {
                    int comp = i / 2;
                    bool pos = (i % 2 == 0);
                    Coord t = c;
                    if (pos) t[comp]++; else t[comp]--;
                    float rv = ref.get(t);
                    minFluid = std::min(minFluid, rv);
                    found |= (rv < 0);
                }
                float* pv = &phi.leafVals(l)[off];
                if (found) *pv = minFluid;
                else if (*pv < 0) *pv = std::max(phi.bg[0], -vs);
            }
        }
    }
    phi.dilate(1, false);  // NN_FACE (:2802)
}

// FLIP_vdb::field_add_vector (FF/FLIP_vdb.cpp:3145-3158) through FieldAddVector::apply
// (FF/nosys/FieldAddVector.cpp:16-31): from_vec3 -> add on every ON voxel -> to_vec3, dt = 1.0.
void node_FieldAddVector(World& w, float x, float y, float z) {
    Packed3 p;
    from_vec3(p, w.velocity, false);
    float f[3] = {x, y, z};
    const float dt = 1.0f;
    for (int c = 0; c < 3; c++)
        for (int l = 0; l < p.v[c].leafCount(); l++) {
            float* v = p.v[c].leafVals(l);
            for (int off = 0; off < 512; off++)
                if (maskGet(p.v[c].masks[l], off)) v[off] = v[off] + f[c] * dt;
        }
    to_vec3(w.velocity, p);
}

// FLIP_vdb::cfl (FF/FLIP_vdb.cpp:3160-3207)
float node_CFL_dt(World& w) {
    const Vec3Grid& vel = w.velocity;
    std::vector<float> maxPerLeaf(vel.leafCount(), 0.f);
    for (int l = 0; l < vel.leafCount(); l++) {
        float mv = 0;
        for (int off = 0; off < 512; off++)
            if (maskGet(vel.masks[l], off))
                for (int c = 0; c < 3; c++) mv = std::max(mv, std::fabs(vel.leafVals(l, c)[off]));
        maxPerLeaf[l] = mv;
    }
    if (maxPerLeaf.empty()) return std::numeric_limits<float>::max() / 2;
    int nleaf = int(maxPerLeaf.size());
    int top90 = nleaf * 99 / 100;
    std::nth_element(maxPerLeaf.begin(), maxPerLeaf.begin() + nleaf - 1, maxPerLeaf.end());
    std::nth_element(maxPerLeaf.begin(), maxPerLeaf.begin() + top90, maxPerLeaf.end());
    // The reference reads element nleaf-1 after both partial sorts. The first call parks the
    // global maximum at nleaf-1; libstdc++'s introselect never moves an element that is greater
    // than every pivot, so the value read is the global maximum of the per-leaf maxima (the
    // "99th percentile" has no effect). The tail maximum below is that same value.
    float mv = *std::max_element(maxPerLeaf.begin() + top90, maxPerLeaf.end());
    return w.dx / (std::fabs(mv) + 1e-6f);
}

// FLIP_vdb::apply_pressure_gradient (FF/FLIP_vdb.cpp:2863-2967); with SurfaceTension > 0 the pressure of an air cell next to the
// face is replaced by tension * (curvature mixed by theta), :2932-2939
static void apply_pressure_gradient(World& w, Packed3& vel, float dt, float dx) {
    const bool enableTension = w.tensionCoef > 0;
    const float tension = 2 * w.tensionCoef / w.density;
    for (int ch = 0; ch < 3; ch++) {
        FloatGrid& g = vel.v[ch];
        for (int l = 0; l < g.leafCount(); l++) {
            Coord o = g.origins[l];
            for (int off = 0; off < 512; off++) {
                if (!maskGet(g.masks[l], off)) continue;
                Coord c(o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
                bool hasUpdate = false;
                float updated = g.leafVals(l)[off];
                float fwt = w.faceWeight.get(ch, c);
                Coord lo = c;
                lo[ch] -= 1;
                if (fwt > 0.0f) {
                    bool hp = w.pressure.isOn(c), hpb = w.pressure.isOn(lo);
                    if (hp || hpb) hasUpdate = true;
                    if (hasUpdate) {
                        float phiThis = w.liquidSDF.get(c), phiBelow = w.liquidSDF.get(lo);
                        float pThis = w.pressure.get(c), pBelow = w.pressure.get(lo);
                        float theta = 1.0f;
                        if (phiThis >= 0 || phiBelow >= 0) {
                            theta = fraction_inside(phiBelow, phiThis);
                            if (theta < 0.02f) theta = 0.02f;
                            if (enableTension) {
                                const float curvThis = w.curvature.get(c), curvBelow = w.curvature.get(lo);
                                if (phiThis >= 0) pThis = tension * (theta * curvThis + (1.f - theta) * curvBelow);
                                else if (phiBelow >= 0) pBelow = tension * (theta * curvBelow + (1.f - theta) * curvThis);
                            }
                        }
                        float velUpdate = -dt * (float)(pThis - pBelow) / dx / theta;
                        updated += velUpdate;
                        if (fwt < 1.0f) {
                            float solidVel = w.solidVelocity.get(ch, c);
                            const float boundary_friction_coef = 0.f;
                            float solidFraction = (1.0f - fwt) * boundary_friction_coef;
                            updated = (1.0f - solidFraction) * updated + (solidFraction)*solidVel;
                        }
                    }
                }
                if (hasUpdate) g.leafVals(l)[off] = updated;
                else maskSet(g.masks[l], off, false);
            }
        }
    }
}

// SubtractPressureGradient::apply (FF/nosys/SubtractPressureGradient.cpp:25-66)
void node_SubtractPressureGradient(World& w, float dt, float dx, int velExtraLayer) {
    Packed3 p;
    from_vec3(p, w.velocity, false);
    apply_pressure_gradient(w, p, dt, dx);
    union_extrapolate(velExtraLayer, p, &w.liquidSDF);
    to_vec3(w.velocity, p);
}

// ---------------------------------------------------------------- VDBRenormalizeSDF
namespace {
// math::GodunovsNormSqrd (openvdb/math/FiniteDifference.h:326-347) of the first-order one-sided differences
// (ISGradientNormSqrd<FIRST_BIAS>, math/Operators.h:249-260: up = forward, down = backward differences, index space)
inline float godunov_norm_sqrd(bool outside, const float m[3], const float p[3]) {
    auto pow2 = [](float x) { return x * x; };
    float s;
    if (outside) {
        s = std::max(pow2(std::max(m[0], 0.f)), pow2(std::min(p[0], 0.f)));
        s += std::max(pow2(std::max(m[1], 0.f)), pow2(std::min(p[1], 0.f)));
        s += std::max(pow2(std::max(m[2], 0.f)), pow2(std::min(p[2], 0.f)));
    } else {
        s = std::max(pow2(std::min(m[0], 0.f)), pow2(std::max(p[0], 0.f)));
        s += std::max(pow2(std::min(m[1], 0.f)), pow2(std::max(p[1], 0.f)));
        s += std::max(pow2(std::min(m[2], 0.f)), pow2(std::max(p[2], 0.f)));
    }
    return s;
}
// one Euler stage of Normalizer::euler<N,D> (LevelSetTracker.h:631-675): every ACTIVE voxel of `cur` gets
// alpha*phi0 + beta*v (v alone when N == 0); the stencil reads `cur` wherever it lands (inactive voxels and the background
// included); inactive voxels keep their value
void renorm_stage(const FloatGrid& cur, const std::vector<float>& phi0, int N, int D, float dt, float invDx, std::vector<float>& out) {
    out = cur.vals;
    const float alpha = D ? float(N) / float(D) : 0.f, beta = 1.0f - alpha;
    for (int l = 0; l < cur.leafCount(); l++) {
        const Coord o = cur.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(cur.masks[l], off)) continue;
            const int x = o.x + (off >> 6), y = o.y + ((off >> 3) & 7), z = o.z + (off & 7);
            const float c = cur.vals[size_t(l) * 512 + off];
            const float up[3] = {cur.get(0, x + 1, y, z) - c, cur.get(0, x, y + 1, z) - c, cur.get(0, x, y, z + 1) - c};
            const float dn[3] = {c - cur.get(0, x - 1, y, z), c - cur.get(0, x, y - 1, z), c - cur.get(0, x, y, z - 1)};
            const float n2 = godunov_norm_sqrd(c > 0.f, dn, up);
            float v = c / (std::sqrt(c * c + n2) + 1e-8f);          // math::Tolerance<float>::value() = 1e-8
            v = c - dt * v * (std::sqrt(n2) * invDx - 1.0f);
            out[size_t(l) * 512 + off] = N ? alpha * phi0[size_t(l) * 512 + off] + beta * v : v;
        }
    }
}
}  // namespace

void node_VDBRenormalizeSDF(FloatGrid& g, float voxelSize, int iterations) {
    const float dt = voxelSize * 1.0f, invDx = 1.0f / voxelSize;   // Normalizer ctor (:519-523): TVD_RK3 -> dt = dx
    std::vector<float> next;
    for (int it = 0; it < iterations; it++) {
        const std::vector<float> phi0 = g.vals;                    // aux buffer 1 after the first swap
        renorm_stage(g, phi0, 0, 1, dt, invDx, next); g.vals.swap(next);   // euler01: Phi_t1
        renorm_stage(g, phi0, 3, 4, dt, invDx, next); g.vals.swap(next);   // euler34: 3/4 Phi_t0 + 1/4 (Phi_t1 - dt V.G)
        renorm_stage(g, phi0, 1, 3, dt, invDx, next); g.vals.swap(next);   // euler13: 1/3 Phi_t0 + 2/3 (Phi_t2 - dt V.G)
    }
}

// VDBErodeSDF::apply (projects/zenvdb/VDBRenormalize.cpp:155-172)
void node_VDBErodeSDF(FloatGrid& g, float depth) {
    for (int l = 0; l < g.leafCount(); l++)
        for (int off = 0; off < 512; off++)
            if (maskGet(g.masks[l], off)) g.vals[size_t(l) * 512 + off] += depth;
}

// VDBSmoothSDF::apply -> Filter::gaussian(width, iterations, nullptr), tiles off: per iteration 4 x (box X, box Z, box Y); each pass
// writes (sum_{k=-w..w} value shifted by k along the axis, float adds in that order) * (1.f / float(2w+1)) to the ACTIVE voxels
void node_VDBSmoothSDF(FloatGrid& g, int width, int iterations) {
    if (iterations <= 0) return;
    const int w = std::max(1, width);
    const float frac = 1.f / float(2 * w + 1);
    static const int order[3] = {0, 2, 1};
    std::vector<float> next;
    for (int it = 0; it < iterations; it++)
        for (int rep = 0; rep < 4; rep++)
            for (int a = 0; a < 3; a++) {
                const int axis = order[a];
                next = g.vals;
                for (int l = 0; l < g.leafCount(); l++) {
                    const Coord o = g.origins[l];
                    for (int off = 0; off < 512; off++) {
                        if (!maskGet(g.masks[l], off)) continue;
                        int c[3] = {o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7)};
                        const int centre = c[axis];
                        float sum = 0.f;
                        for (int k = -w; k <= w; k++) { c[axis] = centre + k; sum += g.get(0, c[0], c[1], c[2]); }
                        next[size_t(l) * 512 + off] = sum * frac;
                    }
                }
                g.vals.swap(next);
            }
}

// ---- FLIPApplyBoundary (FF/nosys/Update_Solid_SDF.cpp:9-31 -> FLIP_vdb::update_solid_sdf, FF/FLIP_vdb.cpp:1976-2046) with one
// moving solid. (1) every leaf of the moving solid touches the static-SDF leaf under its origin (:1982-1990); (2) every particle
// leaf one of whose eight box corners samples the moving solid < 0 touches the static-SDF leaf at ITS origin (:1995-2024, the
// particle-index coordinate is used as it is); (3) EVERY voxel of EVERY leaf of the static SDF becomes min(own value, moving
// solid sampled at the voxel's world position) and active (:2026-2045). New leaves start at the background, inactive.
// Transforms (FF/nosys/FLIP_Creator.cpp:37-43,95-96): particles cell centred (x = i s), the static SDF vertex centred
// (x = i s + t, t = -0.5 s; inverse (x - t) (1 / s), math/Maps.h:1273-1288); the moving solid on either of the two.
void node_FLIPApplyBoundary(World& w, const FloatGrid& moving, bool movingVertexCentred) {
    const double s = double(w.dx), inv = 1.0 / s, t = -0.5 * s, mt = movingVertexCentred ? t : 0.0;
    auto moving_index = [&](double wx, double wy, double wz, double out[3]) {
        if (movingVertexCentred) { out[0] = (wx - mt) * inv; out[1] = (wy - mt) * inv; out[2] = (wz - mt) * inv; }
        else { out[0] = wx * inv; out[1] = wy * inv; out[2] = wz * inv; }
    };
    FloatGrid& solid = w.solidSDF;
    for (const Coord& o : moving.origins) {
        const double wx = movingVertexCentred ? double(o.x) * s + mt : double(o.x) * s, wy = movingVertexCentred ? double(o.y) * s + mt : double(o.y) * s,
                     wz = movingVertexCentred ? double(o.z) * s + mt : double(o.z) * s;
        solid.touchLeaf(int(std::floor((wx - t) * inv)), int(std::floor((wy - t) * inv)), int(std::floor((wz - t) * inv)));
    }
    const Points& pts = w.particles;
    for (int l = 0; l < pts.leafCount(); l++) {
        const Coord o = pts.origins[l];
        bool hit = false;
        for (int ii = 0; ii <= 8 && !hit; ii += 8)
            for (int jj = 0; jj <= 8 && !hit; jj += 8)
                for (int kk = 0; kk <= 8 && !hit; kk += 8) {
                    double ip[3];
                    moving_index(double(o.x + ii) * s, double(o.y + jj) * s, double(o.z + kk) * s, ip);
                    hit = box_sample_f64(moving, ip[0], ip[1], ip[2]) < 0;
                }
        if (hit) solid.touchLeaf(o.x, o.y, o.z);
    }
    for (int l = 0; l < solid.leafCount(); l++) {
        const Coord o = solid.origins[l];
        float* v = solid.leafVals(l);
        for (int off = 0; off < 512; off++) {
            double ip[3];
            moving_index(double(o.x + (off >> 6)) * s + t, double(o.y + ((off >> 3) & 7)) * s + t, double(o.z + (off & 7)) * s + t, ip);
            v[off] = std::min(v[off], box_sample_f64(moving, ip[0], ip[1], ip[2]));
        }
        solid.masks[l].fill(~uint64_t(0));
    }
    w.hasSolidSDF = true;
}

}  // namespace orc
