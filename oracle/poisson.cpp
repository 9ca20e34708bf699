// TEST INFRASTRUCTURE ONLY -- see flip_oracle.h.
// UAAMG pressure Poisson solver (K9-K13): matrix-free 7-point variational Laplacian on
// 8^3 leaves, multigrid hierarchy, red-black Gauss-Seidel smoothing, MGPCG.
// Follows FF/simd_vdb_poisson_uaamg.cpp; explicit FMAs are kept where the reference
// uses _mm256_fmadd_ps so the op sequence is the same.
#include "flip_oracle.h"
#include <cmath>
#include <cstdio>
#include <memory>

namespace orc {
namespace {

struct Level {
    // DOF topology: leaves + DOF mask; dofIndex value per voxel (-1 off)
    std::unordered_map<uint64_t, int> dir;
    std::vector<Coord> origins;
    std::vector<Mask512> dof;
    std::vector<std::array<int, 512>> dofIndex;
    // coefficients, dense per leaf; trimmed flags say "reads as the default"
    std::vector<float> diag, xe, ye, ze, invdiag;  // [leaf][512]
    std::vector<uint8_t> trimDiag, trimX, trimY, trimZ;
    std::vector<std::array<int, 6>> nbr;  // x+ x- y+ y- z+ z- (uaamg.cpp:1382-1394), -1 = none
    float dt, dx, term;                    // term = dt/dx^2
    int numDof = 0;
    int level = 0;

    int leafCount() const { return int(origins.size()); }
    int findLeaf(int x, int y, int z) const {
        auto it = dir.find(leafKeyOf(x, y, z));
        return it == dir.end() ? -1 : it->second;
    }
    bool dofOn(int x, int y, int z) const {
        int l = findLeaf(x, y, z);
        return l >= 0 && maskGet(dof[l], voxelOffset(x, y, z));
    }
    // accessor reads honouring trimmed leaves (missing leaf -> grid background = default)
    float getDiag(int x, int y, int z) const {
        int l = findLeaf(x, y, z);
        if (l < 0 || trimDiag[l]) return 6.0f * term;
        return diag[size_t(l) * 512 + voxelOffset(x, y, z)];
    }
    float getFace(int ch, int x, int y, int z) const {
        int l = findLeaf(x, y, z);
        const std::vector<float>& e = ch == 0 ? xe : (ch == 1 ? ye : ze);
        const std::vector<uint8_t>& t = ch == 0 ? trimX : (ch == 1 ? trimY : trimZ);
        if (l < 0 || t[l]) return -term;
        return e[size_t(l) * 512 + voxelOffset(x, y, z)];
    }
    int addLeaf(const Coord& o) {
        uint64_t k = leafKeyOf(o.x, o.y, o.z);
        auto it = dir.find(k);
        if (it != dir.end()) return it->second;
        int id = leafCount();
        dir.emplace(k, id);
        origins.push_back(Coord(o.x & ~7, o.y & ~7, o.z & ~7));
        Mask512 m; m.fill(0);
        dof.push_back(m);
        return id;
    }
};

using Vec = std::vector<float>;  // [leaf][512] aligned with Level leaves; non-DOF lanes are 0

// LaplacianWithLevel::setDofIndex (uaamg.cpp:786-820)
void setDofIndex(Level& L) {
    L.dofIndex.resize(L.leafCount());
    int next = 0;
    for (int l = 0; l < L.leafCount(); l++) {
        L.dofIndex[l].fill(-1);
        for (int off = 0; off < 512; off++)
            if (maskGet(L.dof[l], off)) L.dofIndex[l][off] = next++;
    }
    L.numDof = next;
}

// LaplacianWithLevel::trimDefaultNodes (uaamg.cpp:1905-1960): a coefficient leaf whose
// ACTIVE values all lie within |default|*1e-5 of the default is deleted -> reads as default.
void trimOne(const Level& L, std::vector<float>& g, std::vector<uint8_t>& flag, float defVal, float eps,
             const std::vector<Mask512>& active) {
    eps = std::fabs(eps);
    flag.assign(L.leafCount(), 0);
    for (int l = 0; l < L.leafCount(); l++) {
        float maxErr = 0;
        for (int off = 0; off < 512; off++)
            if (maskGet(active[l], off)) maxErr = std::max(maxErr, std::fabs(g[size_t(l) * 512 + off] - defVal));
        if (maxErr <= eps) {
            flag[l] = 1;
            std::fill(g.begin() + size_t(l) * 512, g.begin() + size_t(l + 1) * 512, defVal);
        }
    }
}
void trimDefaultNodes(Level& L) {
    float term = -L.dt / (L.dx * L.dx);
    float diagonalTerm = -6.0f * term;
    float epsilon = 1e-5f;
    trimOne(L, L.diag, L.trimDiag, diagonalTerm, diagonalTerm * epsilon, L.dof);
    trimOne(L, L.xe, L.trimX, term, term * epsilon, L.dof);
    trimOne(L, L.ye, L.trimY, term, term * epsilon, L.dof);
    trimOne(L, L.ze, L.trimZ, term, term * epsilon, L.dof);
}

// LaplacianApplySIMD::initLinearizedLeaves + initInvDiagonal (uaamg.cpp:1222-1253,1341-1394)
void initApply(Level& L) {
    L.nbr.resize(L.leafCount());
    for (int l = 0; l < L.leafCount(); l++) {
        Coord o = L.origins[l];
        for (int f = 0; f < 6; f++) {
            int comp = f / 2;
            bool pos = (f % 2 == 0);
            Coord n = o;
            n[comp] += pos ? 8 : -8;
            L.nbr[l][f] = L.findLeaf(n.x, n.y, n.z);
        }
    }
    L.invdiag.resize(L.diag.size());
    float defInv = 1.0f / (6.f * L.term);
    for (int l = 0; l < L.leafCount(); l++)
        for (int off = 0; off < 512; off++) {
            size_t i = size_t(l) * 512 + off;
            if (L.trimDiag[l]) { L.invdiag[i] = defInv; continue; }
            // invdiag tree background 1/bg; ON voxels 1/diag (0 if diag==0)
            if (maskGet(L.dof[l], off)) L.invdiag[i] = L.diag[i] == 0 ? 0.f : 1.0f / L.diag[i];
            else L.invdiag[i] = 1.0f / (6.0f * L.term);
        }
}

// BuildFinestMatrix (uaamg.cpp:623-747) + LaplacianWithLevel ctor (:357-402)
std::unique_ptr<Level> buildFinest(const World& w, float dt, float dx) {
    auto Lp = std::make_unique<Level>();
    Level& L = *Lp;
    L.dt = dt; L.dx = dx; L.term = dt / (dx * dx); L.level = 0;
    const FloatGrid& phi = w.liquidSDF;
    const Vec3Grid& fw = w.faceWeight;
    for (int l = 0; l < phi.leafCount(); l++) L.addLeaf(phi.origins[l]);
    size_t n = size_t(L.leafCount()) * 512;
    L.diag.assign(n, 6 * L.term);
    L.xe.assign(n, -L.term);
    L.ye.assign(n, -L.term);
    L.ze.assign(n, -L.term);
    float dtOverDxSqr = dt / (dx * dx);
    for (int l = 0; l < phi.leafCount(); l++) {
        Coord o = phi.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(phi.masks[l], off)) continue;
            float phiHere = phi.leafVals(l)[off];
            if (!(phiHere < 0.f)) continue;
            Coord g(o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
            float diagonal = 0.f;
            float xyz[3] = {0.f, 0.f, 0.f};
            for (int f = 0; f < 6; f++) {
                int comp = f / 2;
                bool pos = (f % 2 == 0);
                Coord wc = g, pc = g;
                float weight, phiOther;
                if (pos) {
                    wc[comp]++;
                    weight = fw.get(comp, wc);
                    pc[comp]++;
                    phiOther = phi.get(pc);
                } else {
                    weight = fw.get(comp, wc);
                    pc[comp]--;
                    phiOther = phi.get(pc);
                }
                float term = weight * dtOverDxSqr;
                if (phiOther < 0) {
                    diagonal += term;
                    if (!pos) xyz[comp] = -term;
                } else {
                    float theta = fraction_inside(phiHere, phiOther);
                    if (theta < 0.02f) theta = 0.02f;
                    diagonal += term / theta;
                }
            }
            if (diagonal != 0.f) {
                size_t i = size_t(l) * 512 + off;
                maskSet(L.dof[l], off, true);
                L.diag[i] = diagonal;
                L.xe[i] = xyz[0];
                L.ye[i] = xyz[1];
                L.ze[i] = xyz[2];
            }
        }
    }
    setDofIndex(L);
    trimDefaultNodes(L);
    initApply(L);
    return Lp;
}

// LaplacianWithLevel::initializeFromFineLevel (uaamg.cpp:409-620)
std::unique_ptr<Level> coarsen(const Level& F) {
    auto Lp = std::make_unique<Level>();
    Level& L = *Lp;
    L.dt = F.dt; L.dx = 2.0f * F.dx; L.level = F.level + 1;
    L.term = L.dt / (L.dx * L.dx);
    // TouchCoarseLeafReducer (FF/SIMD_UAAMG_Ops.h:4-37): touchLeaf(origin/2)
    for (int l = 0; l < F.leafCount(); l++) {
        Coord o = F.origins[l];
        L.addLeaf(Coord(o.x / 2, o.y / 2, o.z / 2));
    }
    for (int l = 0; l < L.leafCount(); l++) {
        Coord o = L.origins[l];
        for (int off = 0; off < 512; off++) {
            int cx = o.x + (off >> 6), cy = o.y + ((off >> 3) & 7), cz = o.z + (off & 7);
            bool found = false;
            for (int ii = 0; ii < 2 && !found; ii++)
                for (int jj = 0; jj < 2 && !found; jj++)
                    for (int kk = 0; kk < 2 && !found; kk++)
                        if (F.dofOn(2 * cx + ii, 2 * cy + jj, 2 * cz + kk)) found = true;
            maskSet(L.dof[l], off, found);
        }
    }
    setDofIndex(L);
    float dtOverDxSqr = L.dt / (L.dx * L.dx);
    size_t n = size_t(L.leafCount()) * 512;
    L.diag.assign(n, 6.0f * dtOverDxSqr);
    L.xe.assign(n, -dtOverDxSqr);
    L.ye.assign(n, -dtOverDxSqr);
    L.ze.assign(n, -dtOverDxSqr);
    for (int l = 0; l < L.leafCount(); l++) {
        Coord o = L.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(L.dof[l], off)) continue;
            int cx = o.x + (off >> 6), cy = o.y + ((off >> 3) & 7), cz = o.z + (off & 7);
            float diag = 0, x = 0, y = 0, z = 0;
            for (int ii = 0; ii < 2; ii++)
                for (int jj = 0; jj < 2; jj++)
                    for (int kk = 0; kk < 2; kk++) {
                        int fx = 2 * cx + ii, fy = 2 * cy + jj, fz = 2 * cz + kk;
                        if (!F.dofOn(fx, fy, fz)) continue;
                        diag += F.getDiag(fx, fy, fz);
                        if (F.dofOn(fx - 1, fy, fz)) {
                            if (ii == 0) x += F.getFace(0, fx, fy, fz);
                            else diag += 2 * F.getFace(0, fx, fy, fz);
                        }
                        if (F.dofOn(fx, fy - 1, fz)) {
                            if (jj == 0) y += F.getFace(1, fx, fy, fz);
                            else diag += 2 * F.getFace(1, fx, fy, fz);
                        }
                        if (F.dofOn(fx, fy, fz - 1)) {
                            if (kk == 0) z += F.getFace(2, fx, fy, fz);
                            else diag += 2 * F.getFace(2, fx, fy, fz);
                        }
                    }
            const float factor = 0.5f * (1.0f / 8.0f);
            size_t i = size_t(l) * 512 + off;
            L.diag[i] = diag * factor;
            L.xe[i] = x * factor;
            L.ye[i] = y * factor;
            L.ze[i] = z * factor;
        }
    }
    trimDefaultNodes(L);
    initApply(L);
    return Lp;
}

enum class Mode { Laplacian, Residual, RedGS, BlackGS };

// LaplacianApplySIMD::LightWeightApplier::operator() (uaamg.cpp:904-1180), scalar per lane.
// result may alias lhs (in-place Gauss-Seidel, :1536-1559).
void applyOp(const Level& L, Mode mode, Vec& result, const Vec& lhs, const Vec* rhs, float wSor) {
    const float defFace = -L.term, defDiag = 6.f * L.term, defInv = 1.0f / (6.f * L.term);
    const float oneMinusW = 1.0f - wSor;
    const int nl = L.leafCount();
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++) {
        const float* x0 = &lhs[size_t(l) * 512];
        const float* xn[6];
        for (int f = 0; f < 6; f++) xn[f] = L.nbr[l][f] >= 0 ? &lhs[size_t(L.nbr[l][f]) * 512] : nullptr;
        auto faceLeaf = [&](const std::vector<float>& e, const std::vector<uint8_t>& t, int leaf) -> const float* {
            if (leaf < 0 || t[leaf]) return nullptr;
            return &e[size_t(leaf) * 512];
        };
        const float* xT0 = faceLeaf(L.xe, L.trimX, l);
        const float* xT1 = faceLeaf(L.xe, L.trimX, L.nbr[l][0]);
        const float* yT0 = faceLeaf(L.ye, L.trimY, l);
        const float* yT1 = faceLeaf(L.ye, L.trimY, L.nbr[l][2]);
        const float* zT0 = faceLeaf(L.ze, L.trimZ, l);
        const float* zT1 = faceLeaf(L.ze, L.trimZ, L.nbr[l][4]);
        // mXLeavesUpper probes the coefficient tree at origin+8, which exists independently
        // of the DOF leaf; on this path both trees have the same leaves (before trimming).
        const float* dg = L.trimDiag[l] ? nullptr : &L.diag[size_t(l) * 512];
        const float* idg = L.trimDiag[l] ? nullptr : &L.invdiag[size_t(l) * 512];
        const float* b = rhs ? &(*rhs)[size_t(l) * 512] : nullptr;
        float* out = &result[size_t(l) * 512];
        for (int row = 0; row < 64; row++) {
            int vo = row * 8;
            unsigned rowMask = unsigned((L.dof[l][row >> 3] >> ((row & 7) * 8)) & 0xffu);
            if (rowMask == 0) continue;
            int x = row >> 3, y = row & 7;
            float res[8];
            for (int z = 0; z < 8; z++) {
                int off = vo + z;
                float xp, xpT, xm, xmT, yp, ypT, ym, ymT, zp, zpT, zm, zmT;
                if (x < 7) { xp = x0[off + 64]; xpT = xT0 ? xT0[off + 64] : defFace; }
                else { xp = xn[0] ? xn[0][off - 448] : 0.f; xpT = xT1 ? xT1[off - 448] : defFace; }
                if (x > 0) xm = x0[off - 64]; else xm = xn[1] ? xn[1][off + 448] : 0.f;
                xmT = xT0 ? xT0[off] : defFace;
                if (y < 7) { yp = x0[off + 8]; ypT = yT0 ? yT0[off + 8] : defFace; }
                else { yp = xn[2] ? xn[2][off - 56] : 0.f; ypT = yT1 ? yT1[off - 56] : defFace; }
                if (y > 0) ym = x0[off - 8]; else ym = xn[3] ? xn[3][off + 56] : 0.f;
                ymT = yT0 ? yT0[off] : defFace;
                if (z < 7) { zp = x0[off + 1]; zpT = zT0 ? zT0[off + 1] : defFace; }
                else { zp = xn[4] ? xn[4][off - 7] : 0.f; zpT = zT1 ? zT1[off - 7] : defFace; }
                if (z > 0) zm = x0[off - 1]; else zm = xn[5] ? xn[5][off + 7] : 0.f;
                zmT = zT0 ? zT0[off] : defFace;
                // (:1044-1049) ((x+ + x-) + (y+ + y-)) + (z+ + z-)
                float fx = xp * xpT + xm * xmT;
                float fy = yp * ypT + ym * ymT;
                float fz = zp * zpT + zm * zmT;
                float offd = (fx + fy) + fz;
                float xi = x0[off];
                float r = xi;
                switch (mode) {
                case Mode::Residual: {
                    float d = dg ? dg[off] : defDiag;
                    r = b[off] - std::fmaf(xi, d, offd);
                    break;
                }
                case Mode::Laplacian: {
                    float d = dg ? dg[off] : defDiag;
                    r = std::fmaf(xi, d, offd);
                    break;
                }
                case Mode::RedGS:
                case Mode::BlackGS: {
                    float inv = idg ? idg[off] : defInv;
                    float t = ((b[off] - offd) * inv) * wSor;
                    float upd = std::fmaf(xi, oneMinusW, t);
                    bool red = ((x + y + z) & 1) == 0;
                    bool take = (mode == Mode::RedGS) ? red : !red;
                    r = take ? upd : xi;
                    break;
                }
                }
                res[z] = r;
            }
            // in-place safety: the row's own inputs were read before the row is stored,
            // and same-colour neighbours are never read for the colour being written.
            for (int z = 0; z < 8; z++) out[vo + z] = ((rowMask >> z) & 1u) ? res[z] : 0.f;
        }
    }
}

// In-place GS needs care: a row's z-neighbours inside the same row and y/x neighbours in
// other rows are of the other colour, so sequential in-place update equals the SIMD one.
void rbgs(const Level& L, Vec& x, const Vec& b, bool redFirst, float wSor) {
    applyOp(L, redFirst ? Mode::RedGS : Mode::BlackGS, x, x, &b, wSor);
    applyOp(L, redFirst ? Mode::BlackGS : Mode::RedGS, x, x, &b, wSor);
}

// restriction (uaamg.cpp:1773-1833)
void restriction(const Level& fine, const Level& coarse, Vec& outCoarse, const Vec& inFine) {
    for (int l = 0; l < coarse.leafCount(); l++) {
        Coord o = coarse.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(coarse.dof[l], off)) continue;
            int cx = o.x + (off >> 6), cy = o.y + ((off >> 3) & 7), cz = o.z + (off & 7);
            int fl = fine.findLeaf(2 * cx, 2 * cy, 2 * cz);
            if (fl < 0) continue;
            int fbase = voxelOffset(2 * cx, 2 * cy, 2 * cz);
            float sum = 0;
            for (int ii = 0; ii < 2; ii++)
                for (int jj = 0; jj < 2; jj++)
                    for (int kk = 0; kk < 2; kk++) {
                        int fo = fbase + 64 * ii + 8 * jj + kk;
                        if (maskGet(fine.dof[fl], fo)) sum += inFine[size_t(fl) * 512 + fo];
                    }
            outCoarse[size_t(l) * 512 + off] = sum * 0.125f;
        }
    }
}
// prolongation<inplace_add=true> (uaamg.cpp:1835-1903)
void prolongationAdd(const Level& fine, const Level& coarse, Vec& outFine, const Vec& inCoarse, float alpha) {
    for (int l = 0; l < coarse.leafCount(); l++) {
        Coord o = coarse.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(coarse.dof[l], off)) continue;
            int cx = o.x + (off >> 6), cy = o.y + ((off >> 3) & 7), cz = o.z + (off & 7);
            int fl = fine.findLeaf(2 * cx, 2 * cy, 2 * cz);
            if (fl < 0) continue;
            float cv = alpha * inCoarse[size_t(l) * 512 + off];
            int fbase = voxelOffset(2 * cx, 2 * cy, 2 * cz);
            for (int ii = 0; ii < 2; ii++)
                for (int jj = 0; jj < 2; jj++)
                    for (int kk = 0; kk < 2; kk++) {
                        int fo = fbase + 64 * ii + 8 * jj + kk;
                        if (maskGet(fine.dof[fl], fo)) outFine[size_t(fl) * 512 + fo] += cv;
                    }
        }
    }
}

struct Solver {
    std::vector<std::unique_ptr<Level>> levels;
    std::vector<Vec> lhs, rhs, tmp;
    // coarsest level as an explicit CSR matrix (getTriplets, uaamg.cpp:278-355)
    std::vector<int> rowPtr, col;
    std::vector<float> val, cdiagInv;

    void construct() {
        const int maxCoarsestDOF = 4000;  // uaamg.cpp:1965
        while (levels.back()->numDof > maxCoarsestDOF) levels.push_back(coarsen(*levels.back()));
        for (auto& L : levels) {
            size_t n = size_t(L->leafCount()) * 512;
            lhs.emplace_back(n, 0.f);
            rhs.emplace_back(n, 0.f);
            tmp.emplace_back(n, 0.f);
        }
        buildCoarsest();
    }
    void buildCoarsest() {
        const Level& L = *levels.back();
        int n = L.numDof;
        rowPtr.assign(n + 1, 0);
        col.clear(); val.clear();
        cdiagInv.assign(n, 1.f);
        // row-wise assembly in column order (Eigen sorts inner indices in setFromTriplets)
        for (int l = 0; l < L.leafCount(); l++) {
            Coord o = L.origins[l];
            for (int off = 0; off < 512; off++) {
                if (!maskGet(L.dof[l], off)) continue;
                int gx = o.x + (off >> 6), gy = o.y + ((off >> 3) & 7), gz = o.z + (off & 7);
                int row = L.dofIndex[l][off];
                std::vector<std::pair<int, float>> ent;
                float d = L.getDiag(gx, gy, gz);
                ent.emplace_back(row, d);
                for (int ch = 0; ch < 3; ch++) {
                    Coord nc(gx, gy, gz), pc(gx, gy, gz);
                    nc[ch]--; pc[ch]++;
                    int nl = L.findLeaf(nc.x, nc.y, nc.z);
                    if (nl >= 0 && maskGet(L.dof[nl], voxelOffset(nc.x, nc.y, nc.z)))
                        ent.emplace_back(L.dofIndex[nl][voxelOffset(nc.x, nc.y, nc.z)], L.getFace(ch, gx, gy, gz));
                    int pl = L.findLeaf(pc.x, pc.y, pc.z);
                    if (pl >= 0 && maskGet(L.dof[pl], voxelOffset(pc.x, pc.y, pc.z)))
                        ent.emplace_back(L.dofIndex[pl][voxelOffset(pc.x, pc.y, pc.z)], L.getFace(ch, pc.x, pc.y, pc.z));
                }
                std::sort(ent.begin(), ent.end());
                for (auto& e : ent) { col.push_back(e.first); val.push_back(e.second); }
                rowPtr[row + 1] = int(col.size());
                cdiagInv[row] = d != 0.f ? 1.0f / d : 1.0f;
            }
        }
    }
    // Eigen::ConjugateGradient<SparseMatrix<float>> default (Lower|Upper, DiagonalPreconditioner,
    // tolerance = float epsilon), setMaxIterations(10), zero initial guess (uaamg.cpp:2291-2303,
    // :2019-2023). Eigen is an external dependency absent from /root/reference (CI pins vcpkg eigen3,
    // Docker 3.3.7); this follows Eigen's published conjugate_gradient(): PARITY UNPINNED here.
    void coarsestSolve(Vec& outLhs, const Vec& inRhs) {
        const Level& L = *levels.back();
        int n = L.numDof;
        std::vector<float> b(n, 0.f), x(n, 0.f);
        for (int l = 0; l < L.leafCount(); l++)
            for (int off = 0; off < 512; off++)
                if (maskGet(L.dof[l], off)) b[L.dofIndex[l][off]] = inRhs[size_t(l) * 512 + off];
        auto spmv = [&](const std::vector<float>& in, std::vector<float>& out) {
            for (int r = 0; r < n; r++) {
                float s = 0;
                for (int k = rowPtr[r]; k < rowPtr[r + 1]; k++) s += val[k] * in[col[k]];
                out[r] = s;
            }
        };
        auto dot = [&](const std::vector<float>& a, const std::vector<float>& c) {
            float s = 0;
            for (int i = 0; i < n; i++) s += a[i] * c[i];
            return s;
        };
        const int maxIters = 10;
        const float tol = std::numeric_limits<float>::epsilon();
        std::vector<float> residual = b, p(n), z(n), t(n);
        float rhsNorm2 = dot(b, b);
        if (rhsNorm2 != 0) {
            float threshold = std::max(tol * tol * rhsNorm2, std::numeric_limits<float>::min());
            float residualNorm2 = dot(residual, residual);
            if (!(residualNorm2 < threshold)) {
                for (int i = 0; i < n; i++) p[i] = cdiagInv[i] * residual[i];
                float absNew = dot(residual, p);
                int i = 0;
                while (i < maxIters) {
                    spmv(p, t);
                    float alpha = absNew / dot(p, t);
                    for (int k = 0; k < n; k++) x[k] += alpha * p[k];
                    for (int k = 0; k < n; k++) residual[k] -= alpha * t[k];
                    residualNorm2 = dot(residual, residual);
                    if (residualNorm2 < threshold) break;
                    for (int k = 0; k < n; k++) z[k] = cdiagInv[k] * residual[k];
                    float absOld = absNew;
                    absNew = dot(residual, z);
                    float beta = absNew / absOld;
                    for (int k = 0; k < n; k++) p[k] = z[k] + beta * p[k];
                    i++;
                }
            }
        }
        for (int l = 0; l < L.leafCount(); l++)
            for (int off = 0; off < 512; off++)
                if (maskGet(L.dof[l], off)) outLhs[size_t(l) * 512 + off] = x[L.dofIndex[l][off]];
    }

    // muCyclePreconditioner<mu=2, skip_first> with RBGS (uaamg.cpp:1993-2126).
    // x and b are the level's lhs/rhs storage (level 0: caller's vectors).
    void muCyclePrecond(Vec& x, const Vec& b, int level, int n, bool skipFirst) {
        const int nlevel = int(levels.size());
        const Level& L = *levels[level];
        if (level == nlevel - 1) { coarsestSolve(x, b); return; }
        const float w = 1.2f;  // mWeightSOR (uaamg.cpp:1201)
        // setGridToResultAfterFirstRBGS (:1665-1731) == a red-first sweep on a zero guess
        // (the non-DOF lanes and the first colour are overwritten; identical arithmetic).
        if (skipFirst) { std::fill(x.begin(), x.end(), 0.f); rbgs(L, x, b, true, w); }
        for (int i = (skipFirst ? 1 : 0); i < n; i++) rbgs(L, x, b, true, w);
        applyOp(L, Mode::Residual, tmp[level], x, &b, w);
        int parent = level + 1;
        restriction(L, *levels[parent], rhs[parent], tmp[level]);
        muCyclePrecond(lhs[parent], rhs[parent], parent, n, true);
        muCyclePrecond(lhs[parent], rhs[parent], parent, n, false);
        prolongationAdd(L, *levels[parent], x, lhs[parent], 1.0f);
        for (int i = 0; i < n; i++) rbgs(L, x, b, false, w);
    }

    // muCycleIterative<2> with RBGS, w = 1.0 (uaamg.cpp:2127-2288)
    void muCycleIter(Vec& x, const Vec& b, int level, int n, int postSmooth) {
        const int nlevel = int(levels.size());
        const Level& L = *levels[level];
        const float w = 1.0f;
        if (level == nlevel - 1) {
            for (int i = 0; i < 10 * n; i++) rbgs(L, x, b, true, w);
            return;
        }
        for (int i = 0; i < n; i++) rbgs(L, x, b, true, w);
        applyOp(L, Mode::Residual, tmp[level], x, &b, w);
        int parent = level + 1;
        restriction(L, *levels[parent], rhs[parent], tmp[level]);
        // the reference does not reset the parent lhs here (uaamg.cpp:2225-2227): it carries
        // whatever the previous visit / preconditioner call left in mMuCycleLHSs[parent].
        for (int mu = 0; mu < 2; mu++) muCycleIter(lhs[parent], rhs[parent], parent, n, 0);
        prolongationAdd(L, *levels[parent], x, lhs[parent], 0.5f);
        for (int i = 0; i < n; i++) rbgs(L, x, b, false, w);
        for (int i = 0; i < postSmooth && level == 0; i++) rbgs(L, x, b, false, w);
    }
};

// grid_abs_max_op / grid_dot_op (FF/openvdb_grid_math_op.h:4-74); fp32, leaf order.
float absMax(const Level& L, const Vec& v) {
    float m = 0;
    for (int l = 0; l < L.leafCount(); l++)
        for (int off = 0; off < 512; off++)
            if (maskGet(L.dof[l], off)) {
                float a = v[size_t(l) * 512 + off];
                if (!std::isfinite(a)) return a;
                m = std::max(m, std::fabs(a));
            }
    return m;
}
float dotp(const Level& L, const Vec& a, const Vec& b) {
    // per-leaf partial sums then a leaf-order combine, like the TBB reduce over leaf ranges
    float total = 0;
    for (int l = 0; l < L.leafCount(); l++) {
        float s = 0;
        for (int off = 0; off < 512; off++)
            if (maskGet(L.dof[l], off)) s += a[size_t(l) * 512 + off] * b[size_t(l) * 512 + off];
        total += s;
    }
    return total;
}
void axpy(const Level& L, float alpha, const Vec& x, Vec& y) {  // y = a*x + y (uaamg.cpp:2463-2479)
    for (int l = 0; l < L.leafCount(); l++)
        for (int off = 0; off < 512; off++)
            if (maskGet(L.dof[l], off)) y[size_t(l) * 512 + off] += alpha * x[size_t(l) * 512 + off];
}
void xpay(const Level& L, float alpha, const Vec& x, Vec& y) {  // y = x + a*y (uaamg.cpp:2481-2497)
    for (int l = 0; l < L.leafCount(); l++)
        for (int off = 0; off < 512; off++)
            if (maskGet(L.dof[l], off))
                y[size_t(l) * 512 + off] = x[size_t(l) * 512 + off] + alpha * y[size_t(l) * 512 + off];
}

// BuildPoissonRhs (uaamg.cpp:18-93); with SurfaceTension > 0 BuildPoissonRhs_withTension (uaamg.cpp:95-209): a face towards an
// air cell adds dt/dx^2 * weight * tension * (theta curvOther + (1 - theta) curvThis) / theta, tension = 2 coef / density
void buildRhs(const World& w, const Level& L, const Packed3& vel, Vec& rhs, float dx, float dt) {
    float invdx = 1.0f / dx;
    const bool enableTension = w.tensionCoef > 0;
    const float tension = 2 * w.tensionCoef / w.density;
    const float dtOverDxSqr = dt / (dx * dx);
    for (int l = 0; l < L.leafCount(); l++) {
        Coord o = L.origins[l];
        for (int off = 0; off < 512; off++) {
            if (!maskGet(L.dof[l], off)) continue;
            Coord g(o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
            float r = 0, weightSum = 0;
            bool hasNonZero = false;
            const float phiThis = enableTension ? w.liquidSDF.get(g) : 0.f, curvThis = enableTension ? w.curvature.get(g) : 0.f;
            for (int i = 0; i < 6; i++) {
                int ch = i / 2;
                bool pos = (i % 2) == 0;
                Coord nc = g;
                nc[ch] += int(pos);
                float weight = w.faceWeight.get(ch, nc);
                weightSum += weight;
                if (weight != 0.f) hasNonZero = true;
                float v = vel.v[ch].get(nc);
                float sv = w.solidVelocity.get(ch, nc);
                if (pos) r -= invdx * (weight * v + (1.0f - weight) * sv);
                else r += invdx * (weight * v + (1.0f - weight) * sv);
                if (enableTension) {
                    Coord pc = g;
                    pc[ch] += pos ? 1 : -1;
                    const float phiOther = w.liquidSDF.get(pc), curvOther = w.curvature.get(pc);
                    if (phiThis < 0.f && phiOther >= 0.f) {
                        float theta = fraction_inside(phiThis, phiOther);
                        if (theta < 0.02f) theta = 0.02f;
                        r += dtOverDxSqr * weight * tension * (theta * curvOther + (1.f - theta) * curvThis) / theta;
                    }
                }
            }
            if (!hasNonZero || weightSum < 0.1) r = 0;
            rhs[size_t(l) * 512 + off] = r;
        }
    }
}

}  // namespace

// AssembleSolvePPE::apply (FF/nosys/SolvePoissonPressureEqn.cpp:23-64) ->
// FLIP_vdb::solve_pressure_simd_uaamg (FF/FLIP_vdb.cpp:3034-3100), tension disabled.
void node_AssembleSolvePPE(World& w, float dt, float dx) {
    Packed3 vel;
    from_vec3(vel, w.velocity, false);
    w.residualHistory.clear();
    w.pcgIterations = 0;
    w.pcgStatus = 0;
    if (w.liquidSDF.leafCount() == 0) { to_vec3(w.velocity, vel); return; }

    Solver S;
    S.levels.push_back(buildFinest(w, dt, dx));
    S.construct();
    const Level& L0 = *S.levels[0];
    w.mgLevels = int(S.levels.size());
    w.numDof = L0.numDof;
    size_t n = size_t(L0.leafCount()) * 512;
    Vec rhs(n, 0.f), pressure(n, 0.f), r(n, 0.f), p(n, 0.f), z(n, 0.f);
    buildRhs(w, L0, vel, rhs, dx, dt);

    // solveMultigridPCG (uaamg.cpp:2332-2403); mRelativeTolerance 5e-5, mMaxIteration 100
    const float relTol = w.solveRelTol;
    const int maxIter = w.solveMaxIter;
    int status = 1, iter = 0;
    applyOp(L0, Mode::Residual, r, pressure, &rhs, 1.2f);
    float nu = absMax(L0, r);
    float initAbs = nu + 1e-16f;
    float numax = relTol * nu;
    w.residualHistory.push_back(nu / initAbs);
    if (nu <= numax) status = 0;
    else {
        S.muCyclePrecond(p, r, 0, 4, true);
        float rho = dotp(L0, p, r);
        float nuOld = nu;
        for (; iter < maxIter; iter++) {
            applyOp(L0, Mode::Laplacian, z, p, nullptr, 1.2f);
            float sigma = dotp(L0, p, z);
            float alpha = rho / sigma;
            axpy(L0, -alpha, z, r);
            nuOld = nu;
            nu = absMax(L0, r);
            w.residualHistory.push_back(nu / initAbs);
            if (nu <= numax) { axpy(L0, alpha, p, pressure); status = 0; break; }
            if (nu > nuOld && iter > 3) { status = 1; break; }
            S.muCyclePrecond(z, r, 0, 4, true);
            float rhoNew = dotp(L0, z, r);
            float beta = rhoNew / rho;
            rho = rhoNew;
            axpy(L0, alpha, p, pressure);
            xpay(L0, beta, z, p);
        }
    }
    w.pcgIterations = iter;
    w.pcgStatus = status;
    if (status != 0) {
        // fallback: warm start from the previous pressure + pure multigrid (FF/FLIP_vdb.cpp:3089-3097)
        for (int l = 0; l < L0.leafCount(); l++) {
            Coord o = L0.origins[l];
            for (int off = 0; off < 512; off++) {
                if (!maskGet(L0.dof[l], off)) continue;
                float oldp = w.pressure.get(0, o.x + (off >> 6), o.y + ((off >> 3) & 7), o.z + (off & 7));
                if (std::isfinite(oldp)) pressure[size_t(l) * 512 + off] = oldp;
            }
        }
        applyOp(L0, Mode::Residual, r, pressure, &rhs, 1.0f);
        nu = absMax(L0, r);
        numax = relTol * nu;
        if (!(nu <= numax)) {
            for (int it2 = 0; it2 < 100; it2++) {
                S.muCycleIter(pressure, rhs, 0, 8, 8);
                applyOp(L0, Mode::Residual, r, pressure, &rhs, 1.0f);
                nu = absMax(L0, r);
                if (nu <= numax) break;
            }
        }
    }
    w.pcgRelResidual = nu / initAbs;

    // outputs: Pressure = new grid on the DOF topology; Divergence = rhs grid (:3063-3067,3087,3099)
    FloatGrid pg(0.f), rg(0.f);
    for (int l = 0; l < L0.leafCount(); l++) {
        int a = pg.touchLeaf(L0.origins[l]);
        int b = rg.touchLeaf(L0.origins[l]);
        pg.masks[a] = L0.dof[l];
        rg.masks[b] = L0.dof[l];
        std::copy(pressure.begin() + size_t(l) * 512, pressure.begin() + size_t(l + 1) * 512, pg.leafVals(a));
        std::copy(rhs.begin() + size_t(l) * 512, rhs.begin() + size_t(l + 1) * 512, rg.leafVals(b));
    }
    w.pressure = std::move(pg);
    w.divergence = std::move(rg);
    to_vec3(w.velocity, vel);
}

}  // namespace orc
