// TEST INFRASTRUCTURE ONLY -- link stubs for the two legacy solvers FLIP_vdb.cpp references but the
// FastFLIP nodes on the hot path never call: simd_vdb_poisson (FLIP_vdb::solve_pressure_simd, disabled
// with #if 0 in FF/nosys/SolvePoissonPressureEqn.cpp:46-51) and simd_viscosity3d
// (FLIP_vdb::solve_viscosity, the SolveViscousTerm node: SURVEY 8f "next"). oracle/_ref does not compile
// FF/simd_vdb_poisson.cpp / FF/simd_viscosity3d.cpp (they need far more of Eigen); calling into them aborts.
#include "simd_vdb_poisson.h"
#include "simd_viscosity3d.h"
#include <stdexcept>

static void notBuilt(const char* what) { throw std::runtime_error(std::string("oracle/_ref: ") + what + " is not part of the hot path and was not built"); }

void simd_vdb_poisson::construct_levels() { notBuilt("simd_vdb_poisson"); }
void simd_vdb_poisson::build_rhs() { notBuilt("simd_vdb_poisson"); }
openvdb::FloatGrid::Ptr simd_vdb_poisson::Laplacian_with_level::get_zero_vec_grid() { notBuilt("simd_vdb_poisson"); return nullptr; }
bool simd_vdb_poisson::pcg_solve(openvdb::FloatGrid::Ptr, float) { notBuilt("simd_vdb_poisson"); return false; }
void simd_vdb_poisson::smooth_solve(openvdb::FloatGrid::Ptr, int) { notBuilt("simd_vdb_poisson"); }

namespace simd_uaamg {
simd_viscosity3d::simd_viscosity3d(openvdb::FloatGrid::Ptr, openvdb::FloatGrid::Ptr, openvdb::FloatGrid::Ptr, packed_FloatGrid3,
                                   openvdb::Vec3fGrid::Ptr, float, float) { notBuilt("simd_viscosity3d"); }
packed_FloatGrid3 L_with_level::get_zero_vec() const { notBuilt("simd_viscosity3d"); return packed_FloatGrid3(); }
void simd_viscosity3d::pcg_solve(packed_FloatGrid3, float) { notBuilt("simd_viscosity3d"); }
}
