// TEST INFRASTRUCTURE ONLY -- runs the REFERENCE's own node classes on the CPU.
//
// projects/FastFLIP/nosys/{P2G,SheetG2PAdvector,EvalFaceWeight,FixLiquidSDF,FieldAddVector,CFL,SolvePoissonPressureEqn,
// SubtractPressureGradient}.cpp are compiled here UNMODIFIED, from where they lie, against the minimal stand-in of the Zeno
// node runtime (oracle/ref/shims/zeno_nodes) and the real FLIP_vdb.h; their apply() bodies call the FLIP_vdb statics already
// linked into libflipref.so. They are wrapped in a namespace (the drop-in defines node structs with the same names in
// plugin_nodes_test.cpp) and register into their own table. rn_<node> wires real OpenVDB objects to their sockets exactly like
// pn_<node> does for the drop-in (node_harness.inc).
//
// Purpose: ref_driver.cpp claims that each ref_<node> entry point "performs exactly the calls of the node shim it stands for".
// tests/test_ref_pin_cpu.py::test_ref_driver_equals_reference_nodes checks that claim against the node classes themselves, so
// the oracle, the fixtures and the CPU baseline are anchored at the reference's nodes, not at our reading of them.
#include <zeno/zeno.h>
#include <zeno/ZenoInc.h>
#include <zeno/NumericObject.h>
#include <zeno/MeshObject.h>
#include <zeno/VDBGrid.h>
#include <omp.h>
#include <limits>
#include <openvdb/points/PointCount.h>
#include <openvdb/points/PointAdvect.h>
#include <openvdb/tree/LeafManager.h>
#include "levelset_util.h"
#include <zeno/StringObject.h>
#include <zeno/ConditionObject.h>
#include <vector>
#include <openvdb/tools/Morphology.h>
#include <openvdb/tools/MeshToVolume.h>
#include <openvdb/tools/LevelSetTracker.h>
#include <openvdb/tools/Filter.h>
#include <openvdb/tools/LevelSetSphere.h>
#include <openvdb/tools/ChangeBackground.h>
#include "FLIP_vdb.h"
#include "vdb_velocity_extrapolator.h"

namespace zeno {
inline std::map<std::string, NodeClass>& refNodeRegistry() { static std::map<std::string, NodeClass> r; return r; }
template <class T>
int defRefNodeClass(std::string const& name, Descriptor const& desc = {}) {
    refNodeRegistry()[name] = NodeClass{[] { return std::unique_ptr<INode>(new T()); }, desc};
    return 1;
}
}  // namespace zeno

#define defNodeClass defRefNodeClass
namespace refnodes {
namespace zeno { using namespace ::zeno; }
#include "nosys/P2G.cpp"
#include "nosys/SheetG2PAdvector.cpp"
#include "nosys/EvalFaceWeight.cpp"
#include "nosys/FixLiquidSDF.cpp"
#include "nosys/FieldAddVector.cpp"
#include "nosys/CFL.cpp"
#include "nosys/SolvePoissonPressureEqn.cpp"
#include "nosys/SubtractPressureGradient.cpp"
#include "nosys/KillParticles.cpp"
#include "nosys/ParticleAddGravity.cpp"     // ParticleAddDV
#include "nosys/G2P_Advector.cpp"           // the plain advector
#include "nosys/FLIP_Reseed.cpp"            // FluidReseed
#include "nosys/ParticleEmitter.cpp"        // ParticleEmitter
#include "nosys/Update_Solid_SDF.cpp"       // FLIPApplyBoundary
#include "VDBRenormalize.cpp"               // projects/zenvdb: VDBRenormalizeSDF
}  // namespace refnodes
#undef defNodeClass

#include <cstring>
#include "../../include/flipb200.h"   // grid ids only

namespace { std::string g_err; }
#include "shims/seeded_random.h"            // flipref::seed(): what FLIP_vdb.cpp's std::random_device returns in this build
#undef random_device
#define NH_SET_SEED(s) (flipref::seed() = (s))
#define NH_FN(name) rn_##name
#define NH_REGISTRY ::zeno::refNodeRegistry()
#include "node_harness.inc"
