// TEST / BENCH INFRASTRUCTURE -- the NODES of the Zeno-side drop-in (zeno_b200/plugin/flipb200_nodes.cpp) linked against the
// PRODUCT, libflipb200.so: the same harness as plugin_nodes_test.cpp (real OpenVDB objects in a RefWorld, sockets wired like
// the packaged graph, apply() of the node class registered under the reference's node name), but the flipb200_* calls the
// nodes make go to the CUDA library. tests/test_plugin_gpu.py runs one substep through these node classes on the GPU and
// compares with the oracle; bench.py times the substep through them (`e2e_nodes`).
// Built by oracle/ref/build_ref.sh into oracle/_ref/libflipplugin_gpu.so with hidden visibility: the inline functions and the
// node registry of this copy must not be merged with those of the oracle-backed copy inside libflipref.so.
#include <tbb/parallel_for.h>
#define flipb200 flipb200_gpu
#include "../../zeno_b200/plugin/flipb200_nodes.cpp"
#undef flipb200

namespace { std::string g_err; }

#define NH_HAS_PRIMITIVE 1
#define NH_SET_SEED(s) (zeno::seed_base() = (s), zeno::seed_fixed() = true)
#define NH_FN(name) pg_##name
#define NH_REGISTRY ::zeno::nodeRegistry()
#pragma GCC visibility push(default)
#include "node_harness.inc"
#pragma GCC visibility pop
