// TEST INFRASTRUCTURE ONLY -- runs the NODES of the Zeno-side drop-in (zeno_b200/plugin/flipb200_nodes.cpp) on the CPU.
//
// The plugin source is compiled here unchanged against (a) a minimal stand-in for the Zeno node runtime
// (oracle/ref/shims/zeno_nodes: INode sockets/params, defNodeClass, NumericObject, VDBGridWrapper) and (b) a C ABI whose
// flipb200_* entry points are renamed to ob_* and implemented below by the CPU oracle (liboracle.so, dlopen'ed; the same
// flat leaf layout). A "pn world" holds REAL OpenVDB objects (a RefWorld of ref_driver.cpp); every pn_<node> call wires those
// objects to the sockets of the plugin's node class, exactly as the packaged graph does (projects/tools/FLIPtools/stub.cpp),
// and runs its apply(). tests/test_plugin_cpu.py drives the reference nodes (ref_*) and the plugin nodes (pn_*) from the
// same state and compares: together with the GPU parity tests (CUDA library == oracle through the same ABI) this closes
// the drop-in chain  reference node == plugin node o ABI(oracle) == plugin node o ABI(CUDA).
#define flipb200_last_error ob_last_error
#define flipb200_world_create ob_world_create
#define flipb200_world_destroy ob_world_destroy
#define flipb200_grid_upload ob_grid_upload
#define flipb200_grid_leaf_count ob_grid_leaf_count
#define flipb200_grid_download ob_grid_download
#define flipb200_particles_upload ob_particles_upload
#define flipb200_particles_info ob_particles_info
#define flipb200_particles_download ob_particles_download
#define flipb200_p2g ob_p2g
#define flipb200_g2p_advect_sheetty ob_g2p_advect_sheetty
#define flipb200_face_weights ob_face_weights
#define flipb200_pushout_sdf ob_pushout_sdf
#define flipb200_add_vector ob_add_vector
#define flipb200_cfl ob_cfl
#define flipb200_solve_ppe ob_solve_ppe
#define flipb200_subtract_grad ob_subtract_grad
#define flipb200_kill_particles_in_sdf ob_kill_particles_in_sdf
#define flipb200_particles_add_dv ob_particles_add_dv
#define flipb200_fluid_reseed ob_fluid_reseed
#define flipb200_emit_liquid ob_emit_liquid
#define flipb200_apply_boundary ob_apply_boundary
#define flipb200_particles_to_points ob_particles_to_points
#define flipb200_set_surface_tension ob_set_surface_tension
#define flipb200_g2p_advect ob_g2p_advect
#define flipb200_renormalize_sdf ob_renormalize_sdf
#define flipb200_erode_sdf ob_erode_sdf
#define flipb200_smooth_sdf ob_smooth_sdf
#define flipb200_host_alloc ob_host_alloc
#define flipb200_host_free ob_host_free
#define flipb200_grid_download_begin ob_grid_download_begin
#define flipb200_particles_download_begin ob_particles_download_begin
#define flipb200_download_wait ob_download_wait
#include <tbb/parallel_for.h>
#include "../../zeno_b200/plugin/flipb200_nodes.cpp"

#include <dlfcn.h>

// ---------------------------------------------------------------- the oracle behind the C ABI
namespace {
struct OracleApi {
    void* handle = nullptr;
    void* (*world_create)(float) = nullptr;
    void (*world_destroy)(void*) = nullptr;
    int (*grid_set)(void*, int, int, const int32_t*, const uint64_t*, const float*, const float*) = nullptr;
    int (*grid_leaf_count)(void*, int) = nullptr;
    int (*grid_get)(void*, int, int32_t*, uint64_t*, float*, float*) = nullptr;
    int (*particles_set)(void*, int, const int32_t*, const uint32_t*, uint64_t, const uint16_t*, const uint16_t*) = nullptr;
    int (*particles_info)(void*, int*, uint64_t*) = nullptr;
    int (*particles_get)(void*, int32_t*, uint32_t*, uint16_t*, uint16_t*) = nullptr;
    int (*p2g)(void*, float, int) = nullptr;
    int (*g2p)(void*, float, float, int, int, float, float, int) = nullptr;
    int (*face_weights)(void*) = nullptr;
    int (*pushout)(void*, float) = nullptr;
    int (*add_vector)(void*, float, float, float) = nullptr;
    float (*cfl)(void*) = nullptr;
    int (*solve)(void*, float, float, int*, float*, int*) = nullptr;
    int (*subtract)(void*, float, float, int) = nullptr;
    int (*kill)(void*, int, int) = nullptr;
    int (*add_dv)(void*, float, float, float) = nullptr;
    int (*reseed)(void*, uint32_t, const uint64_t*, uint64_t*) = nullptr;
    int (*boundary)(void*, int, int) = nullptr;
    int (*to_points)(void*, float*, float*) = nullptr;
    int (*tension)(void*, float, float) = nullptr;
    int (*emit)(void*, int, float, float, float, uint32_t, const uint64_t*, uint64_t*) = nullptr;
    int (*g2p_plain)(void*, float, float, int, float) = nullptr;
    int (*renorm)(void*, int, int, int) = nullptr;
    int (*erode)(void*, int, float) = nullptr;
    int (*smooth)(void*, int, int, int) = nullptr;
};
OracleApi g_orc;
std::string g_err;
template <typename F> void bind(F& f, const char* name) {
    f = reinterpret_cast<F>(dlsym(g_orc.handle, name));
    if (!f) throw std::runtime_error(std::string("liboracle has no ") + name);
}
}  // namespace

struct flipb200_world { void* orc = nullptr; };

extern "C" {
const char* ob_last_error(void) { return g_err.c_str(); }
int ob_world_create(int, float dx, flipb200_world** out) {
    if (!g_orc.handle) { g_err = "pn_backend was not called"; return FLIPB200_ERR_STATE; }
    *out = new flipb200_world{g_orc.world_create(dx)};
    return 0;
}
int ob_world_destroy(flipb200_world* w) { if (w) { g_orc.world_destroy(w->orc); delete w; } return 0; }
int ob_grid_upload(flipb200_world* w, int grid, int n, const int32_t* o, const uint64_t* m, const float* v, int layout, const float* bg) {
    const int nch = grid <= FLIPB200_FACE_WEIGHT ? 3 : 1;
    if (nch == 3 && layout == FLIPB200_AOS) {   // [leaf][512][3] -> the oracle's [leaf][3][512]
        std::vector<float> soa(size_t(n) * 1536);
        for (size_t l = 0; l < size_t(n); l++)
            for (int i = 0; i < 512; i++)
                for (int c = 0; c < 3; c++) soa[l * 1536 + size_t(c) * 512 + i] = v[l * 1536 + size_t(i) * 3 + c];
        return g_orc.grid_set(w->orc, grid, n, o, m, soa.data(), bg);
    }
    return g_orc.grid_set(w->orc, grid, n, o, m, v, bg);
}
int ob_grid_leaf_count(flipb200_world* w, int grid, int* n) { *n = g_orc.grid_leaf_count(w->orc, grid); return *n < 0 ? FLIPB200_ERR_ARG : 0; }
int ob_grid_download(flipb200_world* w, int grid, int32_t* o, uint64_t* m, float* v, int layout, float* bg) {
    const int nch = grid <= FLIPB200_FACE_WEIGHT ? 3 : 1;
    if (nch == 3 && layout == FLIPB200_AOS) {
        const int n = g_orc.grid_leaf_count(w->orc, grid);
        std::vector<float> soa(size_t(n) * 1536);
        int rc = g_orc.grid_get(w->orc, grid, o, m, soa.data(), bg);
        for (size_t l = 0; l < size_t(n); l++)
            for (int i = 0; i < 512; i++)
                for (int c = 0; c < 3; c++) v[l * 1536 + size_t(i) * 3 + c] = soa[l * 1536 + size_t(c) * 512 + i];
        return rc;
    }
    return g_orc.grid_get(w->orc, grid, o, m, v, bg);
}
int ob_host_alloc(size_t bytes, void** out) { *out = std::malloc(bytes); return *out ? 0 : FLIPB200_ERR_CUDA; }
int ob_host_free(void* p) { std::free(p); return 0; }
int ob_download_wait(flipb200_world*) { return 0; }
int ob_grid_download_begin(flipb200_world* w, int grid, int cap, int32_t* o, uint64_t* m, float* v, int layout, float* bg, int* n) {
    *n = g_orc.grid_leaf_count(w->orc, grid);
    if (*n < 0 || *n > cap) { g_err = "grid_download_begin: capacity"; return FLIPB200_ERR_ARG; }
    return ob_grid_download(w, grid, o, m, v, layout, bg);
}
int ob_particles_upload(flipb200_world* w, int nl, const int32_t* o, const uint32_t* ve, uint64_t np, const uint16_t* P, const uint16_t* V) {
    return g_orc.particles_set(w->orc, nl, o, ve, np, P, V);
}
int ob_particles_info(flipb200_world* w, int* nl, uint64_t* np) { return g_orc.particles_info(w->orc, nl, np); }
int ob_particles_download(flipb200_world* w, int32_t* o, uint32_t* ve, uint16_t* P, uint16_t* V) { return g_orc.particles_get(w->orc, o, ve, P, V); }
int ob_particles_download_begin(flipb200_world* w, int capLeaves, uint64_t capParticles, int32_t* o, uint32_t* ve, uint16_t* P, uint16_t* V, int* nl, uint64_t* np) {
    int rc = g_orc.particles_info(w->orc, nl, np);
    if (rc) return rc;
    if (*nl > capLeaves || *np > capParticles) { g_err = "particles_download_begin: capacity"; return FLIPB200_ERR_ARG; }
    return g_orc.particles_get(w->orc, o, ve, P, V);
}
int ob_p2g(flipb200_world* w, float dx, int n) { return g_orc.p2g(w->orc, dx, n); }
int ob_g2p_advect_sheetty(flipb200_world* w, float dt, float dx, int ss, int rk, float pmin, float pmax, int flags) {
    return g_orc.g2p(w->orc, dt, dx, ss, rk, pmin, pmax, flags);
}
int ob_face_weights(flipb200_world* w) { return g_orc.face_weights(w->orc); }
int ob_pushout_sdf(flipb200_world* w, float dx) { return g_orc.pushout(w->orc, dx); }
int ob_add_vector(flipb200_world* w, float x, float y, float z) { return g_orc.add_vector(w->orc, x, y, z); }
int ob_cfl(flipb200_world* w, float* dt) { *dt = g_orc.cfl(w->orc); return 0; }
int ob_solve_ppe(flipb200_world* w, float dt, float dx, int* it, float* res, int* st) { return g_orc.solve(w->orc, dt, dx, it, res, st); }
int ob_subtract_grad(flipb200_world* w, float dt, float dx, int n) { return g_orc.subtract(w->orc, dt, dx, n); }
int ob_kill_particles_in_sdf(flipb200_world* w, int grid, int keep) { return g_orc.kill(w->orc, grid, keep); }
int ob_particles_add_dv(flipb200_world* w, float x, float y, float z) { return g_orc.add_dv(w->orc, x, y, z); }
int ob_fluid_reseed(flipb200_world* w, uint32_t seed) { return g_orc.reseed(w->orc, seed, nullptr, nullptr); }
int ob_set_surface_tension(flipb200_world* w, float density, float coef) { return g_orc.tension(w->orc, density, coef); }
int ob_particles_to_points(flipb200_world* w, float* pos, float* vel) { return g_orc.to_points(w->orc, pos, vel); }
int ob_apply_boundary(flipb200_world* w, int grid, int vc) { return g_orc.boundary(w->orc, grid, vc); }
int ob_emit_liquid(flipb200_world* w, int grid, float vx, float vy, float vz, uint32_t seed) { return g_orc.emit(w->orc, grid, vx, vy, vz, seed, nullptr, nullptr); }
int ob_g2p_advect(flipb200_world* w, float dt, float dx, int rk, float s) { return g_orc.g2p_plain(w->orc, dt, dx, rk, s); }
int ob_renormalize_sdf(flipb200_world* w, int grid, int it, int dil) { return g_orc.renorm(w->orc, grid, it, dil); }
int ob_erode_sdf(flipb200_world* w, int grid, float d) { return g_orc.erode(w->orc, grid, d); }
int ob_smooth_sdf(flipb200_world* w, int grid, int wd, int it) { return g_orc.smooth(w->orc, grid, wd, it); }
}  // extern "C"

#define NH_HAS_PRIMITIVE 1
#define NH_SET_SEED(s) (zeno::seed_base() = (s), zeno::seed_fixed() = true)
#define NH_FN(name) pn_##name
#define NH_REGISTRY ::zeno::nodeRegistry()
#include "node_harness.inc"

extern "C" {
int pn_backend(const char* liboracle) {
    return guarded([&] {
        if (g_orc.handle) return;
        g_orc.handle = dlopen(liboracle, RTLD_NOW | RTLD_LOCAL);
        if (!g_orc.handle) throw std::runtime_error(std::string("cannot load ") + liboracle + ": " + dlerror());
        bind(g_orc.world_create, "orc_world_create"); bind(g_orc.world_destroy, "orc_world_destroy");
        bind(g_orc.grid_set, "orc_grid_set"); bind(g_orc.grid_leaf_count, "orc_grid_leaf_count"); bind(g_orc.grid_get, "orc_grid_get");
        bind(g_orc.particles_set, "orc_particles_set"); bind(g_orc.particles_info, "orc_particles_info"); bind(g_orc.particles_get, "orc_particles_get");
        bind(g_orc.p2g, "orc_p2g"); bind(g_orc.g2p, "orc_g2p_advect_sheetty"); bind(g_orc.face_weights, "orc_face_weights");
        bind(g_orc.pushout, "orc_pushout_sdf"); bind(g_orc.add_vector, "orc_add_vector"); bind(g_orc.cfl, "orc_cfl");
        bind(g_orc.solve, "orc_solve_ppe"); bind(g_orc.subtract, "orc_subtract_grad"); bind(g_orc.kill, "orc_kill_particles"); bind(g_orc.add_dv, "orc_particles_add_dv"); bind(g_orc.reseed, "orc_fluid_reseed"); bind(g_orc.emit, "orc_emit_liquid"); bind(g_orc.boundary, "orc_apply_boundary"); bind(g_orc.to_points, "orc_particles_to_points"); bind(g_orc.tension, "orc_set_surface_tension"); bind(g_orc.g2p_plain, "orc_g2p_advect"); bind(g_orc.renorm, "orc_renormalize_sdf"); bind(g_orc.erode, "orc_erode_sdf"); bind(g_orc.smooth, "orc_smooth_sdf");
    });
}
}  // extern "C"
