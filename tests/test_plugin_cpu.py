"""The Zeno-side drop-in (zeno_b200/plugin/flipb200_nodes.cpp) must register the reference's node names with the
reference's sockets and params. Where /root/reference exists the descriptors are compared with the
reference sources literally; elsewhere against the table recorded here from the same sources."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "zeno_b200", "plugin", "flipb200_nodes.cpp")
REF = "/root/reference/projects/FastFLIP/nosys"
REF_FILES = {"FLIP_P2G": "P2G.cpp", "G2PAdvectorSheetty": "SheetG2PAdvector.cpp", "AssembleSolvePPE": "SolvePoissonPressureEqn.cpp",
             "SubtractPressureGradient": "SubtractPressureGradient.cpp", "CutCellWeight": "EvalFaceWeight.cpp",
             "PushOutLiquidSDF": "FixLiquidSDF.cpp", "FieldAddVector": "FieldAddVector.cpp", "CFL_dt": "CFL.cpp",
             "KillParticlesInSDF": "KillParticles.cpp", "ParticleAddDV": "ParticleAddGravity.cpp", "FluidReseed": "FLIP_Reseed.cpp", "ParticleEmitter": "ParticleEmitter.cpp", "FLIPApplyBoundary": "Update_Solid_SDF.cpp", "VDBPointsToPrimitive": "../../zenvdb/GetVDBPoints.cpp",
             "G2P_Advector": "G2P_Advector.cpp", "VDBRenormalizeSDF": "../../zenvdb/VDBRenormalize.cpp",
             "VDBErodeSDF": "../../zenvdb/VDBRenormalize.cpp", "VDBSmoothSDF": "../../zenvdb/VDBRenormalize.cpp"}
# (inputs, outputs, params) by name only, recorded from the reference files above
EXPECTED = {
    "FLIP_P2G": (["Dx", "Particles", "Velocity", "PostP2GVelocity", "LiquidSDF"], [], ["dx", "VelExtraLayer"]),
    "G2PAdvectorSheetty": (["dt", "Dx", "pic_min", "pic_max", "Particles", "Velocity", "ViscousVelocity", "LiquidSDF", "PostAdvVelocity",
                            "SolidSDF", "SolidVelocity"], [], ["dx", "RK_ORDER", "pic_smoothness", "surface_size"]),
    "AssembleSolvePPE": (["dt", "Dx", "Density", "SurfaceTension", "LiquidSDF", "Divergence", "Pressure", "CellFWeight", "Velocity",
                          "SolidVelocity", "Curvature"], [], ["dx"]),
    "SubtractPressureGradient": (["dt", "Dx", "Density", "SurfaceTension", "LiquidSDF", "SolidSDF", "Pressure", "CellFWeight", "Velocity",
                                  "SolidVelocity", "Curvature"], [], ["dx", "VelExtraLayer"]),
    "CutCellWeight": (["LiquidSDF", "SolidSDF", "FaceWeight"], [], []),
    "PushOutLiquidSDF": (["Dx", "LiquidSDF", "SolidSDF"], [], ["dx"]),
    "FieldAddVector": (["invec3", "Velocity", "FieldWeight"], [], []),
    "CFL_dt": (["Velocity", "Dx"], ["cfl_dt"], ["dx"]),
    "KillParticlesInSDF": (["Particles", "KillerSDF"], ["Particles"], ["OpType"]),
    "ParticleAddDV": (["Particles", "dv"], [], ["channel", "vx", "vy", "vz"]),
    "FluidReseed": (["Particles", "LiquidSDF", "FluidVel"], [], []),
    "VDBPointsToPrimitive": (["grid"], ["prim"], []),
    "FLIPApplyBoundary": (["Particles", "DynaSolid_SDF", "StatSolid_SDF"], [], []),
    "ParticleEmitter": (["Particles", "ShapeSDF", "VelocityVolume", "VelocityInit", "LiquidSDF"], ["Particles"], ["vx", "vy", "vz"]),
    "G2P_Advector": (["dt", "Dx", "Particles", "Velocity", "PostAdvVelocity", "SolidSDF", "SolidVelocity"], [], ["dx", "RK_ORDER", "pic_smoothness"]),
    "VDBRenormalizeSDF": (["inoutSDF"], ["inoutSDF"], ["method", "iterations", "dilateIters"]),
    "VDBErodeSDF": (["inoutSDF", "depth"], ["inoutSDF"], []),
    "VDBSmoothSDF": (["inoutSDF"], ["inoutSDF"], ["width", "iterations", "DEPRECATED"]),
}


def descriptors(text):
    """name -> (inputs, outputs, params) from `defNodeClass<T>("name", { {inputs}, {outputs}, {params}, {category} })`"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    # both registration forms: defNodeClass<T>("name", {...}) and ZENDEFNODE(name, {...}) (zeno/core/defNode.h:28-34)
    for m in list(re.finditer(r'defNodeClass<\w+>\(\s*"(\w+)"\s*,', text)) + list(re.finditer(r'ZENDEFNODE\(\s*(\w+)\s*,', text)):
        i = text.index("{", m.end())
        depth, j = 0, i
        while True:
            depth += text[j] == "{"
            depth -= text[j] == "}"
            if depth == 0:
                break
            j += 1
        body = text[i + 1:j]
        groups, depth, start = [], 0, None
        for k, ch in enumerate(body):
            if ch == "{":
                if depth == 0:
                    start = k
                depth += 1
            elif ch == "}":
                depth -= 1
                if depth == 0:
                    groups.append(body[start + 1:k])
        def names(g, param=False):
            res = []
            for item in re.finditer(r'\{([^{}]*)\}|"(\w+)"', g):
                if item.group(1) is not None:
                    q = re.findall(r'"([^"]*)"', item.group(1))
                    res.append(q[1])
                else:
                    res.append(item.group(2))
            return res
        out[m.group(1)] = (names(groups[0]), names(groups[1]), names(groups[2]))
    return out


def test_plugin_registers_reference_descriptors():
    mine = descriptors(open(PLUGIN).read())
    assert set(mine) == set(EXPECTED), sorted(mine)
    for name, exp in EXPECTED.items():
        assert mine[name] == tuple(exp) or list(map(list, mine[name])) == list(map(list, exp)), (name, mine[name], exp)
    if os.path.isdir(REF):
        for name, fn in REF_FILES.items():
            ref = descriptors(open(os.path.join(REF, fn)).read())
            assert ref[name] == mine[name], (name, ref[name], mine[name])


def test_plugin_calls_only_the_c_abi():
    src = open(PLUGIN).read()
    code = re.sub(r"//[^\n]*", "", src)  # comments may mention what the file does not use
    assert "FLIP_vdb::" not in code and "cuda" not in code.lower() and "torch" not in code.lower()
    hdr = open(os.path.join(ROOT, "include", "flipb200.h")).read()
    for call in set(re.findall(r"\b(flipb200_\w+)\(", src)):
        assert re.search(r"\b" + call + r"\(", hdr), f"{call} is not declared in include/flipb200.h"


def test_plugin_marshalling_roundtrips_real_reference_objects():
    """The plugin's OpenVDB <-> flat-array marshalling (upload / download / upload_particles / download_particles), compiled
    unchanged into oracle/_ref with a loopback C ABI that hands the leaves back in reverse order, must return every grid and the
    particle tree of a world the REAL reference nodes produced -- same leaves, masks, values, per-voxel offsets and raw codec words
    (oracle/ref/ref_driver.cpp: ref_plugin_roundtrip)."""
    import ctypes as C

    import pytest

    from oracle import pyoracle
    from zeno_b200 import scenes
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref was not built (no /root/reference here)")
    lib = pyoracle.load_ref()
    if not hasattr(lib, "ref_plugin_roundtrip"):
        pytest.skip("oracle/_ref predates the marshalling harness")
    N = 32
    pos, vel, dx = scenes.dam_break_points(N, seed=4, random_velocity=True)
    w = pyoracle.RefWorld(dx)
    w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
    w.PrimToVDBPointDataGrid(pos, vel * 0.2)
    w.FLIP_P2G(dx, 3)
    dt = 0.01
    for _ in range(2):
        w.substep(dt, dx, 4, 3, 0.03, 0.05, (0.0, -9.8, 0.0), 3, True)
    msg = C.create_string_buffer(256)
    rc = lib.ref_plugin_roundtrip(w.h, msg, C.c_int(256))
    assert rc == 0, msg.value.decode()


def test_plugin_nodes_reproduce_the_reference_nodes():
    """The drop-in's NODE classes (same names, sockets, params), instantiated through a minimal stand-in of the Zeno node
    runtime and wired to real OpenVDB objects the way the packaged graph wires them, with the CPU oracle behind the C ABI they
    call (oracle/ref/plugin_nodes_test.cpp). Stage by stage from the reference's state:
      * against the oracle driven directly (no plugin, no OpenVDB): bit-identical -> sockets, params, upload/download sets and
        the marshalling add nothing;
      * against the REAL reference nodes: within the oracle's own pinned tolerances (tests/test_ref_pin_cpu.py).
    With the GPU parity tests (CUDA library == oracle through the same ABI) this closes the drop-in chain."""
    import numpy as np
    import pytest

    from oracle import pyoracle
    from tests import util
    from zeno_b200 import scenes
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref was not built (no /root/reference here)")
    if not hasattr(pyoracle.load_ref(), "pn_backend"):
        pytest.skip("oracle/_ref predates the node harness")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefWorld
    pyoracle.ref_set_threads(0)
    N, dt = 40, 0.006
    pos, vel, dx = scenes.dam_break_points(N, seed=11, random_velocity=True)
    vel *= np.float32(0.25)
    solid = scenes.box_solid_sdf(N, dx)
    pw, ow, rw = PluginWorld(dx), OracleWorld(dx), RefWorld(dx)
    for w in (pw, ow, rw):
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(pos, vel)
    for name, grids in util.REF_STAGES:
        util.sync_state(ow, rw)
        util.sync_state(pw, rw)
        res = [util.run_ref_stage(w, name, dx, dt) for w in (pw, ow, rw)]
        for g in grids:
            util.compare_grids(pw.get_grid(g), ow.get_grid(g), f"plugin node vs oracle: {name}.{g}", tol=0.0, check_inactive=False)
            util.compare_grids(pw.get_grid(g), rw.get_grid(g), f"plugin node vs reference node: {name}.{g}", tol=util.REF_TOL[name], check_inactive=False)
        if name == "g2p":
            a, b, c = (scenes.canonical_particles(w.get_particles()) for w in (pw, ow, rw))
            assert np.array_equal(a, b), "plugin G2PAdvectorSheetty differs from the oracle driven directly"
            m = util.particle_code_report(a, c)
            assert a.shape == c.shape and m["same_voxel"] >= 0.999 and m["P_within_1lsb"] >= 0.999 and m["v_within_1ulp"] >= 0.999, m
        if name == "grad":
            cp, co, cr = pw.CFL_dt(), ow.CFL_dt(), rw.CFL_dt()
            assert cp == co and abs(cp - cr) <= 1e-5 * cr, (cp, co, cr)


def test_plugin_resident_mode():
    """FLIPB200_RESIDENT=1 (grids stay on the device between accelerated nodes; an object is re-uploaded only when its
    tree pointer or leaf count changed) must not change any result: two free-running substeps through the plugin nodes give
    the same digest with and without it, and the same as the oracle driven directly."""
    import subprocess
    import sys

    import pytest

    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "pn_backend"):
        pytest.skip("oracle/_ref with the node harness is not available here")
    worker = os.path.join(ROOT, "tests", "plugin_resident_worker.py")

    def digest(args, env_extra):
        env = dict(os.environ)
        env.pop("FLIPB200_RESIDENT", None)
        env.update(env_extra)
        r = subprocess.run([sys.executable, worker] + args, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return [ln.split()[1] for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][0]

    plain = digest([], {})
    resident = digest([], {"FLIPB200_RESIDENT": "1"})
    direct = digest(["oracle"], {})
    assert plain == direct, "plugin nodes (upload-everything mode) differ from the oracle driven directly"
    assert resident == plain, "FLIPB200_RESIDENT=1 changes the results"


def test_points_to_primitive_node_and_oracle():
    """VDBPointsToPrimitive (projects/zenvdb/GetVDBPoints.cpp:163-177): world position = float((double(P) + double(voxel)) * dx), decoded
    velocity. The oracle's restatement against the same arithmetic written with numpy on the stored codes, and the drop-in's node
    class (oracle behind the ABI, real OpenVDB particle grid in, PrimitiveObject out) against the oracle."""
    import numpy as np
    from oracle import pyoracle
    from oracle.pyoracle import OracleWorld, PluginWorld
    from zeno_b200 import scenes
    pos, vel, dx = scenes.dam_break_points(32, seed=11, random_velocity=True)
    ow = OracleWorld(dx)
    ow.PrimToVDBPointDataGrid(pos, vel)
    p = ow.get_particles()
    counts = np.diff(np.concatenate([np.zeros((p["voxel_end"].shape[0], 1), np.uint32), p["voxel_end"]], axis=1).astype(np.int64), axis=1)   # [leaf, 512]
    off = np.arange(512)
    loc = np.stack([off >> 6, (off >> 3) & 7, off & 7], axis=1)
    vox = (p["origins"][:, None, :].astype(np.int64) + loc[None, :, :]).reshape(-1, 3)
    ijk = np.repeat(vox, counts.reshape(-1), axis=0)
    Pdec = p["P"].astype(np.float32) / np.float32(65535.0) - np.float32(0.5)
    want_pos = ((Pdec.astype(np.float64) + ijk.astype(np.float64)) * np.float64(np.float32(dx))).astype(np.float32)
    want_vel = p["v"].view(np.float16).astype(np.float32)
    got_pos, got_vel = ow.VDBPointsToPrimitive()
    assert np.array_equal(got_pos, want_pos) and np.array_equal(got_vel, want_vel)
    if pyoracle.ref_available() and hasattr(pyoracle.load_ref(), "pn_particles_to_points"):
        pw = PluginWorld(dx)
        pw.PrimToVDBPointDataGrid(pos, vel)
        node_pos, node_vel = pw.VDBPointsToPrimitive()
        a = np.concatenate([node_pos, node_vel], axis=1)
        b = np.concatenate([got_pos, got_vel], axis=1)
        a = a[np.lexsort(tuple(a[:, k] for k in range(5, -1, -1)))]
        b = b[np.lexsort(tuple(b[:, k] for k in range(5, -1, -1)))]
        assert np.array_equal(a, b), "plugin node primitive differs from the oracle"
