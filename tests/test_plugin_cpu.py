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
             "PushOutLiquidSDF": "FixLiquidSDF.cpp", "FieldAddVector": "FieldAddVector.cpp", "CFL_dt": "CFL.cpp"}
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
}


def descriptors(text):
    """name -> (inputs, outputs, params) from `defNodeClass<T>("name", { {inputs}, {outputs}, {params}, {category} })`"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r'defNodeClass<\w+>\(\s*"(\w+)"\s*,', text):
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
