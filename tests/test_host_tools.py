"""Host C++ layer (scene loader, frame writer, option parsing) — CPU tests.

The loader must be bit-identical to the reference loader (initial positions define every parity
comparison), the frame files byte-identical to the reference io/ module's.
"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from oracle import oracle_api

ROOT = Path(__file__).resolve().parent.parent
TOOLS = ROOT / "fluidsimulator_b200" / "bin" / "pbf_host_tools"
APP = ROOT / "fluidsimulator_b200" / "bin" / "fluidsim_b200"


def fnv1a(a: np.ndarray) -> str:
    h = 1469598103934665603
    data = np.ascontiguousarray(a).tobytes()
    # vectorised FNV is awkward; arrays here are <= a few MB, a plain loop over bytes in chunks is fine
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def tool(*args):
    out = subprocess.run([str(TOOLS), *map(str, args)], capture_output=True, text=True, check=True).stdout
    return dict(line.split("=", 1) for line in out.strip().splitlines())


def f32hex(x) -> str:
    return f"{np.float32(x).view(np.uint32):08x}"


@pytest.mark.parametrize("name", ["fluid_large", "fluid_double_side"])
def test_cpp_loader_matches_python_emitter(built, tmp_path, name):
    path = scenes.SCENES[name].write_json(tmp_path / f"{name}.json")
    got = tool("scene", path)
    params, planes, state = scenes.load_scene(scenes.SCENES[name])
    assert int(got["count"]) == len(state[0])
    for key, arr in zip(["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"], state):
        assert got[key] == fnv1a(arr), key
    assert int(got["planes"]) == planes.shape[0]
    for k, key in enumerate(["plane_nx", "plane_ny", "plane_nz", "plane_d"]):
        assert got[key] == fnv1a(planes[:, k].copy())
    assert got["h"] == f32hex(params.h) and got["particle_radius"] == f32hex(params.particle_radius)
    assert got["density"] == f32hex(params.density) and got["epsilon"] == f32hex(params.epsilon)
    assert got["scorr_k"] == f32hex(params.scorr_k) and got["visc_c"] == f32hex(params.visc_c)
    assert int(got["scorr_n"]) == params.scorr_n


def test_cpp_loader_sphere_and_velocity(built, tmp_path):
    obj = scenes.small_block(6).to_json_obj()
    obj["fluid"]["shape"].append({"type": "sphere", "origin": [1.0, 1.1, 0.9], "radius": 0.21, "velocity": [0.5, -1, 0]})
    import json
    path = tmp_path / "s.json"
    path.write_text(json.dumps(obj))
    got = tool("scene", path)
    params, planes, state = scenes.load_scene(obj)
    assert int(got["count"]) == len(state[0])
    for key, arr in zip(["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"], state):
        assert got[key] == fnv1a(arr), key


@pytest.mark.skipif(not oracle_api.available("reference"), reason="needs oracle/_ref")
def test_cpp_test_scene_matches_reference(built):
    """The built-in scene used without --scene (reference init.cpp:119-156)."""
    from oracle.oracle_api import Oracle
    got = tool("test-scene")
    ref = Oracle("reference")
    ref.init_test_scene()
    st = ref.get_state()
    assert int(got["count"]) == len(st[0]) == 46875
    for key, arr in zip(["pos_x", "pos_y", "pos_z"], st):
        assert got[key] == fnv1a(arr)
    p = ref.get_params()
    assert got["h"] == f32hex(p.h) and got["particle_mass"] == f32hex(p.particle_mass)
    assert got["plane_d"] == fnv1a(ref.get_planes()[:, 3].copy())


def test_frame_bytes(built, tmp_path):
    """frame_000000.vtp / series.pvd bytes: setprecision(9) << fixed, three lines per particle."""
    sc = scenes.small_block(4)
    path = sc.write_json(tmp_path / "s.json")
    out = tmp_path / "out"
    subprocess.run([str(TOOLS), "frame", str(path), str(out)], check=True)
    _, _, st = scenes.load_scene(sc)
    n = len(st[0])
    lines = ['<?xml version="1.0"?>', '<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian">',
             "  <PolyData>",
             f'    <Piece NumberOfPoints="{n}" NumberOfVerts="{n}" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="0">',
             "      <Points>", '        <DataArray type="Float32" NumberOfComponents="3" format="ascii">']
    lines += [f"          {float(x):.9f} {float(y):.9f} {float(z):.9f}" for x, y, z in zip(*st[:3])]
    lines += ["        </DataArray>", "      </Points>", "      <Verts>",
              '        <DataArray type="Int32" Name="connectivity" format="ascii">']
    lines += [f"          {i}" for i in range(n)]
    lines += ["        </DataArray>", '        <DataArray type="Int32" Name="offsets" format="ascii">']
    lines += [f"          {i + 1}" for i in range(n)]
    lines += ["        </DataArray>", "      </Verts>", "    </Piece>", "  </PolyData>", "</VTKFile>", ""]
    assert (out / "frame_000000.vtp").read_text() == "\n".join(lines)
    pvd = (out / "series.pvd").read_text()
    assert f'<DataSet timestep="{float(np.float32(1.0 / 120.0)):.9f}" group="" part="0" file="frame_000000.vtp"/>' in pvd


@pytest.mark.skipif(not oracle_api.DROPIN_BIN.exists(), reason="needs oracle/_ref/fluidsim_dropin")
def test_frame_bytes_match_reference_writer(built, tmp_path):
    """Reference application (CPU backend) frame for a scene with zero gravity and no neighbours in range
    == our writer's frame of the initial positions, byte for byte."""
    import json
    obj = scenes.small_block(3).to_json_obj()
    obj["external_forces"] = [0, 0, 0]
    obj["fluid"]["h"] = 0.01          # nobody within h: positions do not move
    path = tmp_path / "still.json"
    path.write_text(json.dumps(obj))
    ref_out, our_out = tmp_path / "ref", tmp_path / "ours"
    subprocess.run([str(oracle_api.DROPIN_BIN), "--backend=cpu", "--scene", str(path), "--steps", "1",
                    "--steps-per-sec", "120", "--output-dir", str(ref_out)], check=True, capture_output=True)
    subprocess.run([str(TOOLS), "frame", str(path), str(our_out)], check=True)
    assert (ref_out / "frame_000000.vtp").read_bytes() == (our_out / "frame_000000.vtp").read_bytes()
    assert (ref_out / "series.pvd").read_bytes() == (our_out / "series.pvd").read_bytes()


@pytest.mark.parametrize("args,msg", [
    (["--steps", "0"], "Steps must be >= 1."),
    (["--steps", "x"], "Invalid steps value."),
    (["--steps-per-sec", "-1"], "steps-per-sec must be > 0."),
    (["--plane-friction", "1.5"], "plane-friction must be in [0, 1]."),
    (["--plane-restitution", "-0.1"], "plane-restitution must be >= 0."),
    (["--bogus"], "Unknown option: --bogus"),
    (["--scene"], "Missing value for option: --scene"),
    (["--no-output=1"], "Option does not take a value: --no-output"),
    (["--backend", "cpu"], "Unsupported backend: cpu"),
    (["--backend", "metal"], "Unsupported backend: metal"),
    (["--devices", "0,x"], "Invalid devices value"),
    (["--devices", "0,-1"], "Invalid devices value"),
    (["--devices", "0,1", "--per-step-host"], "--devices and --per-step-host exclude each other."),
    (["--mode", "sloppy"], "mode must be strict or fast."),
    (["--solver-iterations", "-2"], "solver-iterations must be >= 0."),
])
def test_app_argument_validation(built, args, msg):
    """Same validation and messages as reference app/src/main.cpp:78-166 / cli.cpp:34-101."""
    r = subprocess.run([str(APP), *args], capture_output=True, text=True)
    assert r.returncode == 1
    assert msg in r.stderr


def test_app_help(built):
    r = subprocess.run([str(APP), "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    for opt in ["backend", "no-output", "debug-print", "steps", "steps-per-sec", "enable-scorr", "enable-xsph",
                "enable-vorticity", "plane-restitution", "plane-friction", "threads", "no-omp", "fps", "duration",
                "scene", "output-dir", "solver-iterations", "mode", "device", "devices", "per-step-host"]:
        assert f"--{opt}" in r.stdout
