"""Scene table / emitter parity: scenes.py must reproduce the reference loader."""
import ctypes as C
import json

import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from oracle import oracle_api

import golden_util as G

REF_SCENES = oracle_api.REFERENCE_ROOT / "scene"


def _pbytes(p):
    return bytes(C.string_at(C.byref(p), C.sizeof(p)))


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_emitter_matches_golden_digest(name):
    gold = G.digests()["scenes"][name]
    params, planes, state = scenes.load_scene(scenes.SCENES[name])
    assert len(state[0]) == gold["count"]
    assert [G.digest(a) for a in state[:3]] == gold["pos"]
    assert G.digest(planes) == gold["planes"]
    for k, v in gold["params"].items():
        mine = params.as_dict()[k]
        assert np.float32(mine).tobytes() == np.float32(v).tobytes() if not isinstance(v, list) else \
            [np.float32(x).tobytes() for x in mine] == [np.float32(x).tobytes() for x in v], k


def test_particle_counts():
    """SURVEY.md §6 scene sizes."""
    want = {"fluid_large": 19683, "fluid_double_side": 93312, "fluid_double_dem": 142560,
            "fluid_xlarge": 166375, "fluid_million": 1000000}
    for name, n in want.items():
        assert len(scenes.load_scene(scenes.SCENES[name])[2][0]) == n


def test_json_roundtrip(tmp_path):
    sc = scenes.SCENES["fluid_double_dem"]
    path = sc.write_json(tmp_path / "s.json")
    a = scenes.load_scene(sc)
    b = scenes.load_scene(path)
    assert _pbytes(a[0]) == _pbytes(b[0])
    assert np.array_equal(a[1], b[1])
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_sphere_and_velocity_shapes():
    obj = scenes.small_block(6).to_json_obj()
    obj["fluid"]["shape"].append({"type": "sphere", "origin": [1.0, 1.0, 1.0], "radius": 0.2, "velocity": [0, -1, 0.5]})
    params, planes, state = scenes.load_scene(obj)
    n_cube = len(scenes.load_scene(scenes.small_block(6))[2][0])
    assert len(state[0]) > n_cube
    d = np.stack([state[0][n_cube:] - 1.0, state[1][n_cube:] - 1.0, state[2][n_cube:] - 1.0])
    assert np.all((d * d).sum(0) < 0.2 * 0.2 + 1e-6)
    assert np.all(state[4][n_cube:] == np.float32(-1)) and np.all(state[5][n_cube:] == np.float32(0.5))


@pytest.mark.skipif(not (REF_SCENES.exists() and oracle_api.available("reference")), reason="needs /root/reference")
@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_reference_loader_agrees(name, tmp_path):
    """Reference loader on the reference file == reference loader on our generated JSON == scenes.py."""
    from oracle.oracle_api import Oracle
    a = Oracle("reference")
    a.load_scene(REF_SCENES / f"{name}.json")
    b = Oracle("reference")
    b.load_scene(scenes.SCENES[name].write_json(tmp_path / f"{name}.json"))
    params, planes, state = scenes.load_scene(scenes.SCENES[name])
    assert _pbytes(a.get_params()) == _pbytes(b.get_params()) == _pbytes(params)
    assert np.array_equal(a.get_planes().view(np.uint32), planes.view(np.uint32))
    for x, y, z in zip(a.get_state(), b.get_state(), state):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        assert np.array_equal(x.view(np.uint32), z.view(np.uint32))
