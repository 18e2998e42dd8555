"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle.

STRICT mode is required to be bit-identical to the reference CPU path on every integer
table AND every float array, free-running over many substeps with all flags on.
FAST mode keeps the integer tables bit-exact for one teacher-forced substep and is
gated by the tolerance BASELINE.json states: max-abs position error <= 1e-4 * h.
"""
import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_FAST, PBF_MODE_STRICT

import helpers as H

pytestmark = pytest.mark.gpu

SMALL = scenes.small_block(12)


@pytest.mark.parametrize("flags", [H.NO_FLAGS, H.STABLE_FLAGS, H.ALL_FLAGS], ids=["none", "stable", "all"])
def test_strict_bit_exact_small(built, flags):
    sol, orc, params = H.make_pair(SMALL, flags)
    for step in range(1, 13):
        sol.step(1)
        orc.step(1)
        assert H.compare_integers(sol, orc) == [], f"step {step}"
        assert H.compare_scratch_bits(sol, orc, params) == [], f"step {step}"
        assert H.compare_state_bits(sol, orc) == [], f"step {step}"
    assert np.float32(sol.time) == np.float32(orc.time)


@pytest.mark.parametrize("scene", ["fluid_large", "fluid_double_side"])
def test_strict_bit_exact_scene_all_flags(built, scene):
    sol, orc, params = H.make_pair(scenes.SCENES[scene], H.ALL_FLAGS)
    for step in range(1, 4):
        sol.step(1)
        orc.step(1)
        assert H.compare_integers(sol, orc) == [], f"step {step}"
        assert H.compare_scratch_bits(sol, orc, params) == [], f"step {step}"
        assert H.compare_state_bits(sol, orc) == [], f"step {step}"


def test_strict_free_running_batch(built):
    """One pbf_step(30) batch (graph replay, no host sync) == 30 oracle steps, bit for bit."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, debug=False)
    sol.step(30)
    orc.step(30)
    assert H.compare_state_bits(sol, orc) == []
    assert H.compare_integers(sol, orc) == []


@pytest.mark.parametrize("iterations", [2, 8])
def test_strict_iteration_sweep(built, iterations):
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.ALL_FLAGS, iterations=iterations)
    sol.step(2)
    orc.step(2)
    assert H.compare_state_bits(sol, orc) == []
    assert H.compare_scratch_bits(sol, orc, params) == []


def test_fast_mode_tolerance(built):
    """FAST mode: integer tables exact on the first substep; positions within 1e-4*h after 10
    free-running substeps without vorticity (SURVEY §8c gate G3)."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, mode=PBF_MODE_FAST)
    sol.step(1)
    orc.step(1)
    assert H.compare_integers(sol, orc) == []
    sol.step(9)
    orc.step(9)
    err = H.max_abs_pos_err_in_h(sol, orc, params.h)
    assert err <= 1e-4, f"max-abs position error {err:.3e} h"
    assert H.max_abs_vel_err(sol, orc) <= 1e-4 * params.h / params.dt


def test_fast_mode_vorticity_one_step(built):
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.ALL_FLAGS, mode=PBF_MODE_FAST)
    sol.step(1)
    orc.step(1)
    assert H.compare_integers(sol, orc) == []
    assert H.max_abs_pos_err_in_h(sol, orc, params.h) <= 1e-4


def test_step_host_contract(built):
    """pbf_step_host == the reference cuda_step contract: host arrays in, host arrays out."""
    sol, orc, params = H.make_pair(SMALL, H.STABLE_FLAGS)
    st = [a.copy() for a in orc.get_state()]
    for _ in range(3):
        sol.step_host(st, 1)
        orc.step(1)
    for a, b in zip(st, orc.get_state()):
        assert H.bit_equal(a, b)


def test_empty_state_advances_time(built):
    from fluidsimulator_b200.capi import Solver
    params, planes, st = scenes.load_scene(SMALL)
    sol = Solver(0, 0)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.upload([np.zeros(0, np.float32)] * 6)
    sol.step(3)
    t = np.float32(0)
    for _ in range(3):
        t = np.float32(t + np.float32(params.dt))
    assert np.float32(sol.time) == t
