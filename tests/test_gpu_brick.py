"""The shared-memory staged brick kernels (kernels/brick.cu: TMA bulk copies into tiles, LDS.128
gathers, 16-bit tile-relative lists) against the CPU oracle, both drivers.  The brick family is not
the default (the global-gather kernels measured faster on B200), so it gets its own
parity tests: bit-identical state, scratch arrays, cell table and neighbour lists."""
import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_BRICK_OFF, PBF_BRICK_PER_CTA, PBF_BRICK_PERSISTENT

import helpers as H

pytestmark = pytest.mark.gpu

MODES = {"persistent": PBF_BRICK_PERSISTENT, "per_cta": PBF_BRICK_PER_CTA}


@pytest.mark.parametrize("mode", list(MODES), ids=list(MODES))
@pytest.mark.parametrize("flags", [H.NO_FLAGS, H.STABLE_FLAGS, H.ALL_FLAGS], ids=["none", "stable", "all"])
def test_brick_bit_exact_small(built, mode, flags):
    sol, orc, params = H.make_pair(scenes.small_block(12), flags)
    sol.set_brick(MODES[mode])
    for step in range(1, 9):
        sol.step(1)
        orc.step(1)
        # with vorticity the reference diverges after a few substeps and a tile may outgrow
        # its capacity: that batch is replayed on the global-gather kernels (tested below)
        if not flags["vort"] or step <= 3:
            assert sol.brick_status()["active"], f"step {step}: the batch did not run on the brick path"
        assert H.compare_integers(sol, orc) == [], f"step {step}"
        assert H.compare_scratch_bits(sol, orc, params) == [], f"step {step}"
        assert H.compare_state_bits(sol, orc) == [], f"step {step}"


@pytest.mark.parametrize("mode", list(MODES), ids=list(MODES))
def test_brick_scene_free_running(built, mode):
    """fluid_large (19 683 particles, many bricks per persistent CTA ring): one 30-substep batch."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, debug=False)
    sol.set_brick(MODES[mode])
    sol.step(30)
    orc.step(30)
    assert sol.brick_status()["active"]
    assert H.compare_state_bits(sol, orc) == []
    assert H.compare_integers(sol, orc) == []


@pytest.mark.parametrize("mode", list(MODES), ids=list(MODES))
def test_brick_random_clouds(built, mode):
    """Ragged counts, coincident particles, particles on cell boundaries (tests/helpers.random_cloud)."""
    for seed in range(6):
        params, planes, state, flags = H.random_cloud(seed)
        sol, orc, params = H.make_pair_from(params, planes, state)
        sol.set_brick(MODES[mode])
        for step in range(3):
            sol.step(1)
            orc.step(1)
            assert H.compare_integers(sol, orc) == [], f"seed {seed} step {step}"
            assert H.compare_state_bits(sol, orc) == [], f"seed {seed} step {step}"


def test_brick_falls_back_when_the_reference_diverges(built):
    """With vorticity on the reference blows up: tiles overflow / the cell table turns
    sparse, the batch is replayed on the global-gather kernels, the results stay bit-identical."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.ALL_FLAGS, debug=False)
    sol.set_brick(PBF_BRICK_PERSISTENT)
    for _ in range(8):
        sol.step(5)
        orc.step(5)
        assert H.compare_state_bits(sol, orc) == []
    assert sol.brick_status()["fallbacks"] >= 1


def test_brick_modes_agree_with_default(built):
    """Switching the family between batches of one context changes nothing."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, debug=False)
    for mode in (PBF_BRICK_PERSISTENT, PBF_BRICK_OFF, PBF_BRICK_PER_CTA, PBF_BRICK_OFF):
        sol.set_brick(mode)
        sol.step(4)
        orc.step(4)
        assert H.compare_state_bits(sol, orc) == []
