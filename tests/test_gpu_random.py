"""GPU: the CUDA path against the oracle on adversarial random clouds (see helpers.random_cloud):
particles exactly on cell boundaries, coincident particles, 1 / 31 / 33 / 257 particles, particles
outside the planes, random flags, iteration counts, s_corr exponents and plane sets — STRICT,
every array bit for bit, one context and three slabs."""
import numpy as np
import pytest

from fluidsimulator_b200.capi import PBF_MODE_STRICT, SlabGroup, Solver
from oracle.oracle_api import Oracle, best_kind

import golden_util as G
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(16))
def test_strict_equals_oracle_on_random_clouds(built, seed):
    params, planes, state, flags = H.random_cloud(seed)
    orc = Oracle(best_kind())
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    sol = Solver(0, len(state[0]), PBF_MODE_STRICT)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.debug_enable(True)
    sol.upload(state)
    for step in range(1, 4):
        sol.step(1)
        orc.step(1)
        bad = G.mismatches(G.snapshot_of(sol, flags, True), G.snapshot_of(orc, flags, False))
        assert bad == [], f"seed {seed} (n = {len(state[0])}, flags {flags}), step {step}"


@pytest.mark.parametrize("seed", [3, 5, 7, 11])
def test_slabs_equal_oracle_on_random_clouds(built, seed):
    params, planes, state, flags = H.random_cloud(seed, n=1500)
    orc = Oracle(best_kind())
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    grp = SlabGroup([0, 0, 0], params, planes)
    grp.upload(state)
    grp.step(3)
    orc.step(3)
    for name, a, b in zip(G.STATE, grp.download(), orc.get_state()):
        assert H.bit_equal(a, b), f"seed {seed}: {name}"
    grp.close()
