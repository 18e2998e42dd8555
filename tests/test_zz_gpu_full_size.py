"""Full-size parity on the GPU (named to run after every other test file: these are the longest
GPU tests, and the all-flags and 16 M entries were generated after the round's last GPU minute, so
their first GPU run is the round-end one)."""
import numpy as np  # noqa: F401
import pytest

from fluidsimulator_b200 import scenes

import golden_util as G
import helpers as H

pytestmark = pytest.mark.gpu
FLAGSETS = {"none": H.NO_FLAGS, "stable": H.STABLE_FLAGS, "all": H.ALL_FLAGS}


@pytest.mark.parametrize("run", ["fluid_million:stable", "fluid_million:all", "block_16m:stable", "fluid_double_dem:all"])
def test_full_size_runs_match_the_reference_digests(built, run):
    """The workloads bench.py measures, at full size, against the reference itself: free-running
    state after N substeps must hash to what the unmodified reference CPU solver produced
    (tests/golden/million.json, written by tests/golden/make_golden_million.py — a 25-minute CPU
    run, so the digests are committed): BASELINE.json's fluid_million (1 000 000 particles) through
    280 substeps, the same scene with vorticity for the 3 substeps before the reference blows up,
    the 16 M-particle block of the multi-GPU runs, and BASELINE.json's configs[0] (fluid_double_dem,
    all flags) 80 substeps deep into the reference's own blow-up."""
    import json
    from fluidsimulator_b200.capi import Solver
    path = G.GOLDEN / "million.json"
    gold = json.loads(path.read_text())["runs"] if path.exists() else {}
    if run not in gold:
        pytest.skip(f"tests/golden/million.json has no entry {run}")
    gold = gold[run]
    scene_name, flagname = run.split(":")
    scene = scenes.block_16m() if scene_name == "block_16m" else scenes.SCENES[scene_name]
    params, planes, state = scenes.load_scene(scene)
    params = H.configure(params, FLAGSETS[flagname])
    assert len(state[0]) == gold["particles"]
    sol = Solver(0, len(state[0]))
    sol.set_params(params)
    sol.set_planes(planes)
    sol.upload(state)
    done = 0
    for step in sorted(int(k) for k in gold["steps"]):
        sol.step(step - done)
        done = step
        got = {name: G.digest(a) for name, a in zip(G.STATE, sol.download())}
        bad = [name for name in G.STATE if got[name] != gold["steps"][str(step)][name]]
        assert bad == [], f"{run}, substep {step}: {bad} differ from the reference"
    sol.close()
