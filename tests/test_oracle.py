"""CPU tests: the oracles against the committed golden fixtures and against each other.

The fixtures under tests/golden/ are outputs of the unmodified reference (make_golden.py);
the C restatement (oracle/pbf_oracle.c) must reproduce them bit for bit — that is what pins
the oracle the GPU tests are judged against.
"""
import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from oracle import oracle_api
from oracle.oracle_api import Oracle

import golden_util as G
import helpers as H

FLAGSETS = {"none": H.NO_FLAGS, "stable": H.STABLE_FLAGS, "all": H.ALL_FLAGS}


def _oracle(kind, scene, flags, iterations=None):
    params, planes, state = scenes.load_scene(scene)
    params = H.configure(params, flags, iterations=iterations)
    orc = Oracle(kind)
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    return orc, params


@pytest.mark.parametrize("flagname", list(FLAGSETS))
def test_port_matches_golden_small(built, flagname):
    flags = FLAGSETS[flagname]
    gold = G.load_small(flagname)
    orc, _ = _oracle("port", scenes.small_block(10), flags)
    done = 0
    for step in sorted(gold):
        orc.step(step - done)
        done = step
        assert G.mismatches(G.snapshot_of(orc, flags, False), gold[step]) == [], f"step {step}"
        assert np.float32(orc.time) == gold[step]["time"]


@pytest.mark.parametrize("run", ["fluid_large:stable", "fluid_large:all", "fluid_large:all:iters8"])
def test_port_matches_golden_digests(built, run):
    gold = G.digests()["runs"][run]
    parts = run.split(":")
    flags = FLAGSETS[parts[1]]
    iters = 8 if len(parts) > 2 else None
    orc, _ = _oracle("port", scenes.SCENES[parts[0]], flags, iters)
    done = 0
    for step in sorted(int(s) for s in gold):
        orc.step(step - done)
        done = step
        assert G.digest_mismatches(G.snapshot_of(orc, flags, False), gold[str(step)]) == [], f"step {step}"


@pytest.mark.skipif(not oracle_api.available("reference"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("flagname", list(FLAGSETS))
def test_reference_matches_golden_small(built, flagname):
    """The fixtures are reproducible from the compiled reference on this machine."""
    flags = FLAGSETS[flagname]
    gold = G.load_small(flagname)
    orc, _ = _oracle("reference", scenes.small_block(10), flags)
    done = 0
    for step in sorted(gold):
        orc.step(step - done)
        done = step
        assert G.mismatches(G.snapshot_of(orc, flags, False), gold[step]) == []


@pytest.mark.skipif(not oracle_api.available("reference"), reason="oracle/_ref not built")
def test_port_equals_reference_on_perturbed_state(built):
    """Irregular input: jittered positions, random velocities, ragged particle count."""
    rng = np.random.default_rng(1234)
    params, planes, state = scenes.load_scene(scenes.small_block(11))
    n = len(state[0]) - 37
    h = np.float32(params.h)
    state = [a[:n].copy() for a in state]
    for k in range(3):
        state[k] += (rng.standard_normal(n).astype(np.float32) * np.float32(0.2) * h)
        state[3 + k] = (rng.standard_normal(n) * 0.5).astype(np.float32)
    params = H.configure(params, H.ALL_FLAGS)
    pair = []
    for kind in ("port", "reference"):
        orc = Oracle(kind)
        orc.set_params(params)
        orc.set_planes(planes)
        orc.set_state(state)
        pair.append(orc)
    for step in range(6):
        for orc in pair:
            orc.step(1)
        a = G.snapshot_of(pair[0], H.ALL_FLAGS, False)
        b = G.snapshot_of(pair[1], H.ALL_FLAGS, False)
        assert G.mismatches(a, b) == [], f"step {step + 1}"


def test_oracle_thread_count_invariance(built):
    """The oracle is bitwise identical for 1 and many threads."""
    outs = []
    for threads in (1, 4):
        orc, _ = _oracle("port", scenes.small_block(10), H.ALL_FLAGS)
        orc.set_threads(threads)
        orc.step(5)
        outs.append(G.snapshot_of(orc, H.ALL_FLAGS, False))
    assert G.mismatches(outs[0], outs[1]) == []


def test_oracle_empty_state(built):
    orc = Oracle("port")
    p = scenes.load_scene(scenes.small_block(10))[0]
    orc.set_params(p)
    orc.set_state([np.zeros(0, np.float32)] * 6)
    orc.step(2)
    assert orc.count() == 0
    assert np.float32(orc.time) == np.float32(np.float32(p.dt) + np.float32(p.dt))
