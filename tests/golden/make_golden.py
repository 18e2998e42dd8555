"""Generates the committed golden fixtures from the UNMODIFIED reference CPU solver.

Run in the build container (needs /root/reference -> oracle/_ref):
    python tests/golden/make_golden.py

The reference ships no golden vectors, known-answer tests or fixtures of its own
(SURVEY.md §4), so these are outputs of the reference itself: fluid::init_scene_from_json
and fluid::step executed through oracle/ref_shim.cpp.  Two kinds of fixture:
  * small_*.npz      full arrays (state, grid tables, neighbour lists, scratch) of a
                     ~1000-particle block after a few substeps, several flag sets;
  * digests.json     sha256 of the same arrays for the reference scenes (too big to
                     store), plus the emitted particle sets of all five scenes.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from fluidsimulator_b200 import scenes  # noqa: E402
from fluidsimulator_b200.capi import SCRATCH_IDS  # noqa: E402
from oracle.oracle_api import Oracle  # noqa: E402
import helpers as H  # noqa: E402

OUT = Path(__file__).resolve().parent
FLAGSETS = {"none": H.NO_FLAGS, "stable": H.STABLE_FLAGS, "all": H.ALL_FLAGS}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def snapshot(orc: Oracle) -> dict:
    snap = {}
    for name, a in zip(["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"], orc.get_state()):
        snap[name] = a
    for k, v in orc.grid().items():
        snap["grid_" + k] = v
    prefix, idx = orc.neighbors()
    snap["neighbor_prefix_sum"] = prefix
    snap["neighbor_indices"] = idx
    for name in SCRATCH_IDS:
        snap["scratch_" + name] = orc.scratch(name)
    snap["time"] = np.float32(orc.time)
    return snap


def run(scene, flags, steps, iterations=None):
    params, planes, state = scenes.load_scene(scene)
    params = H.configure(params, flags, iterations=iterations)
    orc = Oracle("reference")
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    out = {}
    done = 0
    for s in steps:
        orc.step(s - done)
        done = s
        out[s] = snapshot(orc)
    return out


def main():
    digests = {"scenes": {}, "runs": {}}
    # 1. emitted particle sets of the five reference scenes, through the REFERENCE loader
    for name in scenes.SCENES:
        orc = Oracle("reference")
        orc.load_scene(f"/root/reference/scene/{name}.json")
        st = orc.get_state()
        p = orc.get_params()
        digests["scenes"][name] = {
            "count": int(st[0].shape[0]),
            "pos": [digest(a) for a in st[:3]],
            "planes": digest(orc.get_planes()),
            "params": {k: (v if not isinstance(v, list) else v) for k, v in p.as_dict().items()},
        }
    # 2. full small fixtures
    small = scenes.small_block(10)
    for fname, flags in FLAGSETS.items():
        snaps = run(small, flags, [1, 5, 20])
        flat = {}
        for step, snap in snaps.items():
            for k, v in snap.items():
                flat[f"s{step}_{k}"] = v
        np.savez_compressed(OUT / f"small_{fname}.npz", **flat)
    # 3. digests on a real scene
    for fname, flags, steps in (("stable", H.STABLE_FLAGS, [1, 2, 10]), ("all", H.ALL_FLAGS, [1, 2, 3])):
        snaps = run(scenes.SCENES["fluid_large"], flags, steps)
        digests["runs"][f"fluid_large:{fname}"] = {
            str(step): {k: digest(np.asarray(v)) for k, v in snap.items()} for step, snap in snaps.items()}
    snaps = run(scenes.SCENES["fluid_large"], H.ALL_FLAGS, [2], iterations=8)
    digests["runs"]["fluid_large:all:iters8"] = {"2": {k: digest(np.asarray(v)) for k, v in snaps[2].items()}}
    # 4. the parity configurations BASELINE.json names (configs[0..2]) and the reference's vorticity
    #    blow-up: by substep 20-25 the box of occupied cells spans kilometres
    for key, scene, flags, steps, iters in (
            ("fluid_double_dem:all", "fluid_double_dem", H.ALL_FLAGS, [1, 5, 20], None),
            ("fluid_double_side:all", "fluid_double_side", H.ALL_FLAGS, [1, 4, 12], None),
            ("fluid_large:all:blowup", "fluid_large", H.ALL_FLAGS, [25], None),
            ("fluid_large:stable:iters2", "fluid_large", H.STABLE_FLAGS, [6], 2),
            ("fluid_xlarge:stable", "fluid_xlarge", H.STABLE_FLAGS, [5], None)):
        snaps = run(scenes.SCENES[scene], flags, steps, iterations=iters)
        digests["runs"][key] = {
            str(step): {k: digest(np.asarray(v)) for k, v in snap.items()} for step, snap in snaps.items()}
    (OUT / "digests.json").write_text(json.dumps(digests, indent=1, sort_keys=True))
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
