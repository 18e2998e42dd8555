"""Full-size golden digests: the UNMODIFIED reference CPU solver (oracle/_ref) on the workloads
bench.py measures, free-running.

    python tests/golden/make_golden_million.py [run ...]     # all runs: ~25 minutes on 8 cores

Runs (tests/golden/million.json, one entry each; existing entries are kept unless re-run):
    fluid_million:stable   scene/fluid_million.json (BASELINE.json's headline workload), stable flags,
                           substeps 65 and 280
    fluid_million:all      the same scene with vorticity on, substep 3 (afterwards the reference blows up)
    block_16m:stable       the synthetic 252^3 block of the multi-GPU runs, substep 2
    fluid_double_dem:all   BASELINE.json configs[0] (142 560 particles, all flags), substeps 40 and 80:
                           far into the reference's own divergence (sparse cell table, K in the thousands)
Per step: sha256 of the six state arrays, and the combined 16-hex digest tools/quick_ab.py prints
(sha256 over the six arrays in a row), so that a digest logged by a GPU run can be compared with the
reference without re-running either.  Kept apart from make_golden.py because of its run time."""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from fluidsimulator_b200 import scenes  # noqa: E402
from oracle.oracle_api import Oracle  # noqa: E402
import helpers as H  # noqa: E402

OUT = Path(__file__).resolve().parent / "million.json"
NAMES = ["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"]
RUNS = {
    "fluid_million:stable": (lambda: scenes.SCENES["fluid_million"], H.STABLE_FLAGS, [65, 280]),
    "fluid_million:all": (lambda: scenes.SCENES["fluid_million"], H.ALL_FLAGS, [3]),
    "block_16m:stable": (scenes.block_16m, H.STABLE_FLAGS, [2]),
    # BASELINE.json configs[0] (the CPU parity reference run: fluid_double_dem, all flags), deep into the
    # reference's own vorticity blow-up (~10 000 particles in one cell by substep 80)
    "fluid_double_dem:all": (lambda: scenes.SCENES["fluid_double_dem"], H.ALL_FLAGS, [40, 80]),
}


def run(key: str) -> dict:
    scene, flags, steps = RUNS[key]
    params, planes, state = scenes.load_scene(scene())
    params = H.configure(params, flags)
    orc = Oracle("reference")
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    out = {"particles": len(state[0]), "steps": {}}
    done = 0
    t0 = time.perf_counter()
    for s in steps:
        while done < s:
            orc.step(1)
            done += 1
            print(f"{key}: substep {done} ({time.perf_counter() - t0:.0f} s)", flush=True)
        st = orc.get_state()
        combined = hashlib.sha256()
        for a in st:
            combined.update(np.ascontiguousarray(a).tobytes())
        out["steps"][str(s)] = {**{n: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for n, a in zip(NAMES, st)},
                                "combined16": combined.hexdigest()[:16]}
    return out


def main():
    doc = json.loads(OUT.read_text()) if OUT.exists() else {"runs": {}}
    for key in sys.argv[1:] or list(RUNS):
        doc["runs"][key] = run(key)
        OUT.write_text(json.dumps(doc, indent=1, sort_keys=True))
        print(f"{key}: {[(s, v['combined16']) for s, v in doc['runs'][key]['steps'].items()]}", flush=True)


if __name__ == "__main__":
    main()
