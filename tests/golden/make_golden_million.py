"""Full-size golden digests: the UNMODIFIED reference CPU solver on scene/fluid_million.json
(1 000 000 particles, BASELINE.json's headline workload), stable flags, free-running.

    python tests/golden/make_golden_million.py        # ~15 minutes on 8 cores

Writes tests/golden/million.json: sha256 of the six state arrays after 65 and 280 substeps, and
the combined 16-hex digest tools/quick_ab.py prints (sha256 over the six arrays in a row), so the
digest a GPU run logged can be compared with the reference without re-running either.
Kept apart from make_golden.py because of its run time."""
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

STEPS = [65, 280]
NAMES = ["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"]


def main():
    params, planes, state = scenes.load_scene(scenes.SCENES["fluid_million"])
    params = H.configure(params, H.STABLE_FLAGS)
    orc = Oracle("reference")
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    out = {"scene": "fluid_million", "flags": "stable", "particles": len(state[0]), "steps": {}}
    done = 0
    t0 = time.perf_counter()
    for s in STEPS:
        while done < s:
            orc.step(1)
            done += 1
            if done % 10 == 0:
                print(f"substep {done} ({time.perf_counter() - t0:.0f} s)", flush=True)
        st = orc.get_state()
        combined = hashlib.sha256()
        for a in st:
            combined.update(np.ascontiguousarray(a).tobytes())
        out["steps"][str(s)] = {**{n: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for n, a in zip(NAMES, st)},
                                "combined16": combined.hexdigest()[:16]}
        (Path(__file__).resolve().parent / "million.json").write_text(json.dumps(out, indent=1, sort_keys=True))
        print(f"substep {s}: combined16 {out['steps'][str(s)]['combined16']}", flush=True)


if __name__ == "__main__":
    main()
