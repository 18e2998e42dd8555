"""Bit-compare a small block against the oracle for flag sets / step counts (debug aid).
usage: dbg_smoke.py [edge] [steps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_STRICT, Solver
from oracle.oracle_api import Oracle, best_kind

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for flags in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 0, 1)):
    params, planes, state = scenes.load_scene(scenes.small_block(edge))
    params.dt = np.float32(1.0 / 120.0)
    params.enable_scorr, params.enable_xsph, params.enable_vorticity = flags
    params.plane_restitution, params.plane_friction = 0.05, 0.1
    sol = Solver(0, len(state[0]), PBF_MODE_STRICT)
    sol.set_params(params); sol.set_planes(planes); sol.upload(state)
    orc = Oracle(best_kind()); orc.set_params(params); orc.set_planes(planes); orc.set_state(state)
    for s in range(steps):
        sol.step(1); orc.step(1)
        bad = [n for n, a, b in zip("px py pz vx vy vz".split(), sol.download(), orc.get_state())
               if not np.array_equal(a.view(np.uint32), b.view(np.uint32))]
        nbad = 0
        if bad:
            a, b = sol.download()[0], orc.get_state()[0]
            nbad = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        print(f"flags {flags} step {s+1}: {'ok' if not bad else 'DIFF ' + str(bad) + ' n=' + str(nbad)} {sol.brick_status()}", flush=True)
        if bad:
            break
    sol.close()
