"""Seconds-scale look at the brick path on the GPU box (no torch): fluid_million, stable flags, STRICT.
ms/substep and per-stage launch times at t0 and settled, the largest tile along the trajectory, how
many batches fell back to the global-gather family, and the state digest after 280 substeps (must be
tests/golden/million.json's 0fc7fad13d5e3129).  PBF_B200_LIB selects a variant library, PBF_BRICK=0|1|2 the
global-gather family (default) / persistent bricks / one CTA per brick.   python tools/quick_brick.py [chunk]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import numpy as np
from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_STRICT, Solver
from quick_ab import digest

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 15
params, planes, state = scenes.load_scene(scenes.SCENES["fluid_million"])
params.dt = np.float32(1.0 / 120.0)
params.enable_scorr, params.enable_xsph, params.enable_vorticity = 1, 1, 0
params.plane_restitution, params.plane_friction = 0.05, 0.1
sol = Solver(0, len(state[0]), PBF_MODE_STRICT)
sol.set_params(params)
sol.set_planes(planes)
sol.upload(state)


def timed(steps):
    t0 = time.perf_counter()
    sol.step(steps)
    return (time.perf_counter() - t0) * 1e3 / steps


def stages(steps=10):
    sol.profile_enable(True)
    sol.profile_reset()
    sol.step(steps)
    prof = sol.profile()
    sol.profile_enable(False)
    return {k: round(1e3 * v["ms"] / v["launches"], 1) for k, v in prof.items() if v["launches"]}


done = 0
sol.step(5); done += 5
print(f"ms_t0 {timed(40):.4f}", sol.brick_status(), flush=True); done += 40
print("stage_us_t0", stages(), sol.brick_status(), flush=True); done += 10
while done < 200:
    ms = timed(chunk); done += chunk
    print(f"  substep {done:3d}: {ms:.4f} ms/substep", sol.brick_status(), flush=True)
print(f"ms_settled {timed(40):.4f}", sol.brick_status(), flush=True); done += 40
print("stage_us_settled", stages(), sol.brick_status(), flush=True); done += 10
sol.step(280 - done)
d = digest(sol.download())
print(f"million_280 {d} {'IDENTICAL' if d == '0fc7fad13d5e3129' else 'DIFFERS'}", sol.brick_status(), flush=True)
