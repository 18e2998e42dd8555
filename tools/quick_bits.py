"""Smallest possible bit check on the GPU box (a couple of seconds): fluid_large, all flags,
40 substeps (blow-up, sparse table) must give the state digest recorded for the kernels of
profiles/ab_r01g_neighbors_mask.txt; then, time permitting, ms/substep of fluid_million."""
import sys
import time
from pathlib import Path

t00 = time.perf_counter()
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import numpy as np
from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_STRICT, Solver
from quick_ab import digest

EXPECT_LARGE_ALL_40 = "e0e7639c7c0fa35e"
EXPECT_MILLION_280 = "0fc7fad13d5e3129"


def setup(scene, vort):
    params, planes, state = scenes.load_scene(scenes.SCENES[scene])
    params.dt = np.float32(1.0 / 120.0)
    params.enable_scorr, params.enable_xsph, params.enable_vorticity = 1, 1, vort
    params.plane_restitution, params.plane_friction = 0.05, 0.1
    sol = Solver(0, len(state[0]), PBF_MODE_STRICT)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.upload(state)
    return sol


sol = setup("fluid_large", 1)
sol.step(3)
print("large_all_3 brick", sol.brick_status(), flush=True)
sol.step(37)
print("large_all_40 brick", sol.brick_status(), flush=True)
d = digest(sol.download())
print(f"large_all_40 {d} {'IDENTICAL' if d == EXPECT_LARGE_ALL_40 else 'DIFFERS'} ({time.perf_counter() - t00:.1f} s)", flush=True)
sol.close()
sol = setup("fluid_million", 0)
sol.step(5)
t0 = time.perf_counter()
sol.step(60)
print(f"million ms_t0 {(time.perf_counter() - t0) * 1e3 / 60:.4f} ({time.perf_counter() - t00:.1f} s)", flush=True)
sol.step(135)
t0 = time.perf_counter()
sol.step(60)
print(f"million ms_200 {(time.perf_counter() - t0) * 1e3 / 60:.4f} ({time.perf_counter() - t00:.1f} s)", flush=True)
sol.profile_enable(True)
sol.profile_reset()
sol.step(20)
prof = sol.profile()
print("stage_us", {k: round(1e3 * v["ms"] / v["launches"], 1) for k, v in prof.items() if v["launches"]}, flush=True)
print("million brick", sol.brick_status(), flush=True)
d = digest(sol.download())
print(f"million_280 {d} {'IDENTICAL' if d == EXPECT_MILLION_280 else 'DIFFERS'} ({time.perf_counter() - t00:.1f} s)", flush=True)
