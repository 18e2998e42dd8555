#!/usr/bin/env python
"""Parity harness (row f2 of SURVEY §8: the "frame dumps + diff tool" the reference's AGENTS.md asks
for and never got).  Runs the CUDA backend and the CPU oracle side by side on any scene / flags /
mode and prints, per checkpoint: bit-equality of the integer tables (cell keys, sorted order, cell
start/end, neighbour lists), and max-abs / RMS error of positions (in units of h) and velocities.

  python tests/parity_tool.py --scene fluid_large --flags all --steps 20 --every 5 --mode strict
  python tests/parity_tool.py --scene scene.json --flags stable --steps 50 --mode fast --slabs 3

TEST INFRASTRUCTURE: it loads oracle/ (the checker) and therefore lives under tests/.
Exit status 0 when every STRICT checkpoint is bit-identical / every FAST checkpoint is within
--tol (default 1e-4 h), 1 otherwise.  Needs a GPU."""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers as H  # noqa: E402
from fluidsimulator_b200 import scenes  # noqa: E402
from fluidsimulator_b200.capi import PBF_MODE_FAST, PBF_MODE_STRICT, SlabGroup, Solver  # noqa: E402
from oracle.oracle_api import Oracle, best_kind  # noqa: E402

FLAGS = {"none": H.NO_FLAGS, "stable": H.STABLE_FLAGS, "all": H.ALL_FLAGS}


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--scene", default="fluid_large", help="a shipped scene name, small_<n>, or a scene JSON file")
    ap.add_argument("--flags", default="all", choices=list(FLAGS))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--every", type=int, default=5, help="compare every N substeps")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--iterations", type=int, default=None)
    ap.add_argument("--slabs", type=int, default=1, help="> 1: x-slabs on cuda:0 (virtual ranks)")
    ap.add_argument("--tol", type=float, default=1e-4, help="FAST mode gate: max-abs position error in h")
    args = ap.parse_args()

    if args.scene in scenes.SCENES:
        sc = scenes.SCENES[args.scene]
    elif args.scene.startswith("small_"):
        sc = scenes.small_block(int(args.scene.split("_")[1]))
    else:
        sc = args.scene  # a scene JSON in the reference's format (scenes.load_scene reads it)
    params, planes, state = scenes.load_scene(sc)
    params = H.configure(params, FLAGS[args.flags], iterations=args.iterations)
    mode = PBF_MODE_STRICT if args.mode == "strict" else PBF_MODE_FAST
    orc = Oracle(best_kind())
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    if args.slabs > 1:
        sol = SlabGroup([0] * args.slabs, params, planes, mode)
        sol.upload(state)
    else:
        sol = Solver(0, len(state[0]), mode)
        sol.set_params(params)
        sol.set_planes(planes)
        sol.upload(state)
    h, dt = float(params.h), float(params.dt)
    print(f"scene {args.scene}: {len(state[0])} particles, flags {args.flags}, mode {args.mode}, slabs {args.slabs}, "
          f"oracle '{orc.kind}'")
    print(f"{'step':>5} {'tables':>10} {'floats':>12} {'max|dp|/h':>11} {'rms|dp|/h':>11} {'max|dv| m/s':>12}")
    ok, done = True, 0
    while done < args.steps:
        k = min(args.every, args.steps - done)
        sol.step(k)
        orc.step(k)
        done += k
        a, b = sol.download(), orc.get_state()
        dp = np.stack([a[c].astype(np.float64) - b[c].astype(np.float64) for c in range(3)])
        dv = np.stack([a[c].astype(np.float64) - b[c].astype(np.float64) for c in range(3, 6)])
        bits = all(H.bit_equal(x, y) for x, y in zip(a, b))
        tables = "n/a"
        if args.slabs == 1:
            tables = "identical" if H.compare_integers(sol, orc) == [] else "DIFFER"
        max_dp, rms_dp = np.abs(dp).max() / h, np.sqrt((dp ** 2).sum(axis=0).mean()) / h
        print(f"{done:5d} {tables:>10} {'bit-exact' if bits else 'differ':>12} {max_dp:11.3e} {rms_dp:11.3e} "
              f"{np.abs(dv).max():12.3e}")
        if args.mode == "strict":
            ok &= bits and tables != "DIFFER"
        else:
            ok &= max_dp <= args.tol and np.abs(dv).max() <= args.tol * h / dt
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
