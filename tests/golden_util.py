"""Loading / comparing the committed golden fixtures (tests/golden/, made by make_golden.py)."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

from fluidsimulator_b200.capi import SCRATCH_IDS

GOLDEN = Path(__file__).resolve().parent / "golden"
STATE = ["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"]
GRID = ["entry_cx", "entry_cy", "entry_cz", "entry_particle", "cell_xyz", "cell_start", "cell_end"]


def digests() -> dict:
    return json.loads((GOLDEN / "digests.json").read_text())


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_small(flagname: str) -> dict:
    z = np.load(GOLDEN / f"small_{flagname}.npz")
    out: dict = {}
    for key in z.files:
        step, name = key.split("_", 1)
        out.setdefault(int(step[1:]), {})[name] = z[key]
    return out


def scratch_names(flags: dict) -> list[str]:
    names = ["pred_x", "pred_y", "pred_z", "delta_x", "delta_y", "delta_z", "lambda", "rho"]
    if flags["xsph"]:
        names += ["dv_x", "dv_y", "dv_z"]
    if flags["vort"]:
        names += ["omega_x", "omega_y", "omega_z", "omega_mag", "eta_x", "eta_y", "eta_z"]
    return names


def snapshot_of(obj, flags: dict, is_solver: bool) -> dict:
    """Same keys as make_golden.snapshot, from a Solver or an Oracle."""
    snap = {}
    state = obj.download() if is_solver else obj.get_state()
    for name, a in zip(STATE, state):
        snap[name] = a
    grid = obj.debug_grid() if is_solver else obj.grid()
    for k, v in grid.items():
        snap["grid_" + k] = v
    prefix, idx = obj.debug_neighbors() if is_solver else obj.neighbors()
    snap["neighbor_prefix_sum"], snap["neighbor_indices"] = prefix, idx
    for name in scratch_names(flags):
        snap["scratch_" + name] = obj.debug_scratch(name) if is_solver else obj.scratch(name)
    return snap


def mismatches(snap: dict, gold: dict) -> list[str]:
    """Keys of `snap` whose bytes differ from the golden arrays."""
    bad = []
    for k, v in snap.items():
        g = gold[k]
        a, b = np.ascontiguousarray(v), np.ascontiguousarray(g)
        if a.shape != b.shape or a.tobytes() != b.tobytes():
            bad.append(k)
    return bad


def digest_mismatches(snap: dict, gold: dict) -> list[str]:
    return [k for k, v in snap.items() if digest(np.asarray(v)) != gold[k]]
