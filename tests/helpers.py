"""Shared helpers for the parity tests: run the CUDA path and an oracle side by side."""
from __future__ import annotations

import numpy as np

from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_FAST, PBF_MODE_STRICT, SCRATCH_IDS, PbfParams, Solver
from oracle.oracle_api import Oracle, best_kind

ALL_FLAGS = dict(scorr=1, xsph=1, vort=1, rest=0.05, fric=0.1)   # the README example flags
STABLE_FLAGS = dict(scorr=1, xsph=1, vort=0, rest=0.05, fric=0.1)
NO_FLAGS = dict(scorr=0, xsph=0, vort=0, rest=0.0, fric=0.0)


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def bit_equal(a, b) -> bool:
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def configure(params: PbfParams, flags: dict, steps_per_sec: float = 120.0, iterations: int | None = None) -> PbfParams:
    p = params.copy()
    p.dt = np.float32(1.0 / steps_per_sec)   # main.cpp:203-205: static_cast<float>(1.0 / steps_per_sec)
    p.enable_scorr, p.enable_xsph, p.enable_vorticity = flags["scorr"], flags["xsph"], flags["vort"]
    p.plane_restitution, p.plane_friction = flags["rest"], flags["fric"]
    if iterations is not None:
        p.solver_iterations = iterations
    return p


def make_pair(scene, flags, mode=PBF_MODE_STRICT, iterations=None, oracle_kind=None, debug=True, state=None):
    """(solver, oracle) loaded with the same scene, flags and state."""
    params, planes, st = scenes.load_scene(scene)
    if state is not None:
        st = state
    params = configure(params, flags, iterations=iterations)
    orc = Oracle(oracle_kind or best_kind())
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(st)
    sol = Solver(0, len(st[0]), mode)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.debug_enable(debug)
    sol.upload(st)
    return sol, orc, params


def compare_integers(sol: Solver, orc: Oracle) -> list[str]:
    """Bit-exact gate G1: entries, cell table, neighbour lists."""
    bad = []
    g_s, g_o = sol.debug_grid(), orc.grid()
    for k in g_o:
        if not np.array_equal(g_s[k], g_o[k]):
            bad.append(f"grid.{k}")
    p_s, i_s = sol.debug_neighbors()
    p_o, i_o = orc.neighbors()
    if not np.array_equal(p_s, p_o):
        bad.append("neighbor_prefix_sum")
    if not np.array_equal(i_s, i_o):
        bad.append("neighbor_indices")
    return bad


def compare_scratch_bits(sol: Solver, orc: Oracle, params: PbfParams) -> list[str]:
    names = ["pred_x", "pred_y", "pred_z", "delta_x", "delta_y", "delta_z", "lambda", "rho"]
    if params.enable_xsph and params.visc_c != 0:
        names += ["dv_x", "dv_y", "dv_z"]
    if params.enable_vorticity and params.vort_epsilon != 0:
        names += ["omega_x", "omega_y", "omega_z", "omega_mag", "eta_x", "eta_y", "eta_z"]
    return [n for n in names if not bit_equal(sol.debug_scratch(n), orc.scratch(n))]


def compare_state_bits(sol: Solver, orc: Oracle) -> list[str]:
    names = ["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"]
    return [n for n, a, b in zip(names, sol.download(), orc.get_state()) if not bit_equal(a, b)]


def max_abs_pos_err_in_h(sol: Solver, orc: Oracle, h: float) -> float:
    a, b = sol.download(), orc.get_state()
    return float(max(np.max(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))) for k in range(3)) / h)


def max_abs_vel_err(sol: Solver, orc: Oracle) -> float:
    a, b = sol.download(), orc.get_state()
    return float(max(np.max(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))) for k in range(3, 6)))
