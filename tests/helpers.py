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


def make_pair_from(params, planes, state, mode=PBF_MODE_STRICT, debug=True):
    """(solver, oracle, params) for an explicit parameter set / plane list / state (random clouds)."""
    orc = Oracle(best_kind())
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    sol = Solver(0, len(state[0]), mode)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.debug_enable(debug)
    sol.upload(state)
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


def random_cloud(seed: int, n: int | None = None):
    """(params, planes, state, flags) of an adversarial random particle cloud: particles snapped onto
    cell boundaries, coincident particles, negative coordinates, particles outside the planes, fast
    and resting ones, ragged counts (n not a multiple of 32).  Deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    if n is None:
        n = int(rng.choice([1, 2, 31, 32, 33, 97, 257, 600]))
    h = np.float32(0.1)
    pos = rng.uniform(-0.35, 0.65, size=(3, n)).astype(np.float32)
    snap = rng.random(n) < 0.25                       # exactly on a cell boundary along some axis
    axis = rng.integers(0, 3, size=n)
    cells = np.round(pos[axis, np.arange(n)] / h).astype(np.float32)
    pos[axis[snap], np.nonzero(snap)[0]] = (cells[snap] * h).astype(np.float32)
    if n > 3:                                          # coincident particles (r = 0 pairs)
        dup = rng.integers(0, n, size=max(1, n // 20))
        src = rng.integers(0, n, size=dup.shape[0])
        pos[:, dup] = pos[:, src]
    vel = (rng.normal(0.0, 1.5, size=(3, n)) * (rng.random(n) < 0.8)).astype(np.float32)
    params = PbfParams.defaults()
    params.particle_mass, params.density, params.h, params.epsilon = 1.0, 6000.0, float(h), 300.0
    params.scorr_k, params.scorr_n, params.visc_c = 0.0005, int(rng.choice([2, 3, 4])), 0.0005
    flags = dict(scorr=int(rng.integers(0, 2)), xsph=int(rng.integers(0, 2)), vort=int(rng.integers(0, 2)),
                 rest=float(rng.choice([0.0, 0.05])), fric=float(rng.choice([0.0, 0.1])))
    params = configure(params, flags, iterations=int(rng.choice([1, 2, 4])))
    lo, hi = -0.3, 0.6                                 # some particles start outside the box
    planes = np.array([[0, 1, 0, lo], [1, 0, 0, lo], [0, 0, 1, lo],
                       [0, -1, 0, -hi], [-1, 0, 0, -hi], [0, 0, -1, -hi]], dtype=np.float32)
    if rng.random() < 0.3:
        planes = planes[: int(rng.integers(0, 4))]     # fewer planes, possibly none
    state = [np.ascontiguousarray(pos[k]) for k in range(3)] + [np.ascontiguousarray(vel[k]) for k in range(3)]
    return params, planes, state, flags
