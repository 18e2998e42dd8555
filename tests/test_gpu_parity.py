"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle.

STRICT mode is required to be bit-identical to the reference CPU path on every integer
table AND every float array, free-running over many substeps with all flags on.
FAST mode keeps the integer tables bit-exact for one teacher-forced substep and is
gated by the tolerance BASELINE.json states: max-abs position error <= 1e-4 * h.
"""
import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_FAST, PBF_MODE_STRICT

import helpers as H

pytestmark = pytest.mark.gpu

SMALL = scenes.small_block(12)


@pytest.mark.parametrize("flags", [H.NO_FLAGS, H.STABLE_FLAGS, H.ALL_FLAGS], ids=["none", "stable", "all"])
def test_strict_bit_exact_small(built, flags):
    sol, orc, params = H.make_pair(SMALL, flags)
    for step in range(1, 13):
        sol.step(1)
        orc.step(1)
        assert H.compare_integers(sol, orc) == [], f"step {step}"
        assert H.compare_scratch_bits(sol, orc, params) == [], f"step {step}"
        assert H.compare_state_bits(sol, orc) == [], f"step {step}"
    assert np.float32(sol.time) == np.float32(orc.time)


@pytest.mark.parametrize("scene", ["fluid_large", "fluid_double_side"])
def test_strict_bit_exact_scene_all_flags(built, scene):
    sol, orc, params = H.make_pair(scenes.SCENES[scene], H.ALL_FLAGS)
    for step in range(1, 4):
        sol.step(1)
        orc.step(1)
        assert H.compare_integers(sol, orc) == [], f"step {step}"
        assert H.compare_scratch_bits(sol, orc, params) == [], f"step {step}"
        assert H.compare_state_bits(sol, orc) == [], f"step {step}"


def test_strict_free_running_batch(built):
    """One pbf_step(30) batch (graph replay, no host sync) == 30 oracle steps, bit for bit."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, debug=False)
    sol.step(30)
    orc.step(30)
    assert H.compare_state_bits(sol, orc) == []
    assert H.compare_integers(sol, orc) == []


@pytest.mark.parametrize("iterations", [2, 8])
def test_strict_iteration_sweep(built, iterations):
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.ALL_FLAGS, iterations=iterations)
    sol.step(2)
    orc.step(2)
    assert H.compare_state_bits(sol, orc) == []
    assert H.compare_scratch_bits(sol, orc, params) == []


def test_fast_mode_tolerance(built):
    """FAST mode: integer tables exact on the first substep; positions within 1e-4*h after 10
    free-running substeps without vorticity (SURVEY §8c gate G3)."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, mode=PBF_MODE_FAST)
    sol.step(1)
    orc.step(1)
    assert H.compare_integers(sol, orc) == []
    sol.step(9)
    orc.step(9)
    err = H.max_abs_pos_err_in_h(sol, orc, params.h)
    assert err <= 1e-4, f"max-abs position error {err:.3e} h"
    assert H.max_abs_vel_err(sol, orc) <= 1e-4 * params.h / params.dt


def test_fast_mode_vorticity_one_step(built):
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.ALL_FLAGS, mode=PBF_MODE_FAST)
    sol.step(1)
    orc.step(1)
    assert H.compare_integers(sol, orc) == []
    assert H.max_abs_pos_err_in_h(sol, orc, params.h) <= 1e-4


def test_step_host_contract(built):
    """pbf_step_host == the reference cuda_step contract: host arrays in, host arrays out."""
    sol, orc, params = H.make_pair(SMALL, H.STABLE_FLAGS)
    st = [a.copy() for a in orc.get_state()]
    for _ in range(3):
        sol.step_host(st, 1)
        orc.step(1)
    for a, b in zip(st, orc.get_state()):
        assert H.bit_equal(a, b)


@pytest.mark.parametrize("flags", [H.STABLE_FLAGS, H.ALL_FLAGS, H.NO_FLAGS], ids=["stable", "all", "none"])
def test_step_host_contract_pinned_graph(built, flags):
    """The same contract on PAGE-LOCKED host arrays runs as one CUDA graph (copies in, substep, copies
    out, the position download overlapped with the tail passes): bit-identical, including across a
    table overflow inside the graph (K grows -> plain replay) and a change of host arrays."""
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], flags, debug=False)

    def pinned(arrays):
        out = [a.copy() for a in arrays]
        for a in out:
            sol.host_register(a)
        return out

    import ctypes as C
    sol.lib.pbf_debug_set_capacity.restype = C.c_int
    sol.lib.pbf_debug_set_capacity.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    assert sol.lib.pbf_debug_set_capacity(sol.ctx, 8, 1 << 16) == 0   # the first substep overflows the neighbour table
    st = pinned(orc.get_state())
    for step in range(4):
        sol.step_host(st, 1)
        orc.step(1)
        for a, b in zip(st, orc.get_state()):
            assert H.bit_equal(a, b), f"step {step}"
    st2 = pinned(st)   # other arrays: re-capture
    for step in range(2):
        sol.step_host(st2, 1)
        orc.step(1)
        for a, b in zip(st2, orc.get_state()):
            assert H.bit_equal(a, b), f"second arrays, step {step}"
    assert np.float32(sol.time) == np.float32(orc.time)
    # and the device-resident state is the same thing
    assert H.compare_state_bits(sol, orc) == []
    for a in st + st2:
        sol.host_unregister(a)


def test_snapshot_is_the_state_of_its_moment(built):
    """pbf_snapshot_begin / _wait (asynchronous frame output): the positions a snapshot returns are
    those of the substep it was begun after, whatever ran in between; two slots are independent."""
    sol, orc, params = H.make_pair(SMALL, H.STABLE_FLAGS, debug=False)
    sol.step(3)
    orc.step(3)
    want0 = [a.copy() for a in orc.get_state()[:3]]
    t0 = sol.time
    sol.snapshot_begin(0)
    sol.step(2)
    orc.step(2)
    want1 = [a.copy() for a in orc.get_state()[:3]]
    sol.snapshot_begin(1)
    sol.step(4)
    got0, time0 = sol.snapshot_wait(0)
    got1, time1 = sol.snapshot_wait(1)
    for a, b in zip(got0, want0):
        assert H.bit_equal(np.array(a), b)
    for a, b in zip(got1, want1):
        assert H.bit_equal(np.array(a), b)
    assert np.float32(time0) == np.float32(t0) and time1 > time0
    with pytest.raises(Exception):
        sol.snapshot_wait(0)          # nothing pending in the slot any more
    sol.snapshot_begin(0)
    with pytest.raises(Exception):
        sol.snapshot_begin(0)         # still pending
    sol.snapshot_wait(0)


def test_strict_exact_domain_is_reported(built):
    """pbf_strict_exact names the two parameter sets where STRICT is not bit-pinned."""
    sol, orc, params = H.make_pair(SMALL, H.ALL_FLAGS, debug=False)
    assert sol.strict_exact() == (True, None)
    p = params.copy()
    p.scorr_n = 5
    sol.set_params(p)
    ok, why = sol.strict_exact()
    assert not ok and "pow" in why
    p = params.copy()
    p.solver_iterations = 0
    sol.set_params(p)
    ok, why = sol.strict_exact()
    assert not ok and "solver_iterations" in why
    p.enable_xsph = p.enable_vorticity = 0
    sol.set_params(p)
    assert sol.strict_exact() == (True, None)


def test_empty_state_advances_time(built):
    from fluidsimulator_b200.capi import Solver
    params, planes, st = scenes.load_scene(SMALL)
    sol = Solver(0, 0)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.upload([np.zeros(0, np.float32)] * 6)
    sol.step(3)
    t = np.float32(0)
    for _ in range(3):
        t = np.float32(t + np.float32(params.dt))
    assert np.float32(sol.time) == t


# ---- against the committed golden fixtures (outputs of the unmodified reference) -----------
import golden_util as G  # noqa: E402

FLAGSETS = {"none": H.NO_FLAGS, "stable": H.STABLE_FLAGS, "all": H.ALL_FLAGS}


def _solver(scene, flags, iterations=None, mode=PBF_MODE_STRICT):
    from fluidsimulator_b200.capi import Solver
    params, planes, state = scenes.load_scene(scene)
    params = H.configure(params, flags, iterations=iterations)
    sol = Solver(0, len(state[0]), mode)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.debug_enable(True)
    sol.upload(state)
    return sol


@pytest.mark.parametrize("flagname", list(FLAGSETS))
def test_strict_matches_golden_small(built, flagname):
    flags = FLAGSETS[flagname]
    gold = G.load_small(flagname)
    sol = _solver(scenes.small_block(10), flags)
    done = 0
    for step in sorted(gold):
        sol.step(step - done)
        done = step
        assert G.mismatches(G.snapshot_of(sol, flags, True), gold[step]) == [], f"step {step}"
        assert np.float32(sol.time) == gold[step]["time"]


@pytest.mark.parametrize("run", ["fluid_large:stable", "fluid_large:all", "fluid_large:all:iters8",
                                 "fluid_large:stable:iters2", "fluid_large:all:blowup", "fluid_double_side:all",
                                 "fluid_double_dem:all", "fluid_xlarge:stable"])
def test_strict_matches_golden_digests(built, run):
    """sha256 of every array the reference leaves behind (state, grid tables, neighbour lists,
    scratch) on the scenes BASELINE.json names, generated from the unmodified reference by
    tests/golden/make_golden.py — no CPU replay at test time."""
    gold = G.digests()["runs"][run]
    parts = run.split(":")
    flags = FLAGSETS[parts[1]]
    iters = next((int(p[5:]) for p in parts[2:] if p.startswith("iters")), None)
    sol = _solver(scenes.SCENES[parts[0]], flags, iters)
    done = 0
    for step in sorted(int(s) for s in gold):
        sol.step(step - done)
        done = step
        assert G.digest_mismatches(G.snapshot_of(sol, flags, True), gold[str(step)]) == [], f"step {step}"


def test_neighbor_capacity_growth_is_transparent(built):
    """A batch that overflows the neighbour table is re-run after growing it; results are unchanged."""
    import ctypes as C
    flags = H.STABLE_FLAGS
    gold = G.load_small("stable")
    sol = _solver(scenes.small_block(10), flags)
    sol.lib.pbf_debug_set_capacity.restype = C.c_int
    sol.lib.pbf_debug_set_capacity.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    assert sol.lib.pbf_debug_set_capacity(sol.ctx, 8, 64) == 0   # K = 8 neighbours, 64 grid cells
    sol.step(5)
    assert G.mismatches(G.snapshot_of(sol, flags, True), gold[5]) == []


def test_fluid_million_properties(built):
    """Full-size scene (BASELINE.json size): size-independent properties instead of a CPU replay.
    Sortedness and completeness of the grid tables, symmetry of the neighbour relation, and
    equality of the STRICT result regardless of batching / graph replay."""
    from fluidsimulator_b200.capi import Solver
    params, planes, state = scenes.load_scene(scenes.SCENES["fluid_million"])
    params = H.configure(params, H.STABLE_FLAGS)
    n = len(state[0])
    outs = []
    for graph, chunks in ((True, [6]), (False, [1, 2, 3])):
        sol = Solver(0, n)
        sol.set_params(params)
        sol.set_planes(planes)
        sol.set_graph(graph)
        sol.upload(state)
        for c in chunks:
            sol.step(c)
        outs.append(sol.download())
        if graph:
            g = sol.debug_grid()
            key = (g["entry_cx"].astype(np.int64) << 42) + (g["entry_cy"].astype(np.int64) << 21) + g["entry_cz"]
            assert np.all(np.diff(key) >= 0)                                  # sorted by (x, y, z)
            same = np.diff(key) == 0
            assert np.all(np.diff(g["entry_particle"].astype(np.int64))[same] > 0)   # ties by ascending id
            assert np.array_equal(np.sort(g["entry_particle"]), np.arange(n, dtype=np.int32))  # a permutation
            assert g["cell_start"][0] == 0 and g["cell_end"][-1] == n
            assert np.array_equal(g["cell_start"][1:], g["cell_end"][:-1])    # cells tile the sorted order
            prefix, idx = sol.debug_neighbors()
            counts = np.diff(np.concatenate([[0], prefix]))
            owner = np.repeat(np.arange(n, dtype=np.int64), counts)
            fwd = np.sort(owner * n + idx)
            bwd = np.sort(idx.astype(np.int64) * n + owner)
            assert np.array_equal(fwd, bwd)                                   # j in N(i) <=> i in N(j)
    for a, b in zip(*outs):
        assert H.bit_equal(a, b)


def test_strict_follows_the_reference_through_its_blow_up(built):
    """With vorticity on the reference diverges: within ~15 substeps particles fly
    kilometres away and the bounding box of occupied cells no longer fits a dense table.  The
    backend switches to the sparse (hashed) cell table and stays bit-identical — state, grid
    tables in reference order, neighbour lists."""
    import ctypes as C
    sol, orc, params = H.make_pair(scenes.SCENES["fluid_large"], H.ALL_FLAGS)
    sol.lib.pbf_debug_grid_is_sparse.restype = C.c_int
    sol.lib.pbf_debug_grid_is_sparse.argtypes = [C.c_void_p]
    seen_sparse = False
    for chunk in range(8):
        sol.step(5)
        orc.step(5)
        assert H.compare_state_bits(sol, orc) == [], f"substep {5 * chunk + 5}"
        if sol.lib.pbf_debug_grid_is_sparse(sol.ctx) == 1 and not seen_sparse:
            seen_sparse = True
            assert H.compare_integers(sol, orc) == [], f"sparse tables at substep {5 * chunk + 5}"
    assert seen_sparse, "the scene was expected to outgrow the dense table"
    assert H.compare_integers(sol, orc) == []
    assert H.compare_scratch_bits(sol, orc, params) == []
