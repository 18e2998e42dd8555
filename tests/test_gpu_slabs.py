"""x-slab decomposition on ONE GPU ("virtual ranks"): several slab contexts on cuda:0
linked by the in-process transport must reproduce the single-context result BIT FOR BIT in STRICT
mode — through migration, ghost build, per-iteration halo refresh, capacity growth and multi-hop
migration.  The single-context result is itself pinned to the CPU oracle by test_gpu_parity.py;
one test here closes the loop against the oracle directly."""
import ctypes as C

import numpy as np
import pytest

from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import PBF_MODE_FAST, PBF_MODE_STRICT, SlabGroup, Solver

import helpers as H

pytestmark = pytest.mark.gpu


def _single(params, planes, state, mode=PBF_MODE_STRICT):
    sol = Solver(0, len(state[0]), mode)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.upload(state)
    return sol


def _scene(scene, flags, iterations=None, vx=None):
    params, planes, state = scenes.load_scene(scene)
    params = H.configure(params, flags, iterations=iterations)
    if vx is not None:
        state = [a.copy() for a in state]
        state[3][:] = vx
    return params, planes, state


def _assert_same(group, sol, what=""):
    names = ["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z"]
    bad = [n for n, a, b in zip(names, group.download(), sol.download()) if not H.bit_equal(a, b)]
    assert bad == [], f"{what}: slabs differ from one GPU in {bad}"


@pytest.mark.parametrize("nslabs", [2, 3, 5])
@pytest.mark.parametrize("flags", [H.NO_FLAGS, H.STABLE_FLAGS, H.ALL_FLAGS], ids=["none", "stable", "all"])
def test_virtual_slabs_bit_exact(built, nslabs, flags):
    params, planes, state = _scene(scenes.SCENES["fluid_large"], flags)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * nslabs, params, planes)
    grp.upload(state)
    assert sum(grp.owned()) == len(state[0])
    for step in range(1, 7):
        grp.step(1)
        sol.step(1)
        _assert_same(grp, sol, f"step {step}")
    grp.step(6)
    sol.step(6)
    _assert_same(grp, sol, "batch of 6")
    grp.close()


@pytest.mark.parametrize("nslabs", [2, 4])
def test_virtual_slabs_direct_peer_stores(built, nslabs):
    """The fused transport: pack kernels store straight into the neighbour's window and an exchange
    is one flag kernel (pbf_slab_set_p2p).  Same bits as one GPU, stepping one substep at a time
    (plain launches) and in batches (CUDA-graph replay with the flag kernels inside)."""
    for flags, batches in ((H.STABLE_FLAGS, 3), (H.ALL_FLAGS, 1)):   # with vorticity the reference blows up soon
        params, planes, state = _scene(scenes.SCENES["fluid_large"], flags, vx=1.5)
        sol = _single(params, planes, state)
        grp = SlabGroup([0] * nslabs, params, planes, p2p=True)
        grp.upload(state)
        for step in range(1, 4):
            grp.step(1)
            sol.step(1)
            _assert_same(grp, sol, f"p2p step {step}")
        for _ in range(batches):
            grp.step(5)
            sol.step(5)
            _assert_same(grp, sol, "p2p batch of 5")
        assert sum(grp.owned()) == len(state[0])
        grp.close()


def test_virtual_slabs_direct_peer_stores_growth(built):
    """Capacity growth re-allocates the peer windows on every slab at once."""
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 3, params, planes, p2p=True)
    for s in grp.slabs:
        s.lib.pbf_debug_set_slab_capacity.restype = C.c_int
        s.lib.pbf_debug_set_slab_capacity.argtypes = [C.c_void_p, C.c_int, C.c_int]
        assert s.lib.pbf_debug_set_slab_capacity(s.ctx, 4, 64) == 0
    grp.upload(state)
    grp.step(3)
    sol.step(3)
    _assert_same(grp, sol, "p2p after growth")
    grp.close()


def test_virtual_slabs_migration(built):
    """The block drifts in +x at 3 m/s (a quarter cell per substep): particles cross the cuts all
    the time, pile up on the +x wall, and owned counts change."""
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, vx=3.0)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 4, params, planes)
    grp.upload(state)
    before = grp.owned()
    grp.step(25)
    sol.step(25)
    _assert_same(grp, sol, "25 substeps with drift")
    assert grp.owned() != before and sum(grp.owned()) == len(state[0])
    stats = grp.slabs[1].slab_stats()
    assert stats["exchanges"] > 0 and stats["ghosts"] > 0
    grp.close()


def test_virtual_slabs_unequal_slabs_share_capacities(built):
    """fluid_xlarge: slabs hold different particle counts, large enough that the capacity
    heuristics leave their floor values; the message capacities are wire format and must agree."""
    params, planes, state = _scene(scenes.SCENES["fluid_xlarge"], H.STABLE_FLAGS)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 3, params, planes)
    grp.upload(state)
    assert len(set(grp.owned())) > 1
    grp.step(4)
    sol.step(4)
    _assert_same(grp, sol, "fluid_xlarge, 3 slabs")
    grp.close()


def test_virtual_slabs_against_oracle(built):
    params, planes, state = _scene(scenes.small_block(12), H.ALL_FLAGS)
    from oracle.oracle_api import Oracle, best_kind
    orc = Oracle(best_kind())
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    grp = SlabGroup([0] * 2, params, planes)
    grp.upload(state)
    grp.step(8)
    orc.step(8)
    for a, b in zip(grp.download(), orc.get_state()):
        assert H.bit_equal(a, b)
    grp.close()


def test_virtual_slabs_capacity_growth_and_hops(built):
    """Tiny message capacities force the overflow -> grow -> replay path; a few very fast particles
    cross more than one slab per substep and force a second migration hop.  Results unchanged."""
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS)
    state = [a.copy() for a in state]
    state[3][::97] = 45.0   # 0.375 m per substep = almost 4 cells: more than one 2-cell slab
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 6, params, planes)
    for s in grp.slabs:
        s.lib.pbf_debug_set_slab_capacity.restype = C.c_int
        s.lib.pbf_debug_set_slab_capacity.argtypes = [C.c_void_p, C.c_int, C.c_int]
        assert s.lib.pbf_debug_set_slab_capacity(s.ctx, 4, 64) == 0
    grp.upload(state)
    grp.step(4)
    sol.step(4)
    _assert_same(grp, sol, "after growth")
    assert grp.slabs[0].slab_stats()["hops"] >= 2
    grp.close()


def test_virtual_slabs_iterations_and_fast_mode(built):
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.ALL_FLAGS, iterations=2)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 3, params, planes)
    grp.upload(state)
    grp.step(3)
    sol.step(3)
    _assert_same(grp, sol, "2 iterations")
    grp.close()
    # FAST mode: same kernels on both sides, so slabs still match one GPU exactly
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS)
    sol = _single(params, planes, state, PBF_MODE_FAST)
    grp = SlabGroup([0] * 3, params, planes, PBF_MODE_FAST)
    grp.upload(state)
    grp.step(5)
    sol.step(5)
    _assert_same(grp, sol, "fast mode")
    grp.close()


@pytest.mark.parametrize("p2p", ["1", "0"], ids=["peer-stores", "nccl-messages"])
def test_nccl_slabs_two_gpus(built, p2p):
    """Two processes, one GPU each (skipped on a 1-GPU box): halos by direct stores into the
    neighbour's cudaIpc window (default) and by ncclSend/ncclRecv messages (PBF_SLAB_P2P=0); the
    second half of the run replays the captured CUDA graph with the exchanges inside."""
    import os
    import subprocess
    import sys
    from fluidsimulator_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = H.scenes.__file__.rsplit("/fluidsimulator_b200/", 1)[0]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517" if p2p == "1" else "29518",
           "-m", "fluidsimulator_b200.multigpu", "--check", "--scene", "fluid_large", "--steps", "10"]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, PBF_SLAB_P2P=p2p))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "slab check ok" in out.stdout


def test_virtual_slabs_rebalance(built):
    """Cuts re-planned in the middle of a run (pbf_slab_set_cuts): whole cell layers change owner
    during the next substep, several hops for some, message capacities grow — same bits as one GPU."""
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, vx=2.0)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 4, params, planes)
    for s in grp.slabs:
        s.set_rebalance(0.0)            # the automatic policy would have moved the cuts already
    grp.upload(state)
    grp.step(20)
    sol.step(20)
    before = [s.cuts() for s in grp.slabs]
    cuts = grp.rebalance(float(params.h))
    assert [s.cuts() for s in grp.slabs] != before and len(cuts) == 5
    grp.step(6)
    sol.step(6)
    _assert_same(grp, sol, "after rebalancing")
    counts = grp.owned()
    assert max(counts) - min(counts) < 0.2 * sum(counts)
    grp.close()


def test_virtual_slabs_fluid_million(built):
    """BASELINE.json's headline scene at full size: two slabs == one context, bit for bit
    (one context is tied to the oracle through properties in test_gpu_parity.py)."""
    params, planes, state = _scene(scenes.SCENES["fluid_million"], H.STABLE_FLAGS)
    sol = _single(params, planes, state)
    grp = SlabGroup([0, 0], params, planes)
    grp.upload(state)
    assert abs(grp.owned()[0] - grp.owned()[1]) <= 20000      # cuts on cell layers of 10 000 particles
    grp.step(1)
    grp.step(3)
    sol.step(4)
    _assert_same(grp, sol, "fluid_million, 2 slabs, 4 substeps")
    grp.close()


def test_virtual_slabs_automatic_rebalancing(built):
    """The block drifts to +x: the left slabs empty, the right ones fill up.  With the automatic
    policy the cuts follow (decided on every rank from the max-reduced status block, planned on an
    all-reduced x-layer histogram); the result stays bit-identical to one GPU."""
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, vx=3.0)
    sol = _single(params, planes, state)
    counts = {}
    for policy, threshold in (("off", 0.0), ("auto", 1.15)):
        grp = SlabGroup([0] * 4, params, planes)
        for s in grp.slabs:
            s.set_rebalance(threshold)
        grp.upload(state)
        for _ in range(8):
            grp.step(5)
        counts[policy] = (grp.owned(), [s.rebalance_count() for s in grp.slabs])
        if policy == "off":
            sol.step(40)
        _assert_same(grp, sol, f"rebalancing {policy}")
        grp.close()
    assert counts["off"][1] == [0, 0, 0, 0]
    assert len(set(counts["auto"][1])) == 1 and counts["auto"][1][0] >= 1      # every slab took the same decisions
    spread = lambda c: max(c) / (sum(c) / len(c))
    assert spread(counts["auto"][0]) < spread(counts["off"][0])
    assert spread(counts["auto"][0]) < 1.5


def test_virtual_slabs_replan_inside_one_long_call(built):
    """ONE pbf_step(80) call on a drifting block: the call is cut into batches of <= 25 substeps and
    the cuts are re-planned between them (round 1 only re-planned between calls, which let the end
    slabs of fluid_million triple during a 100-substep call).  Same bits as one GPU, the default
    policy, no table had to grow for it twice."""
    params, planes, state = _scene(scenes.SCENES["fluid_large"], H.STABLE_FLAGS, vx=3.0)
    sol = _single(params, planes, state)
    grp = SlabGroup([0] * 4, params, planes)
    grp.upload(state)
    grp.step(80)
    sol.step(80)
    _assert_same(grp, sol, "one call of 80 substeps")
    replans = [s.rebalance_count() for s in grp.slabs]
    assert len(set(replans)) == 1 and replans[0] >= 1, replans     # re-planned inside the single call
    assert sum(grp.owned()) == len(state[0])
    grp.close()
