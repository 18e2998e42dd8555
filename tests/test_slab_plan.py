"""Host-side logic of the x-slab decomposition (no GPU): the cut planner of the C ABI
(pbf_slab_plan) and the multi-process plumbing of fluidsimulator_b200/multigpu.py, the latter
with two gloo ranks on CPU."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from fluidsimulator_b200 import capi, multigpu, scenes

ROOT = Path(__file__).resolve().parent.parent
I32_MIN, I32_MAX = np.iinfo(np.int32).min, np.iinfo(np.int32).max


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_plan_partitions_fluid_million(built, nranks):
    params, planes, state = scenes.load_scene(scenes.SCENES["fluid_million"])
    px, h = state[0], float(params.h)
    cuts = capi.slab_plan(px, h, nranks)
    assert cuts[0] == I32_MIN and cuts[-1] == I32_MAX and len(cuts) == nranks + 1
    inner = cuts[1:-1].astype(np.int64)
    cx = multigpu.cell_x(px, h)
    if nranks > 1:
        assert np.all(np.diff(inner) >= 2)                       # every slab >= 2 cell layers wide
        assert inner[0] - cx.min() >= 2 and cx.max() + 1 - inner[-1] >= 2
    masks = [multigpu.owned_mask(px, h, cuts, r) for r in range(nranks)]
    total = np.sum(masks, axis=0)
    assert np.all(total == 1)                                     # a partition: every particle owned once
    counts = np.array([m.sum() for m in masks])
    # cuts sit on cell-layer boundaries (one layer of this lattice = 1-2 % of the particles)
    assert counts.max() - counts.min() <= 2 * np.bincount(cx - cx.min()).max()


def test_plan_rejects_too_many_slabs(built):
    px = np.linspace(0.0, 0.35, 50, dtype=np.float32)             # 4 cell layers at h = 0.1
    assert len(capi.slab_plan(px, 0.1, 2)) == 3
    with pytest.raises(capi.PbfError, match="can not be split"):
        capi.slab_plan(px, 0.1, 3)


def test_plan_uses_the_reference_cell_expression(built):
    """Cell of x is floor(x * (1.0f / h)) in float32 (core.cpp:28-34), not floor(x / h): the two
    differ at cell boundaries, e.g. x = 0.3f, h = 0.1f."""
    h = np.float32(0.1)
    x = np.float32(0.3)
    assert int(np.floor(x * (np.float32(1) / h))) == 3 and int(np.floor(np.float32(x / h))) == 3
    px = np.array([0.05, 0.15, 0.25, 0.3, 0.45, 0.55, 0.65, 0.75], dtype=np.float32)
    cuts = capi.slab_plan(px, float(h), 2)
    cx = multigpu.cell_x(px, float(h))
    assert cx.tolist() == [0, 1, 2, 3, 4, 5, 6, 7]
    assert cuts[1] == 4


def _slab_costs(hist, cuts_rel, w):
    """Modelled cost per slab (include/pbf_b200.h: owned + w * first ghost layers + w/4 * second),
    in the planner's integer units of 1/256 particle."""
    w1 = int(round(w * 256))
    w2 = w1 // 4
    L, R = len(hist), len(cuts_rel) - 1
    at = lambda l: int(hist[l]) if 0 <= l < L else 0
    out = []
    for r in range(R):
        a, b = cuts_rel[r], cuts_rel[r + 1]
        c = 256 * int(np.sum(hist[a:b]))
        if r > 0:
            c += w1 * at(a - 1) + w2 * at(a - 2)
        if r < R - 1:
            c += w1 * at(b) + w2 * at(b + 1)
        out.append(c)
    return out


def _brute_force_best(hist, nranks, w):
    """Minimum over every valid plan (>= 2 layers per slab) of the busiest slab's cost."""
    import itertools
    L = len(hist)
    best = None
    for inner in itertools.combinations(range(2, L - 1), nranks - 1):
        cuts = (0,) + inner + (L,)
        if min(np.diff(cuts)) < 2:
            continue
        c = max(_slab_costs(hist, cuts, w))
        best = c if best is None else min(best, c)
    return best


@pytest.mark.parametrize("w", [0.0, 0.25, 0.5, 1.0])
def test_weighted_plan_is_optimal_on_small_histograms(built, w):
    """The planner minimises the modelled cost of the busiest slab: checked against a brute-force
    search over every valid plan on random histograms (uniform, piled-up, with empty layers)."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        L = int(rng.integers(4, 15))
        nranks = int(rng.integers(2, min(4, L // 2) + 1))
        kind = trial % 3
        if kind == 0:
            hist = rng.integers(50, 60, L)
        elif kind == 1:
            hist = (1000 * np.exp(-np.arange(L) / 3.0)).astype(np.int64) + rng.integers(0, 5, L)
        else:
            hist = rng.integers(0, 100, L) * rng.integers(0, 2, L)
            hist[0] = hist[-1] = 1 + hist[0]
        first = int(rng.integers(-50, 50))
        cuts = capi.slab_plan_hist(hist, first, nranks, w)
        assert cuts[0] == I32_MIN and cuts[-1] == I32_MAX
        rel = [0] + [int(c) - first for c in cuts[1:-1]] + [L]
        assert min(np.diff(rel)) >= 2, (hist, rel)
        assert max(_slab_costs(hist, rel, w)) == _brute_force_best(hist, nranks, w), (hist, rel, w)


def test_weighted_plan_gives_interior_slabs_fewer_particles(built):
    """A uniform block of 58 layers over 8 slabs (7.25 layers each, about fluid_million on 8 GPUs):
    the two 8-layer slabs go to the ends of the scene, which have one ghost side only, and the
    modelled busiest slab is cheaper than with equal owned counts.  A ghost side weighs about one
    cell layer, so this is all the model can change: one layer per slab."""
    hist = np.full(58, 18_000, dtype=np.uint64)
    equal = capi.slab_plan_hist(hist, 10, 8, 0.0)
    weighted = capi.slab_plan_hist(hist, 10, 8, 0.5)
    rel_e = [0] + [int(c) - 10 for c in equal[1:-1]] + [58]
    rel_w = [0] + [int(c) - 10 for c in weighted[1:-1]] + [58]
    assert np.diff(rel_w).tolist() == [8, 7, 7, 7, 7, 7, 7, 8]
    assert sorted(np.diff(rel_e).tolist()) == [7] * 6 + [8] * 2       # w = 0: equal owned counts
    assert max(_slab_costs(hist, rel_w, 0.5)) <= max(_slab_costs(hist, rel_e, 0.5))


def test_plan_hist_rejects_bad_arguments(built):
    with pytest.raises(capi.PbfError):
        capi.slab_plan_hist(np.ones(8, dtype=np.uint64), 0, 2, 1.5)
    with pytest.raises(capi.PbfError, match="can not be split"):
        capi.slab_plan_hist(np.ones(3, dtype=np.uint64), 0, 2, 0.5)
    with pytest.raises(capi.PbfError, match="can not be split"):
        capi.slab_plan_hist(np.zeros(8, dtype=np.uint64), 0, 2, 0.5)   # no particles
    assert capi.slab_plan_hist(np.zeros(0, dtype=np.uint64), 0, 1, 0.5).tolist() == [I32_MIN, I32_MAX]


def test_two_gloo_ranks_split_and_gather():
    """world_size 2 on CPU: unique-id broadcast, ownership split and gather back to original order."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "gloo_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "gloo slab plumbing ok" in out.stdout
