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


def test_two_gloo_ranks_split_and_gather():
    """world_size 2 on CPU: unique-id broadcast, ownership split and gather back to original order."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "gloo_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "gloo slab plumbing ok" in out.stdout
