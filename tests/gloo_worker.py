"""Worker of test_slab_plan.py::test_two_gloo_ranks_split_and_gather (2 CPU ranks, gloo)."""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from fluidsimulator_b200 import capi, multigpu, scenes  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cpu = torch.device("cpu")
    payload = bytes(range(128)) if rank == 0 else None
    got = multigpu.broadcast_bytes(dist, payload, 128, cpu)
    assert got == bytes(range(128))

    params, planes, state = scenes.load_scene(scenes.small_block(14))
    n = len(state[0])
    cuts = capi.slab_plan(state[0], float(params.h), world)       # same arrays => same cuts on every rank
    all_cuts = [None] * world
    dist.all_gather_object(all_cuts, cuts.tolist())
    assert all(c == all_cuts[0] for c in all_cuts)
    mask = multigpu.owned_mask(state[0], float(params.h), cuts, rank)
    gid = np.nonzero(mask)[0].astype(np.int64)                    # ascending global ids
    mine = [a[mask].copy() for a in state]
    mine[4] = mine[4] - np.float32(1.0)                           # some per-rank "work"
    full = multigpu.gather_global(dist, gid, mine, n, cpu)
    if rank == 0:
        expect = [a.copy() for a in state]
        expect[4] = expect[4] - np.float32(1.0)
        for a, b in zip(full, expect):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        # the multi-GPU bench line's parity witness hashes exactly this gathered state
        import bench
        assert bench.combined16(full) == bench.combined16(expect)
        step, digest = bench.golden_digest("fluid_million", "stable", 4, "strict")
        assert step == 65 and len(digest) == 16
        assert bench.golden_digest("fluid_million", "stable", 2, "strict") is None   # other workloads: no digest
        print("gloo slab plumbing ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
