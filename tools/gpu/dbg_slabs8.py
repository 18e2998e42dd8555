"""Debug aid: fluid_million on 8 virtual slabs (one GPU, one process) through the bench's batch
sequence, printing owned counts, cuts, re-balances and retried batches per batch."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from fluidsimulator_b200.capi import PBF_MODE_STRICT, SlabGroup

nslabs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
params, planes, state = B.load_scene("fluid_million", B.FLAGSETS["stable"], 4)
grp = SlabGroup([0] * nslabs, params, planes, PBF_MODE_STRICT)
grp.upload(state)
done = 0
for batch in (65, 35, 10) + (25,) * 6:
    t0 = time.perf_counter()
    grp.step(batch)
    dt = (time.perf_counter() - t0) * 1e3 / batch
    done += batch
    s0 = grp.slabs[0]
    print(f"after {done:4d}: {dt:.3f} ms/substep owned {[s.owned() for s in grp.slabs]} cuts {[s.cuts() for s in grp.slabs]} "
          f"rebalanced {s0.rebalance_count()} retried {s0.batches_retried()} stats {s0.slab_stats()}", flush=True)
