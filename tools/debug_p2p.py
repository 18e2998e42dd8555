import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from fluidsimulator_b200 import scenes
from fluidsimulator_b200.capi import SlabGroup, Solver, PbfError
import helpers as H

graph = (sys.argv[1] == "graph") if len(sys.argv) > 1 else True
flags = H.ALL_FLAGS if (len(sys.argv) > 2 and sys.argv[2] == "all") else H.STABLE_FLAGS
params, planes, state = scenes.load_scene(scenes.SCENES["fluid_large"])
params = H.configure(params, flags)
state = [a.copy() for a in state]; state[3][:] = 1.5
sol = Solver(0, len(state[0])); sol.set_params(params); sol.set_planes(planes); sol.upload(state)
grp = SlabGroup([0, 0], params, planes, p2p=True)
for s in grp.slabs: s.set_graph(graph)
grp.upload(state)
for step in range(1, 12):
    try:
        grp.step(1)
    except PbfError as e:
        print("step", step, "ERROR", e); break
    sol.step(1)
    a, b = grp.download(), sol.download()
    bad = [k for k in range(6) if not np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))]
    print("step", step, "owned", grp.owned(), "bad", bad, [s.slab_stats() for s in grp.slabs], flush=True)
