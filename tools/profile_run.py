"""Driver for ncu: advances the scene untimed, then runs a few un-graphed substeps between
cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Not a benchmark."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from fluidsimulator_b200.capi import Solver, PBF_MODE_STRICT, PBF_MODE_FAST

scene = sys.argv[1] if len(sys.argv) > 1 else "fluid_million"
flagname = sys.argv[2] if len(sys.argv) > 2 else "stable"
presteps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
mode = PBF_MODE_FAST if (len(sys.argv) > 5 and sys.argv[5] == "fast") else PBF_MODE_STRICT
params, planes, state = bench.load_scene(scene, bench.FLAGSETS[flagname], 4)
n = len(state[0])
sol = Solver(0, n, mode); sol.set_params(params); sol.set_planes(planes); sol.upload(state)
if presteps:
    sol.step(presteps)
sol.set_graph(False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sol.step(steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", steps, "substeps of", scene, flagname, "after", presteps, "presteps; avg nbrs", sol.debug_sizes()[1] / n)
