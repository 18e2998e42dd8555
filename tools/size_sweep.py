"""ms/substep of one GPU over scene sizes (fixed-overhead fit; not a benchmark)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from fluidsimulator_b200.capi import Solver, PBF_MODE_STRICT
for scene in ["small_12", "fluid_large", "fluid_double_side", "fluid_xlarge", "fluid_million", "weak_1"]:
    params, planes, state = bench.load_scene(scene, bench.FLAGSETS["stable"], 4)
    n = len(state[0])
    stream = torch.cuda.Stream()
    sol = Solver(0, n, PBF_MODE_STRICT); sol.set_params(params); sol.set_planes(planes); sol.set_stream(stream.cuda_stream); sol.upload(state)
    with torch.cuda.stream(stream):
        sol.step(10)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); sol.step(100); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    print(f"{scene:18s} n={n:8d} {ms*1e3:8.1f} us/substep  {n/ms*1e3:.3e} particle-substeps/s", flush=True)
    sol.close()
