"""GPU exploration: ms/substep and neighbour statistics along a trajectory (not a test)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from fluidsimulator_b200.capi import Solver, PBF_MODE_STRICT, PBF_MODE_FAST

scene = sys.argv[1] if len(sys.argv) > 1 else "fluid_million"
flagname = sys.argv[2] if len(sys.argv) > 2 else "stable"
total = int(sys.argv[3]) if len(sys.argv) > 3 else 600
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 50
for mode, mname in ((PBF_MODE_STRICT, "strict"), (PBF_MODE_FAST, "fast")):
    params, planes, state = bench.load_scene(scene, bench.FLAGSETS[flagname], 4)
    n = len(state[0])
    stream = torch.cuda.Stream()
    sol = Solver(0, n, mode); sol.set_params(params); sol.set_planes(planes); sol.set_stream(stream.cuda_stream); sol.upload(state)
    with torch.cuda.stream(stream):
        sol.step(3)
        done = 3
        while done < total:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); sol.step(chunk); e1.record(stream); torch.cuda.synchronize()
            done += chunk
            ms = e0.elapsed_time(e1) / chunk
            nn = sol.debug_sizes()[1]
            print(f"{scene} {flagname} {mname} step {done}: {ms:.3f} ms/substep = {n/ms*1e3:.3e} p-substeps/s, avg nbrs {nn/n:.1f}, K={sol.lib.pbf_launch_count(sol.ctx)}", flush=True)
        sol.profile_enable(True); sol.profile_reset(); sol.step(20); torch.cuda.synchronize()
        prof = sol.profile()
        print({k: round(v["ms"]/20, 4) for k, v in prof.items() if v["launches"]})
