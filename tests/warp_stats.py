"""Warp-level utilisation of the list-driven passes and of the neighbour kernel's test loop, counted on
the CPU from the reference's own grid and neighbour lists (test-infrastructure oracle; not a product
path): python tests/warp_stats.py fluid_xlarge 250.  Numbers quoted in DESIGN.md §9."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
from fluidsimulator_b200 import scenes
from oracle.oracle_api import Oracle
import helpers as H
scene, steps = sys.argv[1], int(sys.argv[2])
params, planes, state = scenes.load_scene(scenes.SCENES[scene])
params = H.configure(params, H.STABLE_FLAGS)
orc = Oracle("reference"); orc.set_params(params); orc.set_planes(planes); orc.set_state(state)
t=time.time(); orc.step(steps); print("oracle s", round(time.time()-t,1))
g = orc.grid(); prefix, idx = orc.neighbors()
n = len(prefix)
counts_orig = np.diff(np.concatenate([[0], prefix]))          # per original particle
order = g["entry_particle"]                                     # sorted slot -> particle
cnt = counts_orig[order]                                        # per sorted slot
print("particles", n, "avg nbrs", cnt.mean())
# solver passes: warp = 32 consecutive sorted slots; pairs per lane = ceil(cnt/2)
pad = (-n) % 32
pairs = np.concatenate([(cnt+1)//2, np.zeros(pad, int)]).reshape(-1,32)
print("solver: lane utilisation (sum pairs / 32*max pairs) =", pairs.sum()/ (32*pairs.max(axis=1)).sum())
# neighbour kernel: per lane, per stencil cell, candidate count
cs, ce, cxyz = g["cell_start"], g["cell_end"], g["cell_xyz"]
occ = (ce-cs)
key = {tuple(c): k for k,c in enumerate(map(tuple,cxyz))}
cell_of_slot = np.repeat(np.arange(len(cs)), occ)
ncell = len(cs)
# occupancy of the 27 stencil cells per cell
st = np.zeros((ncell,27), int)
offs = [(dx,dy,dz) for dz in (-1,0,1) for dy in (-1,0,1) for dx in (-1,0,1)]
import itertools
cx = cxyz
lut = {}
for k,(x,y,z) in enumerate(cx): lut[(x,y,z)] = k
for k,(x,y,z) in enumerate(cx):
    for o,(dx,dy,dz) in enumerate(offs):
        j = lut.get((x+dx,y+dy,z+dz))
        if j is not None: st[k,o] = occ[j]
lane_c = st[cell_of_slot]                                       # [n,27] candidates per lane per stencil cell
lane_c = np.concatenate([lane_c, np.zeros((pad,27),int)]).reshape(-1,32,27)
pairs_c = (lane_c+1)//2
warp_iters = pairs_c.max(axis=1).sum(axis=1)                    # sum over cells of max over lanes
lane_iters = pairs_c.sum(axis=2)                                # per lane total
print("neighbour test loop: candidates/particle", lane_c.sum()/n, " pair-iterations per warp", warp_iters.mean(),
      " mean per lane", lane_iters.sum()/n, " utilisation", lane_iters.sum()/(32*warp_iters.sum()))
print("  if the 3 cells of a row were one loop:", (pairs_c.reshape(-1,32,9,3).sum(axis=3).max(axis=1).sum(axis=1)).mean())
print("  if all 27 cells were one loop:", lane_iters.reshape(-1,32).max(axis=1).mean())
