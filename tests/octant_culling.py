"""Octant culling of neighbour candidates, counted on the CPU (oracle state: test infrastructure, not a
product path): candidates per particle with and without skipping the 2x2x2 sub-boxes of the stencil
cells that lie farther than h from the particle, per lane and per WARP (32 consecutive slots of the
reference's sorted order; a warp waits for its slowest lane in every cell).  fluid_xlarge, substep 250.
Numbers quoted in DESIGN.md section 9.      python tests/octant_culling.py   (needs /tmp caches or ~3 min)"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
from fluidsimulator_b200 import scenes
import helpers as H
from oracle.oracle_api import Oracle
params, planes, state = scenes.load_scene(scenes.SCENES['fluid_xlarge'])
params = H.configure(params, H.STABLE_FLAGS)
if not (os.path.exists('/tmp/xlarge250_state.npz') and os.path.exists('/tmp/xlarge250.npz')):
    orc = Oracle("reference"); orc.set_params(params); orc.set_planes(planes); orc.set_state(state)
    orc.step(249)
    np.savez('/tmp/xlarge250_state.npz', *orc.get_state())   # the state substep 250 starts from
    orc.step(1)
    prefix, idx = orc.neighbors()
    np.savez('/tmp/xlarge250.npz', prefix=prefix, idx=idx, **orc.grid())
z=np.load('/tmp/xlarge250_state.npz'); st=[z[f'arr_{k}'] for k in range(6)]
gz=np.load('/tmp/xlarge250.npz')
order=gz['entry_particle']
h=np.float32(params.h); dt=np.float32(params.dt)
g=np.array([params.external_force[k] for k in range(3)],dtype=np.float32)
vel=np.stack(st[3:],1)+g*dt
pred=np.stack(st[:3],1)+vel*dt
u=pred*(np.float32(1)/h)
c=np.floor(u).astype(int); f=(u-c).astype(np.float64)
# consistency with the oracle's grid
assert np.array_equal(c[order,0],gz['entry_cx']) and np.array_equal(c[order,2],gz['entry_cz'])
n=len(c)
c-=c.min(axis=0)-1
dims=c.max(axis=0)+2
o=(f>=0.5).astype(int); oid=o[:,0]*4+o[:,1]*2+o[:,2]
cellid=(c[:,0]*dims[1]+c[:,1])*dims[2]+c[:,2]
occ=np.zeros((dims.prod(),8),int); np.add.at(occ,(cellid,oid),1)
# sorted order, warps of 32
cs=c[order]; fs=f[order]
pad=(-n)%32
W=(n+pad)//32
full=np.zeros((n,27),int); need=np.zeros((n,27),int)
k=0
for dz in (-1,0,1):
  for dy in (-1,0,1):
    for dx in (-1,0,1):
      cid=((cs[:,0]+dx)*dims[1]+cs[:,1]+dy)*dims[2]+cs[:,2]+dz
      oc=occ[cid]                                  # [n,8]
      full[:,k]=oc.sum(axis=1)
      acc=np.zeros(n,int)
      for ox in (0,1):
        lo=dx+0.5*ox; ex=np.maximum(np.maximum(lo-fs[:,0],fs[:,0]-(lo+0.5)),0)
        for oy in (0,1):
          lo=dy+0.5*oy; ey=np.maximum(np.maximum(lo-fs[:,1],fs[:,1]-(lo+0.5)),0)
          for oz in (0,1):
            lo=dz+0.5*oz; ez=np.maximum(np.maximum(lo-fs[:,2],fs[:,2]-(lo+0.5)),0)
            keep=(ex*ex+ey*ey+ez*ez)<1.0
            acc+=np.where(keep,oc[:,ox*4+oy*2+oz],0)
      need[:,k]=acc
      k+=1
def warp_steps(a):
    a=np.concatenate([a,np.zeros((pad,27),int)]).reshape(W,32,27)
    p=(a+1)//2
    return p.max(axis=1).sum(axis=1).mean(), p.sum()/n
print("full:  pair-steps per warp %.1f, per lane %.1f" % warp_steps(full))
print("culled: pair-steps per warp %.1f, per lane %.1f" % warp_steps(need))
nz=lambda a:(np.concatenate([a,np.zeros((pad,27),int)]).reshape(W,32,27).max(axis=1)>0).sum(axis=1).mean()
print("cells a warp must visit: full %.1f culled %.1f" % (nz(full), nz(need)))
