"""How many 128-byte lines one warp-wide neighbour gather touches, under different STORAGE orders of
the particles (the storage order is free: lists are emitted in the reference's traversal order and the
debug surface re-orders on the host, as the sparse table already does).  Counted on the CPU from the
reference's own grid and neighbour lists of fluid_xlarge at substep 250 (oracle: test infrastructure,
not a product path).  Results quoted in DESIGN.md §9.      python tests/gather_lines.py   (~4 min)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np, os
from fluidsimulator_b200 import scenes
from oracle.oracle_api import Oracle
import helpers as H
cache='/tmp/xlarge250.npz'
if not os.path.exists(cache):
    params, planes, state = scenes.load_scene(scenes.SCENES['fluid_xlarge'])
    params = H.configure(params, H.STABLE_FLAGS)
    orc = Oracle("reference"); orc.set_params(params); orc.set_planes(planes); orc.set_state(state)
    orc.step(250)
    g = orc.grid(); prefix, idx = orc.neighbors()
    np.savez(cache, prefix=prefix, idx=idx, **{k:v for k,v in g.items()})
z=np.load(cache)
prefix, idx = z['prefix'], z['idx']
n=len(prefix)
order=z['entry_particle']; ecx,ecy,ecz=z['entry_cx'],z['entry_cy'],z['entry_cz']
start=np.concatenate([[0],prefix[:-1]]); cnt=prefix-start
cell=np.zeros((n,3),int); cell[order,0]=ecx; cell[order,1]=ecy; cell[order,2]=ecz   # per particle
cell-=cell.min(axis=0)
def morton(c):
    def part(v):
        v=v.astype(np.uint64); r=np.zeros_like(v)
        for b in range(10): r|=((v>>b)&1)<<(3*b)
        return r
    return part(c[:,2])|(part(c[:,1])<<1)|(part(c[:,0])<<2)
def blocked(c,B):
    blk=c//B; inn=c%B
    dims=blk.max(axis=0)+1
    return ((blk[:,0]*dims[1]+blk[:,1])*dims[2]+blk[:,2])*(B**3)+((inn[:,0]*B+inn[:,1])*B+inn[:,2])
def slots_from_key(key):
    # particles ordered by (key, id)
    perm=np.lexsort((np.arange(n),key)); slot=np.empty(n,int); slot[perm]=np.arange(n); return slot
dims=cell.max(axis=0)+1
orders={
 'reference (x,y,z)': (cell[:,0]*dims[1]+cell[:,1])*dims[2]+cell[:,2],
 'morton cells': morton(cell),
 'blocked 2x2x2': blocked(cell,2),
 'blocked 4x4x4': blocked(cell,4),
 'column pairs (x, y/2, z, y%2)': ((cell[:,0]*(dims[1]//2+1)+cell[:,1]//2)*dims[2]+cell[:,2])*2+cell[:,1]%2,
 'z-columns 2x2 (x/2,y/2,z,x%2,y%2)': (((cell[:,0]//2)*(dims[1]//2+1)+cell[:,1]//2)*dims[2]+cell[:,2])*4+(cell[:,0]%2)*2+cell[:,1]%2,
}
K=int(cnt.max())
for name,key in orders.items():
    slot=slots_from_key(key)
    inv=np.empty(n,int); inv[slot]=np.arange(n)       # particle at slot
    # lists in slot space, per slot
    pad=(-n)%32
    W=(n+pad)//32
    tot_lines=0; tot_gathers=0; tot_lines32=0
    # build padded matrix [n+pad, K] of neighbour slots (-1 = none)
    M=np.full((n+pad,K),-1,int)
    pc=cnt[inv]; ps=start[inv]
    for k in range(K):
        m=pc>k
        M[:n][m,k]=slot[idx[ps[m]+k]]
    M=M.reshape(W,32,K)
    for rec,label in ((8,'16B'),(4,'32B')):
        lines=np.where(M>=0,M//rec,-1)
        # distinct lines per (warp,k), ignoring -1
        s=np.sort(lines,axis=1)
        d=(np.diff(s,axis=1)!=0).sum(axis=1)+1 - (s[:,0,:]<0)   # distinct values minus the -1 group if present
        active=(M>=0).any(axis=1)
        print(f"{name:36s} {label}: distinct 128-B lines per warp gather = {d[active].mean():.2f}  (gathers {active.sum()})")
print("---- more orders")
def zcol(c,bx,by):
    return (((c[:,0]//bx)*(dims[1]//by+1)+c[:,1]//by)*dims[2]+c[:,2])*(bx*by)+(c[:,0]%bx)*by+c[:,1]%by
more={'z-columns 3x3':zcol(cell,3,3),'z-columns 4x4':zcol(cell,4,4),'z-columns 2x3':zcol(cell,2,3),'z-columns 1x3':zcol(cell,1,3),'z-columns 3x1':zcol(cell,3,1),
      'z-columns 2x2, 2 z-levels': ((((cell[:,0]//2)*(dims[1]//2+1)+cell[:,1]//2)*(dims[2]//2+1)+cell[:,2]//2)*8+(cell[:,2]%2)*4+(cell[:,0]%2)*2+cell[:,1]%2)}
for name,key in more.items():
    slot=slots_from_key(key)
    inv=np.empty(n,int); inv[slot]=np.arange(n)
    pad=(-n)%32; W=(n+pad)//32
    M=np.full((n+pad,K),-1,int)
    pc=cnt[inv]; ps=start[inv]
    for k in range(K):
        m=pc>k
        M[:n][m,k]=slot[idx[ps[m]+k]]
    M=M.reshape(W,32,K)
    for rec,label in ((8,'16B'),):
        lines=np.where(M>=0,M//rec,-1)
        s=np.sort(lines,axis=1)
        d=(np.diff(s,axis=1)!=0).sum(axis=1)+1 - (s[:,0,:]<0)
        active=(M>=0).any(axis=1)
        print(f"{name:36s} {label}: distinct 128-B lines per warp gather = {d[active].mean():.2f}  (gathers {active.sum()})")
print("---- x fastest")
more={'x fastest (z,y,x)': (cell[:,2]*dims[1]+cell[:,1])*dims[0]+cell[:,0],
      'x fastest (y,z,x)': (cell[:,1]*dims[2]+cell[:,2])*dims[0]+cell[:,0],
      'x fastest, 2x2 rows (z/2,y/2,x,z%2,y%2)': (((cell[:,2]//2)*(dims[1]//2+1)+cell[:,1]//2)*dims[0]+cell[:,0])*4+(cell[:,2]%2)*2+cell[:,1]%2,
      'x fastest, 1x2 rows (z,y/2,x,y%2)': ((cell[:,2]*(dims[1]//2+1)+cell[:,1]//2)*dims[0]+cell[:,0])*2+cell[:,1]%2,
      'x fastest, 1x3 rows (z,y/3,x,y%3)': ((cell[:,2]*(dims[1]//3+1)+cell[:,1]//3)*dims[0]+cell[:,0])*3+cell[:,1]%3}
for name,key in more.items():
    slot=slots_from_key(key)
    inv=np.empty(n,int); inv[slot]=np.arange(n)
    pad=(-n)%32; W=(n+pad)//32
    M=np.full((n+pad,K),-1,int)
    pc=cnt[inv]; ps=start[inv]
    for k in range(K):
        m=pc>k
        M[:n][m,k]=slot[idx[ps[m]+k]]
    M=M.reshape(W,32,K)
    for rec,label in ((8,'16B'),(4,'32B')):
        lines=np.where(M>=0,M//rec,-1)
        s=np.sort(lines,axis=1)
        d=(np.diff(s,axis=1)!=0).sum(axis=1)+1 - (s[:,0,:]<0)
        active=(M>=0).any(axis=1)
        print(f"{name:40s} {label}: distinct 128-B lines per warp gather = {d[active].mean():.2f}  (gathers {active.sum()})")
