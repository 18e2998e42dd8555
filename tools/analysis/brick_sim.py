"""CPU-side sizing of the CTA-brick shared-memory path: on the reference's own grid and
neighbour lists (oracle: analysis infrastructure, not a product path) count, for bricks of
BX x BY z-columns x BZ cells,
  * owned particles and halo records per brick (shared-memory tile size),
  * the bank-conflict degree of a warp-wide LDS.128 gather at list position k (a 128-bit shared load
    is served per quarter warp; lanes conflict when their 16-byte chunks fall into the same chunk
    column mod 8 at different addresses) — cycles per gather = sum over quarters of the max
    multiplicity,
  * for comparison the distinct 128-byte lines of the same gather through L1 (global order).
  python tools/analysis/brick_sim.py [scene] [substeps]"""
import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
from fluidsimulator_b200 import scenes
from oracle.oracle_api import Oracle
import helpers as H

scene = sys.argv[1] if len(sys.argv) > 1 else 'fluid_xlarge'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 250
cache = f'/tmp/{scene}_{steps}.npz'
if not os.path.exists(cache):
    params, planes, state = scenes.load_scene(scenes.SCENES[scene])
    params = H.configure(params, H.STABLE_FLAGS)
    orc = Oracle("reference"); orc.set_params(params); orc.set_planes(planes); orc.set_state(state)
    orc.step(steps)
    g = orc.grid(); prefix, idx = orc.neighbors()
    np.savez(cache, prefix=prefix, idx=idx, **{k: v for k, v in g.items()})
z = np.load(cache)
prefix, idx = z['prefix'].astype(np.int64), z['idx'].astype(np.int64)
n = len(prefix)
order = z['entry_particle'].astype(np.int64)
cell_sorted = np.stack([z['entry_cx'], z['entry_cy'], z['entry_cz']], 1).astype(np.int64)
cell_sorted -= cell_sorted.min(axis=0) - 1          # one padding layer like the device table
slot_of = np.empty(n, np.int64); slot_of[order] = np.arange(n)
start = np.concatenate([[0], prefix[:-1]]); cnt = prefix - start
dims = cell_sorted.max(axis=0) + 2
key = (cell_sorted[:, 0] * dims[1] + cell_sorted[:, 1]) * dims[2] + cell_sorted[:, 2]
ncell = int(dims.prod())
cstart = np.searchsorted(key, np.arange(ncell), 'left'); cend = np.searchsorted(key, np.arange(ncell), 'right')
print(f"{scene} after {steps} substeps: n={n}, avg neighbours {cnt.mean():.1f}, max {cnt.max()}, dims {dims}")
# neighbour slots per sorted slot
nbr_slot = slot_of[idx]                      # CSR over ORIGINAL particle ids
def lists_of_slot(s):                        # list of slot s
    p = order[s]; return nbr_slot[start[p]:start[p] + cnt[p]]

def run(BX, BY, BZ, maxwarps=4000):
    nb = [-(-(int(dims[a]) - 2) // b) for a, b in ((0, BX), (1, BY), (2, BZ))]
    own_counts, halo_counts = [], []
    cyc, lines, gathers = 0, 0, 0
    rng = np.random.default_rng(0)
    bricks = [(i, j, k) for i in range(nb[0]) for j in range(nb[1]) for k in range(nb[2])]
    warps_done = 0
    for (bi, bj, bk) in bricks:
        x0, y0, z0 = 1 + bi * BX, 1 + bj * BY, 1 + bk * BZ
        z1 = min(z0 + BZ, int(dims[2]) - 1)
        # halo runs
        tile_base = {}; tot = 0; own = []
        for X in range(x0 - 1, min(x0 + BX, int(dims[0]) - 1) + 1):
            for Y in range(y0 - 1, min(y0 + BY, int(dims[1]) - 1) + 1):
                c0 = (X * dims[1] + Y) * dims[2] + (z0 - 1); c1 = (X * dims[1] + Y) * dims[2] + z1
                s, e = cstart[c0], cend[c1]
                tile_base[(X, Y)] = (tot, s); tot += e - s
                if x0 <= X < x0 + BX and y0 <= Y < y0 + BY and X <= dims[0] - 2 and Y <= dims[1] - 2:
                    os_, oe = cstart[c0 + 1], cend[c1 - 1]
                    own.extend(range(os_, oe))
        if not own: continue
        own_counts.append(len(own)); halo_counts.append(tot)
        if warps_done >= maxwarps: continue
        # tile index of a sorted slot j (must be in the halo)
        def tidx(j):
            c = cell_sorted[j]; b, s = tile_base[(int(c[0]), int(c[1]))]; return b + (j - s)
        own = np.array(own)
        for w0 in range(0, len(own), 32):
            lanes = own[w0:w0 + 32]
            L = [lists_of_slot(s) for s in lanes]
            T = [np.array([tidx(j) for j in l]) for l in L]
            kmax = max(len(l) for l in L)
            for k in range(kmax):
                act = [q for q in range(len(lanes)) if len(L[q]) > k]
                ti = {q: T[q][k] for q in act}
                # LDS.128: per quarter warp
                for qw in range(4):
                    m = {}
                    for q in act:
                        if q // 8 == qw:
                            m.setdefault(ti[q] % 8, set()).add(ti[q])
                    if m: cyc += max(len(v) for v in m.values())
                lines += len({L[q][k] // 8 for q in act})
                gathers += 1
            warps_done += 1
    oc, hc = np.array(own_counts), np.array(halo_counts)
    print(f"brick {BX}x{BY}x{BZ}: {len(oc)} non-empty bricks; owned mean {oc.mean():.0f} max {oc.max()}; halo records mean {hc.mean():.0f} "
          f"p99 {np.percentile(hc, 99):.0f} max {hc.max()} ({hc.max() * 16 / 1024:.0f} KB); halo/owned {hc.sum() / oc.sum():.2f}; "
          f"LDS.128 cycles/gather {cyc / max(gathers,1):.2f}; L1 lines/gather {lines / max(gathers,1):.2f} ({gathers} gathers)")

for shape in ((4, 4, 8), (2, 2, 16), (4, 4, 4), (3, 3, 8), (2, 4, 8), (4, 4, 16)):
    run(*shape, maxwarps=600)
