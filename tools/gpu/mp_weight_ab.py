"""torchrun: the 16 M block on N slabs with the work-weighted cuts (PBF_SLAB_GHOST_WEIGHT=0.5, default)
and with equal owned counts (=0): ms per substep (CUDA events, max over ranks), owned per rank."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import bench as B
from fluidsimulator_b200 import multigpu
from fluidsimulator_b200.capi import PBF_MODE_STRICT

scene = sys.argv[1] if len(sys.argv) > 1 else "block_16m"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
params, planes, state = B.load_scene(scene, B.FLAGSETS["stable"], 4)
n = len(state[0])
for w in ("0.5", "0", "0.5", "0"):
    os.environ["PBF_SLAB_GHOST_WEIGHT"] = w
    sol = multigpu.make_slab(dist, local, params, planes, state, PBF_MODE_STRICT)
    sol.step(4); sol.step(4)
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); sol.step(20); ev1.record(); torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    own = torch.zeros(world, dtype=torch.int64, device=dev); own[rank] = sol.owned(); dist.all_reduce(own)
    if rank == 0:
        print(f"ghost weight {w}: {float(t.item()) / 20:.4f} ms/substep ({n * 20 / float(t.item()) * 1e3:.3e} particle-substeps/s) "
              f"substeps 8..28, owned {own.tolist()}", flush=True)
    sol.close()
dist.barrier()
dist.destroy_process_group()
