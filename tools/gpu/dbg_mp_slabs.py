"""Debug aid (torchrun): fluid_million on N real GPUs through the bench's substep range in batches of
`chunk` substeps; per batch: wall ms/substep (max over ranks), owned per rank, re-plans, replays.
  torchrun --nproc-per-node N tools/gpu/dbg_mp_slabs.py [chunk] [total]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import bench as B
from fluidsimulator_b200 import multigpu
from fluidsimulator_b200.capi import PBF_MODE_STRICT

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 25
total = int(sys.argv[2]) if len(sys.argv) > 2 else 260
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
params, planes, state = B.load_scene("fluid_million", B.FLAGSETS["stable"], 4)
sol = multigpu.make_slab(dist, local, params, planes, state, PBF_MODE_STRICT)
if len(sys.argv) > 3:
    sol.set_rebalance(float(sys.argv[3]))
done = 0
while done < total:
    k = min(chunk, total - done)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    sol.step(k)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    own = torch.zeros(world, dtype=torch.int64, device=dev); own[rank] = sol.owned(); dist.all_reduce(own)
    done += k
    if rank == 0:
        print(f"after {done:4d}: {1e3 * float(t.item()) / k:.3f} ms/substep owned {own.tolist()} rebalanced {sol.rebalance_count()} "
              f"retried {sol.batches_retried()} {sol.slab_stats()}", flush=True)
dist.barrier()
dist.destroy_process_group()
