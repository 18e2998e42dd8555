"""Multi-process driver of the x-slab decomposition: one rank per GPU (torchrun); halos and migrants
move inside libpbf_b200.so by direct stores into the x-neighbours' cudaIpc windows (default) or by
ncclSend/ncclRecv.  torch.distributed is plumbing only: it ships the NCCL unique id,
gathers results for checks and the parity witness, and takes the max over ranks of the device-timed
region.

  torchrun --nproc-per-node N -m fluidsimulator_b200.multigpu --check --scene fluid_large --steps 10
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

from .capi import PBF_MODE_FAST, PBF_MODE_STRICT, SlabSolver, Solver, comm_unique_id, slab_plan


# ---- host-side logic (also exercised on CPU with the gloo backend, tests/test_slab_plan.py::test_two_gloo_ranks_split_and_gather + tests/gloo_worker.py) ---
def cell_x(px: np.ndarray, h: float) -> np.ndarray:
    """x-cell of a position: floor(x * (1.0f / h)) in float32 (reference core.cpp:28-34)."""
    inv = np.float32(1.0) / np.float32(h)
    return np.floor(px.astype(np.float32) * inv).astype(np.int64)


def owned_mask(px: np.ndarray, h: float, cuts: np.ndarray, rank: int) -> np.ndarray:
    c = cell_x(px, h)
    return (c >= int(cuts[rank])) & (c < int(cuts[rank + 1]))


def broadcast_bytes(dist, payload: bytes | None, nbytes: int, device) -> bytes:
    """Rank 0's `payload` to every rank (uint8 tensor on `device`: works for nccl and gloo)."""
    import torch
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def gather_global(dist, gid: np.ndarray, state6, n_global: int, device):
    """Every rank's (global ids, SoA state) assembled on rank 0 in original particle order.
    Returns the six global arrays on rank 0, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = gid.shape[0]
    dist.all_reduce(counts)
    nmax = int(counts.max().item())
    mine = torch.zeros((7, nmax), dtype=torch.float64, device=device)  # float64 holds ids < 2^53 and float32 exactly
    mine[0, : gid.shape[0]] = torch.from_numpy(gid.astype(np.float64)).to(device)
    for k in range(6):
        mine[1 + k, : gid.shape[0]] = torch.from_numpy(state6[k].astype(np.float64)).to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    if rank != 0:
        return None
    out = [np.full(n_global, np.nan, dtype=np.float32) for _ in range(6)]
    seen = 0
    for r in range(world):
        n = int(counts[r].item())
        block = parts[r].cpu().numpy()
        ids = block[0, :n].astype(np.int64)
        for k in range(6):
            out[k][ids] = block[1 + k, :n].astype(np.float32)
        seen += n
    if seen != n_global:
        raise RuntimeError(f"slabs hold {seen} particles, expected {n_global}")
    return out


def rebalance(dist, sol: SlabSolver, h: float, device) -> np.ndarray:
    """Collective: gathers the x coordinates of every slab, plans new cuts on them
    (pbf_slab_plan — the same planner the upload uses) and installs them.  Misplaced particles
    migrate during the next substep; results do not depend on the cuts."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    _, st = sol.slab_download()
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = st[0].shape[0]
    dist.all_reduce(counts)
    nmax = int(counts.max().item())
    mine = torch.zeros(nmax, dtype=torch.float32, device=device)
    mine[: st[0].shape[0]] = torch.from_numpy(st[0]).to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    px = np.concatenate([parts[r][: int(counts[r].item())].cpu().numpy() for r in range(world)])
    cuts = slab_plan(px, h, world)
    sol.set_cuts(cuts[rank], cuts[rank + 1])
    return cuts


# ---- GPU side ------------------------------------------------------------------------------------
def make_slab(dist, local: int, params, planes, state, mode, stream_ptr=None) -> SlabSolver:
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.device("cuda", local)
    uid = broadcast_bytes(dist, comm_unique_id() if rank == 0 else None, 128, device)
    sol = SlabSolver(local, 0, mode)
    sol.set_params(params)
    sol.set_planes(planes)
    if stream_ptr is not None:
        sol.set_stream(stream_ptr)
    sol.comm_init(rank, world, uid)
    sol.slab_upload(state)
    return sol


def extra_run(dist, B, args, flags, scene: str, mode, rank, world, local, device, stream, barrier) -> dict:
    """One more workload on the same ranks: K substeps after a warm-up, CUDA events, max over ranks;
    with a parity witness when tests/golden/million.json holds a reference digest for the scene."""
    import torch
    params, planes, state = B.load_scene(scene, flags, args.iterations)
    n = len(state[0])
    sol = make_slab(dist, local, params, planes, state, mode, stream.cuda_stream)
    parity, done = None, 0
    gold = B.golden_digest(scene, args.flags, args.iterations, args.mode)
    with torch.cuda.stream(stream):
        if gold is not None:
            step, expected = gold
            sol.step(step)
            done = step
            gid_w, st_w = sol.slab_download()
            full = gather_global(dist, gid_w, st_w, n, device)
            if rank == 0:
                got = B.combined16(full)
                parity = {"step": step, "combined16": got, "expected": expected, "ok": got == expected, "slabs": world}
            del full
        warm = max(3, 3 - done)
        sol.step(warm)      # first batches: plain launches (NCCL warm-up), then the captured graph
        sol.step(warm)
        steps = min(args.steps, 10)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        sol.step(steps)
        ev1.record(stream)
        barrier()
        ms_t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        owned_t = torch.zeros(world, dtype=torch.int64, device=device)
        owned_t[rank] = sol.owned()
        dist.all_reduce(owned_t)
    ms = float(ms_t.item())
    out = {"value": n * steps / (ms * 1e-3), "ms_per_step": ms / steps, "particles": n, "scene": scene,
           "substeps": [done + 2 * warm, done + 2 * warm + steps], "transport": sol.transport(),
           "owned_per_rank": [int(x) for x in owned_t.tolist()], "parity": parity}
    sol.close()
    return out


def bench(args, flags, rank: int, world: int, local: int):
    """bench.py's N > 1 arm: strong scaling of one scene over `world` slabs (or --weak)."""
    import torch
    import torch.distributed as dist
    import bench as B

    scene = args.scene if not args.weak else f"weak_{world}"
    params, planes, state = B.load_scene(scene, flags, args.iterations)
    n = len(state[0])
    mode = PBF_MODE_STRICT if args.mode == "strict" else PBF_MODE_FAST
    device = torch.device("cuda", local)
    stream = torch.cuda.Stream()
    sol = make_slab(dist, local, params, planes, state, mode, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # parity witness inside the driver-run line: the slabs' state, gathered by global particle id,
        # must hash to what the unmodified reference CPU solver produced at the same substep
        parity, done = None, 0
        gold = None if args.weak else B.golden_digest(scene, args.flags, args.iterations, args.mode)
        if gold is not None:
            step, expected = gold
            sol.step(step)
            done = step
            gid_w, st_w = sol.slab_download()
            full = gather_global(dist, gid_w, st_w, n, device)
            if rank == 0:
                got = B.combined16(full)
                parity = {"step": step, "combined16": got, "expected": expected, "ok": got == expected,
                          "slabs": world, "source": "tests/golden/million.json (unmodified reference CPU solver)"}
            del full
        if args.presteps > done:
            sol.step(args.presteps - done)
        sol.step(args.warmup)
        barrier()
        sampler = B.ClockSampler(local)
        launches0, retried0, replans0 = sol.launch_count(), sol.batches_retried(), sol.rebalance_count()
        stats0 = sol.slab_stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        ev0.record(stream)
        sol.step(args.steps)
        ev1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms_t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        ms = float(ms_t.item())
        launches = sol.launch_count() - launches0
        retried, replans = sol.batches_retried() - retried0, sol.rebalance_count() - replans0
        stats1 = sol.slab_stats()
        owned = sol.owned()
        payload = sol.payload_bytes()
        transport = sol.transport()

        # per-stage split of this rank (profiling mode: CUDA events around every stage)
        sol.profile_enable(True)
        sol.profile_reset()
        sol.step(min(args.steps, 20))
        torch.cuda.synchronize()
        prof = sol.profile()
        sol.profile_enable(False)
        psteps = min(args.steps, 20)

        # e2e: every rank round-trips ITS particles through pinned host memory each substep
        room = int(sol.owned() * 1.25) + 4096
        pinned = (torch.empty(room, dtype=torch.int64).pin_memory().numpy(),
                  [torch.empty(room, dtype=torch.float32).pin_memory().numpy() for _ in range(6)])
        gid, host = sol.slab_download(pinned)
        e2e_steps = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            # positions and velocities of this rank's particles in from pinned host memory, one substep,
            # out again (the ids of the particles a rank holds are read back whenever they are needed,
            # not shipped both ways every substep)
            sol.slab_upload_owned(None, host)
            sol.step(1)
            _, host = sol.slab_download(pinned, ids=False)
        barrier()
        e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)

    # SURVEY §8e's larger workloads, device-timed the same way (extra keys, not the headline): the
    # 16 M block on the same N slabs (strong scaling against extra.block_16m of the 1-GPU line) and
    # 2 M particles per GPU (weak scaling)
    extras = {}
    if not args.no_extras and not args.weak:
        sol.close()
        for key, name in (("block_16m", "block_16m"), (f"weak_2m_per_gpu", f"weak_{world}")):
            try:
                extras[key] = extra_run(dist, B, args, flags, name, mode, rank, world, local, device, stream, barrier)
            except Exception as e:  # an extra must never take the headline down
                extras[key] = {"error": str(e)[:200]}

    # every rank's per-stage split (interior slabs carry two ghost sides, edge slabs one)
    my_stages = {k: v["ms"] / psteps for k, v in prof.items() if v["launches"]}
    all_stages = [None] * world
    dist.all_gather_object(all_stages, my_stages)
    tot_launch = torch.tensor([launches], dtype=torch.int64, device=device)
    dist.all_reduce(tot_launch)
    owned_t = torch.zeros(world, dtype=torch.int64, device=device)
    owned_t[rank] = owned
    dist.all_reduce(owned_t)
    if rank != 0:
        return
    value = n * args.steps / (ms * 1e-3)
    peak, peak_src = B.measured_peaks()
    stages = {k: {"ms_per_step": v["ms"] / psteps, "launches_per_step": v["launches"] / psteps}
              for k, v in prof.items() if v["launches"]}
    solver = {k: v for k, v in stages.items() if k in ("lambda", "delta")}
    dom = max(solver, key=lambda k: solver[k]["ms_per_step"]) if solver else None
    roofline = None
    if dom:
        per_launch_s = 1e-3 * prof[dom]["ms"] / prof[dom]["launches"]
        achieved = B.ALG_BYTES[dom] * owned / per_launch_s / 1e9   # rank 0's slab, rank 0's kernel
        roofline = {"bound": "hbm", "kernel": f"k_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "scope": "rank 0, per GPU",
                    "alg_bytes_per_particle": B.ALG_BYTES[dom], "avg_launch_ms": per_launch_s * 1e3,
                    "whole_step_frac": B.b_alg(args.iterations, flags) * value / 1e9 / (peak * world)}
    line = {
        "metric": B.METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": scene, "particles": n, "solver_iterations": args.iterations, "flags": args.flags,
                   "mode": args.mode, "dt": "1/120", "presteps": args.presteps,
                   "decomposition": f"{world} x-slabs, 2 ghost layers, " + (
                       "direct peer stores into cudaIpc windows of the x-neighbours over NVLink + flag kernels "
                       "(NCCL: IPC handles and the per-batch status max-reduce only)" if transport == "peer-stores"
                       else "ncclSend/ncclRecv between x-neighbours"),
                   "transport": transport,
                   "owned_per_rank": [int(x) for x in owned_t.tolist()],
                   "l2": "per-slab working set (neighbour list + particle arrays) re-streamed every substep; no explicit flush"},
        "parity": parity, "extra": extras,
        "roofline": roofline, "cpu_baseline": None,
        "e2e": {"value": n * e2e_steps / float(e2e_t.item()), "unit": "particle-substeps/s",
                "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 24 * n, "steps": e2e_steps,
                "call": "per rank: pbf_slab_upload_owned (pos, vel) + pbf_step(1) + pbf_slab_download (pos, vel)"},
        "gpu_launches": int(tot_launch.item()), "batches_replayed_in_timed_region": retried,
        "cut_replans_in_timed_region": replans, "clocks": clocks, "stages": stages,
        "stages_per_rank_ms": [{k: round(v, 4) for k, v in st.items()} for st in all_stages],
        "exchange": {"per_substep": (stats1["exchanges"] - stats0["exchanges"]) / args.steps,
                     "payload_bytes_per_substep_rank0": payload,
                     "capacity_bytes_per_substep_rank0": (stats1["bytes_sent"] - stats0["bytes_sent"]) / args.steps,
                     "ghosts_rank0": stats1["ghosts"], "hops": stats1["hops"]},
    }
    print(json.dumps(line), flush=True)


def main():
    import torch
    import torch.distributed as dist
    import bench as B
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="compare the slab result with one GPU, bit for bit")
    ap.add_argument("--scene", default="fluid_large")
    ap.add_argument("--flags", default="all")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=4)
    ap.add_argument("--drift", type=float, default=0.0, help="initial x velocity of every particle (forces migration / re-balancing)")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    params, planes, state = B.load_scene(args.scene, B.FLAGSETS[args.flags], args.iterations)
    if args.drift:
        state = [a.copy() for a in state]
        state[3][:] = np.float32(args.drift)
    sol = make_slab(dist, local, params, planes, state, PBF_MODE_STRICT)
    sol.step(args.steps // 2)            # first batch: plain stream launches (NCCL warm-up)
    sol.step(args.steps - args.steps // 2)   # second batch: the captured CUDA graph, exchanges included
    gid, st = sol.slab_download()
    full = gather_global(dist, gid, st, len(state[0]), torch.device("cuda", local))
    ok = True
    if rank == 0:
        ref = Solver(local, len(state[0]), PBF_MODE_STRICT)
        ref.set_params(params)
        ref.set_planes(planes)
        ref.upload(state)
        ref.step(args.steps)
        bad = [k for k, (a, b) in enumerate(zip(full, ref.download()))
               if not np.array_equal(a.view(np.uint32), b.view(np.uint32))]
        ok = not bad
        print(f"slab check {'ok' if ok else 'FAILED ' + str(bad)}: {args.scene}, {dist.get_world_size()} slabs, "
              f"{args.steps} substeps, stats {sol.slab_stats()}, re-balanced {sol.rebalance_count()}x, "
              f"owned on rank 0: {sol.owned()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    raise SystemExit(0 if ok else 1)


if __name__ == "__main__":
    main()
