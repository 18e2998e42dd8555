#!/usr/bin/env python
"""bench.py — particle-substeps/s of the PBF substep on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--scene fluid_million] [--flags none|stable|all] [--mode strict|fast]
                  [--presteps P] [--iterations I]

A "step" is one PBF substep (reference core/src/core.cpp:119-615) over the whole scene.
Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline`, `cpu_baseline`,
`e2e`, `clocks`, `gpu_launches` and `stages` are described in DESIGN.md §6.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# BASELINE.json: "particle-substeps/sec @1M particles, 4 iters (1/2/4/8 B200) + % HBM roofline"
METRIC = "particle-substeps/sec (PBF substep, 4 solver iterations)"

FLAGSETS = {
    "none": dict(scorr=0, xsph=0, vort=0, rest=0.0, fric=0.0),
    "stable": dict(scorr=1, xsph=1, vort=0, rest=0.05, fric=0.1),
    "all": dict(scorr=1, xsph=1, vort=1, rest=0.05, fric=0.1),
}

# Algorithmic bytes per particle per launch (SURVEY.md §8d / DESIGN.md §5): each per-particle
# array a pass logically consumes or produces, counted once; neighbour gathers and the
# neighbour list are NOT algorithmic.
ALG_BYTES = {
    "predict": 48 + 8, "sort": 52, "cells": 6 + 52, "neighbors": 16,
    "lambda": 16, "delta": 28, "xsph": 44, "vort_omega": 40, "vort_apply": 52,
}


def b_alg(iterations: int, flags: dict) -> int:
    """B_alg = 254 + 44*I + 44*[xsph] + 92*[vorticity]."""
    return 254 + 44 * iterations + 44 * int(bool(flags["xsph"])) + 92 * int(bool(flags["vort"]))


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(scene: str, presteps: int, stage: str, brick: bool):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the SAME regime
    (profiles/ncu_traffic.json, keyed "<scene>:<t0|settled>:<kernel>", written by tools/ncu_traffic.py),
    or None when there is no capture of that kernel in that regime."""
    tf = ROOT / "profiles" / "ncu_traffic.json"
    try:
        table = json.loads(tf.read_text())
    except Exception:
        return None
    regime = "t0" if presteps < 60 else ("post_impact" if presteps < 160 else "settled")
    return table.get(f"{scene}:{regime}:{stage}{'_brick' if brick else ''}")


def load_scene(name: str, flags: dict, iterations: int):
    from fluidsimulator_b200 import scenes
    if name in scenes.SCENES:
        sc = scenes.SCENES[name]
    elif name == "block_16m":
        sc = scenes.block_16m()
    elif name.startswith("weak_"):
        sc = scenes.weak_block(int(name.split("_")[1]))
    elif name.startswith("small_"):
        sc = scenes.small_block(int(name.split("_")[1]))
    else:
        raise SystemExit(f"unknown scene {name}")
    params, planes, state = scenes.load_scene(sc)
    params.dt = np.float32(1.0 / 120.0)
    params.enable_scorr, params.enable_xsph, params.enable_vorticity = flags["scorr"], flags["xsph"], flags["vort"]
    params.plane_restitution, params.plane_friction = flags["rest"], flags["fric"]
    params.solver_iterations = iterations
    return params, planes, state


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        if self._thr:
            self._stop.set()
            self._thr.join()
            if not self.samples:
                try:
                    self._once()
                except Exception:
                    pass
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def time_cpu_reference(params, planes, state, steps: int, warmup: int):
    """Times the reference's own CPU implementation (oracle/_ref when present, else the C port)
    with all host threads.  Returns (seconds per step list, kind, threads)."""
    from oracle.oracle_api import Oracle, best_kind
    orc = Oracle(best_kind())
    threads = os.cpu_count() or 1
    orc.set_threads(threads)
    orc.set_params(params)
    orc.set_planes(planes)
    orc.set_state(state)
    for _ in range(warmup):
        orc.step(1)
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.step(1)
        per.append(time.perf_counter() - t0)
    return per, orc.kind, min(threads, orc.max_threads())


def run_reference(args, flags):
    """--impl reference: the reference CPU path on the host cores, same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params, planes, state = load_scene(args.scene, flags, args.iterations)
    n = len(state[0])
    # bounded sample: keep the whole arm within a few minutes (≈2-3 s per 1M-particle substep)
    budget_steps = args.steps + args.warmup
    est = 2.5e-6 * n * budget_steps
    sample = f"{args.scene} from t0, full scene ({n} particles) per step"
    if est > 240.0:
        keep = max(20000, int(n * 240.0 / est))
        order = np.argsort(state[1], kind="stable")[:keep]  # the bottom `keep` particles of the block
        order.sort()
        state = [a[order].copy() for a in state]
        n = keep
        sample = f"{args.scene} from t0, bottom {keep} particles of the block per step (bounded sample)"
    per, kind, threads = time_cpu_reference(params, planes, state, args.steps, args.warmup)
    total = float(sum(per))
    value = n * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-substeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.scene, "particles": n, "solver_iterations": args.iterations,
                   "flags": args.flags, "dt": "1/120", "backend": "cpu", "presteps": 0,
                   "note": "the CPU arm is timed from t0 (a substep costs it 1.6-2.5 s; pre-stepping to the settled "
                           "regime would take minutes): its t0 substeps are its CHEAPEST, so the ratio is conservative"},
        "cpu_baseline": {"value": value, "unit": "particle-substeps/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def golden_digest(scene: str, flagname: str, iterations: int, mode: str):
    """(substep, expected combined16) of the reference's own state digest for this workload
    (tests/golden/million.json, written by tests/golden/make_golden_million.py from the unmodified
    reference CPU solver), or None when no digest is committed for it."""
    if iterations != 4 or mode != "strict":
        return None
    path = ROOT / "tests" / "golden" / "million.json"
    try:
        run = json.loads(path.read_text())["runs"][f"{scene}:{flagname}"]
    except Exception:
        return None
    step = min(int(k) for k in run["steps"])
    return step, run["steps"][str(step)]["combined16"]


def combined16(state6) -> str:
    import hashlib
    h = hashlib.sha256()
    for a in state6:
        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    return h.hexdigest()[:16]


def parity_witness(sol, state, scene: str, flagname: str, iterations: int, mode: str, done: int = 0):
    """Steps `sol` (at substep `done` of a run from t0) to the substep the committed reference digest
    was taken at and compares.  Returns (witness dict or None, substeps now done)."""
    gold = golden_digest(scene, flagname, iterations, mode)
    if gold is None or gold[0] < done:
        return None, done
    step, expected = gold
    sol.step(step - done)
    got = combined16(sol.download())
    return {"step": step, "combined16": got, "expected": expected, "ok": got == expected,
            "source": "tests/golden/million.json (unmodified reference CPU solver)"}, step


def timed_steps(sol, stream, steps: int) -> float:
    """CUDA-event milliseconds of `steps` device-resident substeps on `stream`."""
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    sol.step(steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)


def extra_regimes(args, local: int, stream, mode) -> dict:
    """SURVEY §8d's other regimes, device-timed the same way as `value` (extra keys, not the headline):
    the free-fall lattice at t0, no flags, all flags for the 3 substeps before the reference blows
    up, BASELINE.json's fluid_large iteration sweep, fluid_xlarge, and the 16 M block."""
    from fluidsimulator_b200.capi import Solver
    out = {}

    def run(key, scene, flagname, iterations, presteps, warm, steps, restart=False, stages=False):
        try:
            params, planes, state = load_scene(scene, FLAGSETS[flagname], iterations)
            n = len(state[0])
            sol = Solver(local, n, mode)
            sol.set_params(params)
            sol.set_planes(planes)
            sol.set_stream(stream.cuda_stream)
            sol.upload(state)
            sol.step(presteps + warm)
            if restart:  # graph captured and tables sized; the timed substeps start from t0 again
                sol.upload(state)
            ms = timed_steps(sol, stream, steps)
            out[key] = {"value": n * steps / (ms * 1e-3), "ms_per_step": ms / steps, "particles": n, "scene": scene,
                        "flags": flagname, "solver_iterations": iterations,
                        "substeps": [0 if restart else presteps + warm, (0 if restart else presteps + warm) + steps],
                        "brick_path": sol.brick_status()["active"]}
            if stages:  # per-launch time of the solver passes in this regime (CUDA events, un-graphed)
                sol.profile_enable(True)
                sol.profile_reset()
                sol.step(10)
                prof = sol.profile()
                sol.profile_enable(False)
                out[key]["launch_ms"] = {k: v["ms"] / v["launches"] for k, v in prof.items()
                                         if v["launches"] and k in ("lambda", "delta", "neighbors", "xsph")}
                # and the cuda_step contract (pinned host arrays in and out every substep) in this regime
                import torch
                host = [torch.from_numpy(a).pin_memory().numpy() for a in sol.download()]
                sol.step_host(host, 1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    sol.step_host(host, 1)
                torch.cuda.synchronize()
                out[key]["e2e"] = n * 10 / (time.perf_counter() - t0)
            sol.close()
        except Exception as e:  # an extra must never take the headline down
            out[key] = {"error": str(e)[:200]}

    k, w = args.steps, args.warmup
    # 40 substeps from t0 (free-fall lattice; the floor impact is at substep ~82)
    run("value_t0", args.scene, args.flags, args.iterations, 0, min(w, 10), min(k, 40), stages=True)
    run("value_noflags", args.scene, "none", args.iterations, 0, min(w, 10), min(k, 40))
    run("value_allflags_3steps", args.scene, "all", args.iterations, 0, 3, 3, restart=True)
    for it in (2, 4, 8):
        run(f"fluid_large_iters{it}", "fluid_large", "all", it, 0, 3, 3, restart=True)
    run("fluid_xlarge", "fluid_xlarge", args.flags, args.iterations, 0, w, k)
    run("block_16m", "block_16m", args.flags, args.iterations, 0, 3, min(k, 10))
    return out


def run_ours(args, flags):
    import torch
    from fluidsimulator_b200.capi import PBF_MODE_FAST, PBF_MODE_STRICT, Solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if world > 1:
        from fluidsimulator_b200 import multigpu
        return multigpu.bench(args, flags, rank, world, local)

    params, planes, state = load_scene(args.scene, flags, args.iterations)
    n = len(state[0])
    mode = PBF_MODE_STRICT if args.mode == "strict" else PBF_MODE_FAST
    stream = torch.cuda.Stream()
    sol = Solver(local, n, mode)
    sol.set_params(params)
    sol.set_planes(planes)
    sol.set_stream(stream.cuda_stream)
    sol.upload(state)
    # With vorticity on the REFERENCE trajectory blows up after a few substeps: every
    # phase then restarts from t0 so that no phase runs deeper into the blow-up than the timed one.
    restart = bool(flags["vort"])
    presteps = 0 if restart else args.presteps

    with torch.cuda.stream(stream):
        # parity witness inside the driver-run line: the state this very context reaches must hash to
        # what the unmodified reference CPU solver produced (then the run continues to the timed regime)
        parity, done = parity_witness(sol, state, args.scene, args.flags, args.iterations, args.mode)
        if restart and done:
            sol.upload(state)
            done = 0
        if presteps > done:
            sol.step(presteps - done)
        sol.step(args.warmup)
        torch.cuda.synchronize()
        # the state the timed region starts from: the per-stage profile and the e2e loop below re-run the
        # SAME substeps (the scene changes fast after the floor impact: substeps 125-145 cost 35 % more
        # than 105-125)
        start_state = state if restart else sol.download()
        sampler = ClockSampler(local)
        launches0, retried0 = sol.launch_count(), sol.batches_retried()
        sampler.start()
        ms = timed_steps(sol, stream, args.steps)   # K substeps, device resident, one graph replay per substep
        clocks = sampler.stop()
        launches = sol.launch_count() - launches0
        retried = sol.batches_retried() - retried0
        nbr_total = sol.debug_sizes()[1]
        brick = sol.brick_status()

        sol.upload(start_state)
        if restart:
            sol.step(args.warmup)
        # per-stage CUDA-event timing of the same K substeps (profiling disables the graph)
        sol.profile_enable(True)
        sol.profile_reset()
        ms_prof = timed_steps(sol, stream, args.steps)
        prof = sol.profile()
        sol.profile_enable(False)

        # e2e: the reference-facing call (cuda_step contract): pinned host arrays in and out every step,
        # over the same substeps again
        host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in start_state]
        e2e_steps = max(3, min(args.steps, 20))
        if restart:
            e2e_steps = min(e2e_steps, args.steps)
        sol.step_host(host, 1)   # captures the contract graph (untimed); the loop continues from its result
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sol.step_host(host, 1)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        sol.close()
        extras = {} if args.no_extras else extra_regimes(args, local, stream, mode)

    value = n * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    stages = {}
    for name, rec in prof.items():
        if rec["launches"] and rec["ms"] > 0:
            stages[name] = {"ms_per_step": rec["ms"] / args.steps, "launches_per_step": rec["launches"] / args.steps}
    solver = {k: v for k, v in stages.items() if k in ("lambda", "delta")}
    dom = max(solver, key=lambda k: solver[k]["ms_per_step"]) if solver else None
    roofline = None
    if dom:
        per_launch_s = 1e-3 * prof[dom]["ms"] / prof[dom]["launches"]
        achieved = ALG_BYTES[dom] * n / per_launch_s / 1e9
        roofline = {"bound": "hbm", "kernel": f"k_{dom}{'_brick' if brick['active'] else ''}", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(args.scene, presteps, dom, brick["active"]), "peak_source": peak_src,
                    "alg_bytes_per_particle": ALG_BYTES[dom], "avg_launch_ms": per_launch_s * 1e3,
                    "whole_step_frac": b_alg(args.iterations, flags) * value / 1e9 / peak}

    # the same fraction in the free-fall lattice regime (25 neighbours per particle instead of 30-40)
    t0 = extras.get("value_t0", {}) if isinstance(extras, dict) else {}
    if roofline and "launch_ms" in t0 and dom in t0["launch_ms"]:
        a0 = ALG_BYTES[dom] * n / (t0["launch_ms"][dom] * 1e-3) / 1e9
        roofline["t0"] = {"achieved": a0, "frac": a0 / peak, "avg_launch_ms": t0["launch_ms"][dom],
                          "traffic": ncu_traffic(args.scene, 0, dom, brick["active"]),
                          "whole_step_frac": b_alg(args.iterations, flags) * t0["value"] / 1e9 / peak}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        per, kind, threads = time_cpu_reference(params, planes, state, steps=2, warmup=1)
        cpu_baseline = {"value": n * len(per) / float(sum(per)), "unit": "particle-substeps/s", "cores": threads,
                        "kind": kind, "sample": f"{args.scene} from t0, 1 warm-up + {len(per)} timed substeps, all host threads"}

    entry_bytes = 2 if brick["active"] else 4
    line = {
        "metric": METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.scene, "particles": n, "solver_iterations": args.iterations, "flags": args.flags,
                   "mode": args.mode, "dt": "1/120", "presteps": presteps,
                   "timed_substeps": [presteps + args.warmup, presteps + args.warmup + args.steps],
                   "kernels": "brick (shared-memory staged, 16-bit lists)" if brick["active"] else "global gather (32-bit lists)",
                   "brick": brick,
                   "l2": f"working set exceeds L2: neighbour list {entry_bytes * nbr_total / 1e6:.0f} MB + {n * 16 * 9 / 1e6:.0f} MB of "
                         "per-particle arrays are re-streamed every substep (no explicit flush)",
                   "avg_neighbors": nbr_total / n},
        "parity": parity,
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": n * e2e_steps / e2e_s, "unit": "particle-substeps/s", "h2d_bytes_per_step": 24 * n,
                "d2h_bytes_per_step": 24 * n, "steps": e2e_steps,
                "call": "pbf_step_host (cuda_step contract; one CUDA graph per call on pinned arrays)",
                "value_t0": t0.get("e2e")},
        "gpu_launches": launches, "batches_replayed_in_timed_region": retried, "clocks": clocks, "stages": stages,
        "ms_per_step_profiled": ms_prof / args.steps,
        "extra": extras,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=None)
    ap.add_argument("--flags", default="stable", choices=list(FLAGSETS))
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--presteps", type=int, default=100,
                    help="untimed substeps before the warm-up: 100 puts the timed region after the floor impact of "
                         "fluid_million (substep ~82), SURVEY §8d; `extra.value_t0` is the free-fall lattice")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra regimes (extra.*)")
    ap.add_argument("--iterations", type=int, default=4)
    ap.add_argument("--weak", action="store_true", help="multi-GPU: 2M particles per GPU instead of one fixed scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.scene is None:
        args.scene = "fluid_million"
    flags = FLAGSETS[args.flags]
    if args.impl == "reference":
        run_reference(args, flags)
    else:
        run_ours(args, flags)


if __name__ == "__main__":
    main()
