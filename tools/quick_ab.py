"""Seconds-scale A/B of library variants on the GPU box, without torch (its import alone can take
a minute on a fresh box): for `base` (fluidsimulator_b200/lib/libpbf_b200.so) and every
fluidsimulator_b200/lib/variants/<name>/libpbf_b200.so named on the command line,

  * ms/substep of fluid_million (stable flags, STRICT) over substeps 5..65 and 200..260, wall clock
    around the blocking pbf_step (one sync per batch), and the per-launch time of the neighbour
    kernel from the stage profile;
  * a SHA-256 of the state after 280 substeps, and of fluid_large after 40 substeps with all flags
    (the reference's blow-up: sparse cell table, cells with hundreds of particles) — a variant
    must reproduce the bits of `base`.

One process per library (their symbols would interpose each other inside one process).
Not a benchmark: bench.py is.   python tools/quick_ab.py base mask_u1 mask_u2h
"""
import hashlib
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def lib_of(name: str) -> Path:
    pkg = ROOT / "fluidsimulator_b200" / "lib"
    return pkg / "libpbf_b200.so" if name == "base" else pkg / "variants" / name / "libpbf_b200.so"


def digest(arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(a.tobytes())
    return h.hexdigest()[:16]


def one(name: str):
    import numpy as np
    from fluidsimulator_b200 import scenes
    from fluidsimulator_b200.capi import PBF_MODE_STRICT, Solver

    def setup(scene, scorr, xsph, vort):
        params, planes, state = scenes.load_scene(scenes.SCENES[scene])
        params.dt = np.float32(1.0 / 120.0)
        params.enable_scorr, params.enable_xsph, params.enable_vorticity = scorr, xsph, vort
        params.plane_restitution, params.plane_friction = 0.05, 0.1
        sol = Solver(0, len(state[0]), PBF_MODE_STRICT)
        sol.set_params(params)
        sol.set_planes(planes)
        sol.upload(state)
        return sol, len(state[0])

    def timed(sol, steps):
        t0 = time.perf_counter()
        sol.step(steps)
        return (time.perf_counter() - t0) * 1e3 / steps

    out = {"name": name}
    sol, n = setup("fluid_million", 1, 1, 0)
    sol.step(5)
    out["ms_t0"] = round(timed(sol, 60), 4)
    sol.step(135)
    out["ms_200"] = round(timed(sol, 60), 4)
    sol.profile_enable(True)
    sol.profile_reset()
    sol.step(20)
    prof = sol.profile()
    sol.profile_enable(False)
    out["stage_us"] = {k: round(1e3 * v["ms"] / v["launches"], 1) for k, v in prof.items() if v["launches"]}
    out["sha_million_280"] = digest(sol.download())
    out["nbr_total"] = sol.debug_sizes()[1]
    sol.close()
    sol, _ = setup("fluid_large", 1, 1, 1)
    sol.step(40)
    out["sha_large_all_40"] = digest(sol.download())
    out["large_sparse"] = int(sol.lib.pbf_debug_grid_is_sparse(sol.ctx))
    sol.close()
    print(json.dumps(out), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        return one(sys.argv[2])
    names = sys.argv[1:] or ["base"]
    results = []
    for name in names:
        lib = lib_of(name)
        if not lib.exists():
            print(f"{name}: {lib} missing", flush=True)
            continue
        env = dict(os.environ, PBF_B200_LIB=str(lib))
        t0 = time.perf_counter()
        r = subprocess.run([sys.executable, __file__, "--one", name], env=env, capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            print(f"{name}: FAILED rc={r.returncode} {r.stderr[-400:]}", flush=True)
            continue
        d = json.loads(lines[-1])
        d["wall_s"] = round(time.perf_counter() - t0, 1)
        results.append(d)
        print(json.dumps(d), flush=True)
    if results:
        ref = results[0]
        for d in results[1:]:
            same = all(d[k] == ref[k] for k in ("sha_million_280", "sha_large_all_40", "nbr_total"))
            print(f"{d['name']:10s} vs {ref['name']}: bits {'IDENTICAL' if same else 'DIFFER'}; "
                  f"t0 {ref['ms_t0']:.3f} -> {d['ms_t0']:.3f} ms, settled {ref['ms_200']:.3f} -> {d['ms_200']:.3f} ms, "
                  f"neighbours {ref['stage_us'].get('neighbors')} -> {d['stage_us'].get('neighbors')} us", flush=True)


if __name__ == "__main__":
    main()
