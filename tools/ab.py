"""A/B timing of library variants on the GPU box (not a benchmark): for every
fluidsimulator_b200/lib/variants/<name>/libpbf_b200.so given on the command line, ms/substep of
the graph-replayed substep and the per-stage split at t0 and after 200 substeps of fluid_million."""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
names = sys.argv[1:]
for name in names:
    lib = ROOT / "fluidsimulator_b200" / "lib" / "variants" / name / "libpbf_b200.so"
    env = dict(os.environ, PBF_B200_LIB=str(lib))
    for pre in (0, 200):
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "60", "--warmup", "5", "--presteps", str(pre),
                              "--no-cpu-baseline"], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(name, pre, "FAILED", out.stderr[-500:])
            continue
        d = json.loads(line[-1])
        st = {k: round(v["ms_per_step"] * 1e3) for k, v in d["stages"].items()}
        print(f"{name:10s} pre={pre:3d} {d['ms_per_step']*1e3:7.1f} us/substep  {d['value']:.3e}  {st}", flush=True)
