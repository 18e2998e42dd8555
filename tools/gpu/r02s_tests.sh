mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 ) 2>&1 | tail -n 22
timeout 120 python tools/quick_brick.py 50 2>&1 | grep -E "ms_t0|stage_us_t0|ms_settled|million_280"
python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; tail -n 2 gpurun_out/r02s_bench.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02s_bench.json").read().strip().splitlines()[-1])
print("value", l["value"], "ms", l["ms_per_step"], "launches", l["gpu_launches"], "e2e", l["e2e"]["value"], l["e2e"].get("value_t0"), "parity", l["parity"]["ok"])
print({k: (v.get("value"), v.get("ms_per_step")) for k, v in l["extra"].items()})
PY
