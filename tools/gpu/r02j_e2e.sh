mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_app.py -m gpu -x -q ) 2>&1 | tail -n 12
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -n 3 gpurun_out/r02j_bench.err
python bench.py --no-extras --no-cpu-baseline --presteps 0 --steps 40 > gpurun_out/r02j_bench_t0.json 2>> gpurun_out/r02j_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r02j_bench.json", "gpurun_out/r02j_bench_t0.json"):
    l = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"])
PY
