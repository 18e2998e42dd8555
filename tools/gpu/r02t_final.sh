mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -n 6
python -c "import __graft_entry__ as g; g.smoke()"
( time python bench.py --no-cpu-baseline ) > gpurun_out/r02t_bench1.json 2> gpurun_out/r02t_bench1.err; tail -n 3 gpurun_out/r02t_bench1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench1_driverlike.json 2> gpurun_out/r02t_bench1_driverlike.err
python - <<'PY'
import json
for f in ("gpurun_out/r02t_bench1.json", "gpurun_out/r02t_bench1_driverlike.json"):
    l = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", l["value"], "ms", l["ms_per_step"], "launches", l["gpu_launches"], "e2e", l["e2e"]["value"], l["e2e"].get("value_t0"),
          "parity", l["parity"]["ok"], "frac", l["roofline"]["frac"], l["roofline"].get("t0", {}).get("frac"), "replays", l["batches_replayed_in_timed_region"])
PY
