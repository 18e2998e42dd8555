mkdir -p gpurun_out
echo "== persist"; timeout 120 python tools/gpu/dbg_smoke.py 12 3 2>&1 | tail -20
echo "== once"; PBF_BRICK_PERSIST=0 timeout 120 python tools/gpu/dbg_smoke.py 12 3 2>&1 | tail -20
echo "== legacy"; PBF_BRICK=0 timeout 120 python tools/gpu/dbg_smoke.py 12 3 2>&1 | tail -20
