mkdir -p gpurun_out
nvidia-smi -L
timeout 200 python tools/quick_brick.py 50 > gpurun_out/r02d_persist.log 2>&1; echo rc=$? >> gpurun_out/r02d_persist.log
PBF_BRICK_PERSIST=0 timeout 200 python tools/quick_brick.py 50 > gpurun_out/r02d_once.log 2>&1; echo rc=$? >> gpurun_out/r02d_once.log
PBF_BRICK=0 timeout 200 python tools/quick_brick.py 50 > gpurun_out/r02d_legacy.log 2>&1; echo rc=$? >> gpurun_out/r02d_legacy.log
timeout 600 python tests/quick_check.py > gpurun_out/r02d_check.log 2>&1; echo rc=$? >> gpurun_out/r02d_check.log
tail -5 gpurun_out/r02d_*.log
