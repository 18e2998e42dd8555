mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 300 python tools/quick_bits.py > gpurun_out/r02b_bits_brick.log 2>&1; echo rc=$? >> gpurun_out/r02b_bits_brick.log
PBF_BRICK=0 timeout 300 python tools/quick_bits.py > gpurun_out/r02b_bits_legacy.log 2>&1; echo rc=$? >> gpurun_out/r02b_bits_legacy.log
timeout 600 python tests/quick_check.py > gpurun_out/r02b_check_brick.log 2>&1; echo rc=$? >> gpurun_out/r02b_check_brick.log
tail -5 gpurun_out/r02b_bits_brick.log gpurun_out/r02b_bits_legacy.log gpurun_out/r02b_check_brick.log
