mkdir -p gpurun_out
T="tests/test_gpu_brick.py::test_brick_bit_exact_small tests/test_gpu_brick.py::test_brick_random_clouds tests/test_gpu_parity.py::test_step_host_contract_pinned_graph tests/test_gpu_parity.py::test_snapshot_is_the_state_of_its_moment tests/test_gpu_parity.py::test_strict_bit_exact_small tests/test_gpu_slabs.py -k 'not two_gpus'"
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool" 
  eval timeout 1500 compute-sanitizer --tool $tool --target-processes all python -m pytest $T -m gpu -x -q 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | sort | uniq -c | sort -rn | head -12
done > gpurun_out/r02r_sanitizer.txt 2>&1
cat gpurun_out/r02r_sanitizer.txt
