mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02g_gputests.log 2>&1
( time python bench.py ) > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err
tail -n 8 gpurun_out/r02g_gputests.log; tail -n 5 gpurun_out/r02g_bench.err gpurun_out/r02g_ref.err
