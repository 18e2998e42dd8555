mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -n 6
( time python bench.py ) > gpurun_out/r02p_bench1.json 2> gpurun_out/r02p_bench1.err; tail -n 4 gpurun_out/r02p_bench1.err
( time python bench.py --impl reference ) > gpurun_out/r02p_ref.json 2> gpurun_out/r02p_ref.err; tail -n 4 gpurun_out/r02p_ref.err
python -c "import __graft_entry__ as g; g.smoke()"
