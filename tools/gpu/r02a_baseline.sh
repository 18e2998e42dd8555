mkdir -p gpurun_out
set -x
python tools/quick_bits.py > gpurun_out/r02a_quick_bits.log 2>&1
for P in 5 200; do
ncu --profile-from-start off --set full --import-source on -k regex:'k_neighbors|k_lambda|k_delta|k_xsph' -c 10 -f -o gpurun_out/r02a_step$P python tools/profile_run.py fluid_million stable $P 1 > gpurun_out/r02a_ncu_step$P.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02a_step$P.csv python bench.py --steps 3 --warmup 3 --presteps $P --no-cpu-baseline > gpurun_out/r02a_launch_bench_$P.log 2>&1
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_t0.json 2> gpurun_out/r02a_bench_t0.err
python bench.py --steps 20 --warmup 5 --presteps 200 --no-cpu-baseline > gpurun_out/r02a_bench_200.json 2> gpurun_out/r02a_bench_200.err
ls -la gpurun_out
