mkdir -p gpurun_out
for P in 5 110; do
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/launches_r02k_step$P.csv python tools/profile_run.py fluid_million stable $P 2 > gpurun_out/r02k_$P.log 2>&1
done
tail -n 2 gpurun_out/r02k_*.log
