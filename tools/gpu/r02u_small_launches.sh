mkdir -p gpurun_out
for S in fluid_xlarge fluid_large; do
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv --log-file gpurun_out/launches_r02u_$S.csv python tools/profile_run.py $S stable 20 2 > gpurun_out/r02u_$S.log 2>&1
done
