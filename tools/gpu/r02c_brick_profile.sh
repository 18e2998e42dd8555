mkdir -p gpurun_out
python tools/quick_brick.py > gpurun_out/r02c_base.log 2>&1
for v in z4c3072 z4c2048 z4c2048t128 z8t512; do
  PBF_B200_LIB=fluidsimulator_b200/lib/variants/$v/libpbf_b200.so python tools/quick_brick.py 50 > gpurun_out/r02c_$v.log 2>&1
done
ncu --profile-from-start off --set full --import-source on -k regex:'brick' -c 12 -f -o gpurun_out/r02c_brick_t0 python tools/profile_run.py fluid_million stable 5 1 > gpurun_out/r02c_ncu.log 2>&1
tail -4 gpurun_out/r02c_*.log
