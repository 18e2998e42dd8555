mkdir -p gpurun_out
PBF_BRICK=0 ncu --profile-from-start off --set full --import-source on -k regex:'k_lambda|k_delta' -c 4 -f -o gpurun_out/r02f_legacy_t0 python tools/profile_run.py fluid_million stable 5 1 > gpurun_out/r02f_legacy.log 2>&1
ncu --profile-from-start off --set full --import-source on -k regex:'k_brick_persist' -c 4 -f -o gpurun_out/r02f_persist_t0 python tools/profile_run.py fluid_million stable 5 1 > gpurun_out/r02f_persist.log 2>&1
tail -3 gpurun_out/r02f_*.log; ls -la gpurun_out
