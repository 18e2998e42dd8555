mkdir -p gpurun_out
ncu --profile-from-start off --set full --import-source on -k regex:'k_neighbors|k_lambda|k_delta|k_xsph' -c 10 -f -o gpurun_out/r02v_step110 python tools/profile_run.py fluid_million stable 110 1 > gpurun_out/r02v_step110.log 2>&1
tail -n 2 gpurun_out/r02v_step110.log; ls -la gpurun_out/r02v_step110.ncu-rep
