mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10 ) > gpurun_out/r02i_bench2.json 2> gpurun_out/r02i_bench2.err
tail -n 6 gpurun_out/r02i_bench2.err; tail -c 1500 gpurun_out/r02i_bench2.json
