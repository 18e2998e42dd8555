mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 100 --warmup 10 ) > gpurun_out/r02l_bench8.json 2> gpurun_out/r02l_bench8.err
tail -n 5 gpurun_out/r02l_bench8.err; tail -c 400 gpurun_out/r02l_bench8.json
