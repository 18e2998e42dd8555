mkdir -p gpurun_out
N=$1
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N ) > gpurun_out/r02q_bench$N.json 2> gpurun_out/r02q_bench$N.err
tail -n 4 gpurun_out/r02q_bench$N.err; tail -c 300 gpurun_out/r02q_bench$N.json
