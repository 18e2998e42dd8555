mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q ) > gpurun_out/r02h_slabtests.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10 ) > gpurun_out/r02h_bench2.json 2> gpurun_out/r02h_bench2.err
tail -n 6 gpurun_out/r02h_slabtests.log; tail -n 4 gpurun_out/r02h_bench2.err; tail -c 600 gpurun_out/r02h_bench2.json
