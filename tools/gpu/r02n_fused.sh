mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -n 3
for F in 1 0; do
echo "== PBF_SLAB_FUSED=$F"
PBF_SLAB_FUSED=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 tools/gpu/dbg_mp_slabs.py 50 300 2>&1 | grep -E "after|rror" | cut -c1-150
done
