for cfg in "12 1.1" "12 1.15" "25 1.1" "50 1.15"; do
set -- $cfg
echo "== chunk $1 threshold $2"
PBF_SLAB_CHUNK=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29811 tools/gpu/dbg_mp_slabs.py 100 310 $2 2>&1 | grep -E "after|rror" | cut -c1-140
done
