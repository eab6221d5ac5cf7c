timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_parity.py > gpurun_out/shard160.log 2>&1; grep -v Warning gpurun_out/shard160.log | grep -v "^\[rank" | tail -4
if grep -q "shard parity ok" gpurun_out/shard160.log; then
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/shard_parity.py --numchans 8000 --rows 1200 --steps 2 > gpurun_out/shard8000.log 2>&1; grep -v Warning gpurun_out/shard8000.log | grep -v "^\[rank" | tail -4
fi
