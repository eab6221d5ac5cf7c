timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1_n2b.json 2> gpurun_out/bench_r1_n2b.err; grep '^{' gpurun_out/bench_r1_n2b.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k: round(v['ms_per_step'],3) for k,v in d['roofline']['stages'].items()}, d['sync_hits_last_step'])
"; grep -v Warning gpurun_out/bench_r1_n2b.err | tail -3
