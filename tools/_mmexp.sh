timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_chain.py -m gpu -q -x -k "fused or chain" 2>&1 | tail -2
run() {
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_tmp$1.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp$1.json"))
print("$1", round(d["value"]), round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d["roofline"]["stages"].items()}, d["sync_hits_last_step"])
PY
}
run overlap
GRCUDA_CHAIN_NO_OVERLAP=1 run serial
