GRCUDA_CHAIN_3STAGE=1 timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -2
run() {
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_tmp$1.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp$1.json"))
print("$1", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), {k: round(v["ms_per_step"],3) for k,v in d["roofline"]["stages"].items()}, d["sync_hits_last_step"])
PY
}
run overlap2
GRCUDA_CHAIN_3STAGE=1 run overlap3
GRCUDA_CHAIN_3STAGE=1 GRCUDA_FFT_VARIANT=0 run overlap3_fft128
