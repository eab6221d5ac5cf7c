timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1_v7.json 2> gpurun_out/bench_r1_v7.err; tail -2 gpurun_out/bench_r1_v7.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r1_v7.json"))
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["gpu_launches"], round(d["cpu_baseline"]["value"],1), d["cpu_baseline"]["cores"], d["clocks"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"])
print({k: round(v["ms_per_step"],3) for k,v in d["roofline"]["stages"].items()})
PY
