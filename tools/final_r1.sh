#!/bin/bash
# final round-1 measurement pass: tests, bench, reference arm, ncu launch list (no full captures)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("VALUE", d["value"], "E2E", d["e2e"]["value"], "FRAC", d["roofline"]["frac"], d["kernels_ms_per_step"], d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["parity_mismatch_words"])
PY
python tools/config_sweep.py > /dev/null 2>&1; python - <<'PY'
import json
r=json.load(open("gpurun_out/configs.json"))
print({k:(v.get("gates_per_s_kernels") or v.get("chain_latency_ms_best")) for k,v in r.items() if isinstance(v,dict) and ("gates_per_s_kernels" in v or "chain_latency_ms_best" in v)})
PY
