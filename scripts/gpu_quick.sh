#!/bin/bash
# quick loop: timeline of the fused backward + bench (no cpu/eager) + per-kernel times
mkdir -p gpurun_out
FVAE_TIMELINE=1 timeout 200 python scripts/step_traffic.py cfg2 2>&1 | grep -A8 "fvae timeline" | head -10
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
    print("step ms", d["ms_per_step"], "K1 ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "loss", d["loss"], "e2e ms", d["e2e"]["ms_per_step"])
except Exception as e: print("ERR", e, open("gpurun_out/quick_bench.err").read()[-800:])
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/quick_times.csv python scripts/step_traffic.py cfg2 > /dev/null 2>&1
python scripts/kernel_times.py gpurun_out/quick_times.csv | head -8
