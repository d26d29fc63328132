#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_parity_bench_config_gpu.py tests/test_parity_gpu.py tests/test_panel_gpu.py -m gpu -q -s -p no:cacheprovider -k "bench_shapes or chain or index or multi_tile or more_tiles" 2>&1 | tail -12) > gpurun_out/r2h_pytest.log 2>&1
grep -n "cfg[2-5]:\|passed\|failed\|FAILED\|fvae:" gpurun_out/r2h_pytest.log | head
FVAE_TIMELINE=1 timeout 200 python scripts/step_traffic.py cfg2 2>&1 | grep -A12 "fvae timeline" | head -14
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["loss"], d["e2e"]["ms_per_step"], d["e2e"]["value"])
except Exception as e: print("ERR", e, open("gpurun_out/r2h_bench.err").read()[-800:])
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_times.csv python scripts/step_traffic.py cfg2 > /dev/null 2>&1
python scripts/kernel_times.py gpurun_out/r2h_times.csv | head -20
