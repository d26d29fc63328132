#!/bin/bash
# fused-backward validation + A/B timing (one GPU)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_parity_bench_config_gpu.py tests/test_parity_gpu.py tests/test_panel_gpu.py tests/test_shard_invariance_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/r2f_pytest.log 2>&1
for v in 0 1; do
  if [ $v = 1 ]; then export FVAE_BACK_STREAM=1; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r2f_bench_v$v.json 2> gpurun_out/r2f_bench_v$v.err
done
unset FVAE_BACK_STREAM
grep -n "cfg[2-5]:\|passed\|failed\|FAILED\|fvae:" gpurun_out/r2f_pytest.log | head -40
python - <<PY
import json
for v in (0,1):
    try:
        d=json.loads(open("gpurun_out/r2f_bench_v%d.json"%v).read().strip().splitlines()[-1])
        print(v, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["loss"], d["e2e"]["ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(v, "ERR", e, open("gpurun_out/r2f_bench_v%d.err"%v).read()[-800:])
PY
