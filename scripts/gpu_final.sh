#!/bin/bash
# round-end evidence on ONE box: full GPU test set, smoke, the profiles/ inputs, per-kernel times of the big-H micro-batches
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6) | tee gpurun_out/r02_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke.log
bash scripts/gpu_profiles.sh > /dev/null 2>&1
python scripts/kernel_times.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench.txt 2>/dev/null; head -16 gpurun_out/r02_launches_bench.txt
for wl in cfg4 cfg3; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_times_$wl.csv python scripts/step_traffic.py $wl > /dev/null 2>&1
  echo "== $wl"; python scripts/kernel_times.py gpurun_out/r02_times_$wl.csv 2>/dev/null | head -14 | tee gpurun_out/r02_times_$wl.txt
done
for wl in cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-e2e > gpurun_out/r02_bench_${wl}_1gpu.json 2> gpurun_out/r02_bench_${wl}_1gpu.err
  python -c "import json; d=json.loads(open('gpurun_out/r02_bench_${wl}_1gpu.json').read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], d['value'], d['roofline']['step']['frac'] if d['roofline'] and 'step' in d['roofline'] else d['roofline'])"
done
