#!/bin/bash
# round-end profiles on ONE box, sized for the 64 MiB return limit: the profiles/ inputs of gpu_profiles.sh with the two front
# captures summarised on the box, then full captures of the GRU / heads kernels (summaries kept, reports dropped)
mkdir -p gpurun_out
bash scripts/gpu_profiles.sh > /dev/null 2>&1
python scripts/kernel_times.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench.txt 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_tc_front_tma.ncu-rep gpurun_out/r02_k1_front_tma > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r02_tc_back_tma.ncu-rep gpurun_out/r02_back_tma > /dev/null 2>&1
rm -f gpurun_out/r02_tc_front_tma.ncu-rep gpurun_out/r02_tc_back_tma.ncu-rep
for K in tc_gru_bwd tc_gru_fwd heads_tc_fwd heads_tc_sweep; do
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$K --launch-count 1 -o gpurun_out/r02_$K -f python scripts/step_traffic.py cfg2 > gpurun_out/r02_ncu_$K.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r02_$K.ncu-rep gpurun_out/r02_$K > /dev/null 2>&1
  rm -f gpurun_out/r02_$K.ncu-rep
done
ls -la gpurun_out | tail -30; du -sh gpurun_out
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_cfg2_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['stream_new_rows']['ms_per_step'], d['clocks'])"
