#!/bin/bash
# the evidence under profiles/: launch list of the bench command, per-kernel DRAM bytes of one step, full captures of the two front kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(tc_|heads_|loss_reduce|p2p_|window_index|gather_windows|adam_)' -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager --no-e2e > gpurun_out/r02_launches_bench.log 2>&1
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_step_traffic.csv python scripts/step_traffic.py cfg2 > gpurun_out/r02_traffic.log 2>&1
python scripts/step_traffic.py --parse gpurun_out/r02_step_traffic.csv cfg2 > gpurun_out/r02_step_traffic.json 2>> gpurun_out/r02_traffic.log
for K in tc_front_tma tc_back_tma; do
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$K --launch-count 1 -o gpurun_out/r02_$K -f python scripts/step_traffic.py cfg2 > gpurun_out/r02_ncu_$K.log 2>&1
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2_1gpu.json 2> gpurun_out/r02_bench_cfg2_1gpu.err
tail -c 1500 gpurun_out/r02_bench_cfg2_1gpu.json
