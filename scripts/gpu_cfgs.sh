#!/bin/bash
# kernel-time breakdown of one micro-batch step of cfg3 / cfg4 / cfg5 shapes on one GPU
mkdir -p gpurun_out
for wl in cfg4 cfg3; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_times_$wl.csv python scripts/step_traffic.py $wl > /dev/null 2>&1
  echo "== $wl"; python scripts/kernel_times.py gpurun_out/r2_times_$wl.csv 2>/dev/null | head -16
done
