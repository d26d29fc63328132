#!/bin/bash
# full ncu captures of several kernels of one cfg2 step: args = kernel-name regexes
mkdir -p gpurun_out
for K in "$@"; do
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$K --launch-count 1 -o gpurun_out/r02_$K -f python scripts/step_traffic.py cfg2 > gpurun_out/r02_ncu_$K.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
