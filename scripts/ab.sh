#!/bin/bash
# A/B kernel timing on ONE box: gpurun_ab/lib_<tag>.so are alternative builds of the library (same ABI); run via
#   gpurun -- 'bash scripts/ab.sh A B'      (box-to-box variance is ~2 %, run-to-run on one box ~0.05 %)
for i in 1 2 3; do
  for v in "$@"; do
    FVAE_B200_LIB=$PWD/gpurun_ab/lib_$v.so timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 |
      python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$v', round(j['ms_per_step'],4), round(j['roofline']['kernel_ms'],4))"
  done
done
