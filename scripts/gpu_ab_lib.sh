#!/bin/bash
# A/B on ONE box: the in-tree library vs other builds of the same ABI (factorvae_b200/libfvae_<name>.so), 3 rounds
# usage: gpu_ab_lib.sh name1 [name2 ...]      ("new" = the in-tree library)
for i in 1 2 3; do
  for v in "$@" new; do
    if [ $v = new ]; then unset FVAE_B200_LIB; else export FVAE_B200_LIB=$PWD/factorvae_b200/libfvae_$v.so; fi
    timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$v', round(j['ms_per_step'],4), round(j['roofline']['kernel_ms'],4), j['loss'])"
  done
done
for v in "$@" new; do
  if [ $v = new ]; then unset FVAE_B200_LIB; else export FVAE_B200_LIB=$PWD/factorvae_b200/libfvae_$v.so; fi
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/quick_times_$v.csv python scripts/step_traffic.py cfg2 > /dev/null 2>&1
  echo "== $v"; python scripts/kernel_times.py gpurun_out/quick_times_$v.csv 2>/dev/null | grep "gru"
done
unset FVAE_B200_LIB
