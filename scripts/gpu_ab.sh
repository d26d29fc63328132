#!/bin/bash
# A/B on ONE box: fused backward (default) vs split-role (FVAE_BACK_SPLIT=1), 3 rounds
for i in 1 2 3; do
  for v in split fused; do
    if [ $v = split ]; then export FVAE_BACK_SPLIT=1; else unset FVAE_BACK_SPLIT; fi
    timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$v', round(j['ms_per_step'],4), round(j['roofline']['kernel_ms'],4))"
  done
done
unset FVAE_BACK_SPLIT
