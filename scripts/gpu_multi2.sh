#!/bin/bash
# 2-GPU: NCCL/p2p shard-invariance worker + bench cfg2 with both collectives + 1-GPU reference on the same box
N=${1:-2}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_shard_invariance_gpu.py -m gpu -q -s -p no:cacheprovider -k nccl 2>&1 | grep "G=\|passed\|failed\|Error" | head -12) > gpurun_out/r2_shard_nccl_${N}gpu.log 2>&1
cat gpurun_out/r2_shard_nccl_${N}gpu.log
P=29600
for coll in p2p nccl p2p nccl; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e --collective $coll 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$coll', j['n_gpus'], round(j['ms_per_step'],4), j['detail']['collective'][-40:])"
done
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('1gpu', round(j['ms_per_step'],4))"
