#!/bin/bash
# 8 GPUs: NCCL/p2p shard-invariance worker at G=8 + cfg5 with 128-date micro-batches + cfg2 with both collectives
N=8
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_shard_invariance_gpu.py -m gpu -q -s -p no:cacheprovider -k nccl 2>&1 | grep "G=\|passed\|failed\|Error" | head -12) > gpurun_out/r2_shard_nccl_${N}gpu.log 2>&1
cat gpurun_out/r2_shard_nccl_${N}gpu.log
P=29700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r2_bench_cfg5_${N}gpu.json 2> gpurun_out/r2_bench_cfg5_${N}gpu.err
tail -1 gpurun_out/r2_bench_cfg5_${N}gpu.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg5', j['n_gpus'], round(j['ms_per_step'],3), j['value'], j['roofline']['step']['frac'])"
for coll in p2p nccl; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e --collective $coll 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg2 $coll', j['n_gpus'], round(j['ms_per_step'],4), j['value'])"
done
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg2 1gpu', round(j['ms_per_step'],4), j['value'])"
