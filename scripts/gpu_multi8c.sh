#!/bin/bash
# 8 GPUs, one box: shard-invariance worker at G=8 (NCCL + NVLink kernel), cfg4 (global 512 dates, strong split), cfg2 weak scaling
N=8
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_shard_invariance_gpu.py -m gpu -q -s -p no:cacheprovider -k nccl 2>&1 | grep "G=\|passed\|failed\|Error" | head -12) > gpurun_out/r02_shard_nccl_${N}gpu.log 2>&1
cat gpurun_out/r02_shard_nccl_${N}gpu.log
P=29800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r02_bench_cfg4_${N}gpu.json 2> gpurun_out/r02_bench_cfg4_${N}gpu.err
tail -1 gpurun_out/r02_bench_cfg4_${N}gpu.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg4', j['n_gpus'], round(j['ms_per_step'],3), j['value'], j['roofline']['step']['frac'])"
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r02_bench_cfg2_${N}gpu.json 2> gpurun_out/r02_bench_cfg2_${N}gpu.err
tail -1 gpurun_out/r02_bench_cfg2_${N}gpu.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg2', j['n_gpus'], round(j['ms_per_step'],4), j['value'], j['e2e']['ms_per_step'] if j.get('e2e') else None)"
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg2 1gpu', round(j['ms_per_step'],4), j['value'])"
