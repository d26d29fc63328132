#!/bin/bash
# multi-GPU evidence on ONE box with N GPUs: NCCL shard-invariance test + bench lines (cfg2 weak, cfg4 / cfg5 strong)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
(timeout 600 python -m pytest tests/test_shard_invariance_gpu.py -m gpu -q -s -p no:cacheprovider -k nccl 2>&1 | tail -15) > gpurun_out/r2_shard_nccl_${N}gpu.log 2>&1
P=29500
for wl in cfg2 cfg4 cfg5; do
  P=$((P+1))
  steps=20; if [ $wl = cfg5 ]; then steps=5; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r2_bench_${wl}_${N}gpu.json 2> gpurun_out/r2_bench_${wl}_${N}gpu.err
  tail -c 600 gpurun_out/r2_bench_${wl}_${N}gpu.json | cut -c1-300
done
cat gpurun_out/r2_shard_nccl_${N}gpu.log | tail -5
