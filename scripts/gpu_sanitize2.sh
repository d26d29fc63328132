#!/bin/bash
# compute-sanitizer over what round 2 added late: sub-module calls (fvae_heads_parts), the cluster form of the fp32 heads forward
# (forced cluster sizes), graph capture with the device step counter, and the packed-math GRU / front kernels at a cfg2 shape
mkdir -p gpurun_out
(FVAE_HEADS_CLUSTER=4 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_submodules_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "alone or golden or fp32" 2>&1 | tail -12) > gpurun_out/r02_sanitizer_memcheck_cluster_parts.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_graph_step_gpu.py tests/test_parity_bench_config_gpu.py -m gpu -q -p no:cacheprovider -x -k "graph or (bench_shapes and cfg2)" 2>&1 | tail -12) > gpurun_out/r02_sanitizer_memcheck_graph_gru.log 2>&1
(FVAE_HEADS_CLUSTER=8 timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_submodules_gpu.py -m gpu -q -p no:cacheprovider -x -k "alone" 2>&1 | tail -8) > gpurun_out/r02_sanitizer_synccheck_cluster.log 2>&1
tail -5 gpurun_out/r02_sanitizer_memcheck_cluster_parts.log; tail -5 gpurun_out/r02_sanitizer_memcheck_graph_gru.log; tail -4 gpurun_out/r02_sanitizer_synccheck_cluster.log
