#!/bin/bash
# compute-sanitizer over the round-2 kernels (TMA front forward, fused / split front backward, gather4 path, p2p not included: 1 GPU)
mkdir -p gpurun_out
SEL='bench_shapes and cfg2 or index or replay or more_tiles'
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_parity_bench_config_gpu.py tests/test_panel_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "$SEL" 2>&1 | tail -25) > gpurun_out/r02_sanitizer_memcheck.log 2>&1
(FVAE_BACK_SPLIT=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_parity_bench_config_gpu.py -m gpu -q -p no:cacheprovider -x -k "bench_shapes and cfg2" 2>&1 | tail -12) > gpurun_out/r02_sanitizer_memcheck_split.log 2>&1
(timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_parity_bench_config_gpu.py -m gpu -q -p no:cacheprovider -x -k "bench_shapes and cfg2" 2>&1 | tail -12) > gpurun_out/r02_sanitizer_synccheck.log 2>&1
tail -6 gpurun_out/r02_sanitizer_memcheck.log; tail -4 gpurun_out/r02_sanitizer_memcheck_split.log; tail -4 gpurun_out/r02_sanitizer_synccheck.log
