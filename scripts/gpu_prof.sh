#!/bin/bash
# per-kernel times + DRAM bytes of one cfg2 step, and a full ncu capture of one kernel (regex $1)
K=${1:-tc_back_tma}
TAG=${2:-r2g}
mkdir -p gpurun_out
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_step_traffic.csv python scripts/step_traffic.py cfg2 > gpurun_out/${TAG}_traffic.log 2>&1
python scripts/step_traffic.py --parse gpurun_out/${TAG}_step_traffic.csv cfg2 > gpurun_out/${TAG}_step_traffic.json 2>> gpurun_out/${TAG}_traffic.log
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$K --launch-count 1 -o gpurun_out/${TAG}_${K} -f python scripts/step_traffic.py cfg2 > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_step_traffic.json"))
print(d["dram_bytes_per_step"]/1e9, d["ratio_to_algorithmic"], d["sum_kernel_us_under_ncu"])
for k in d["kernels"]:
    print("%-34s %8.1f us rd %8.1f MB wr %8.1f MB" % (k["kernel"][:34], k["us"], k["dram_read"]/1e6, k["dram_write"]/1e6))
PY
