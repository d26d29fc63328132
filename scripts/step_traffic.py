"""One ELBO step of a bench workload between cudaProfilerStart/Stop, for per-kernel ncu metrics of a whole step:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/step_traffic.csv python scripts/step_traffic.py cfg2
    python scripts/step_traffic.py --parse gpurun_out/step_traffic.csv cfg2 > profiles/r02_step_traffic.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(path, workload):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    per = {}
    order = []
    for r in rows[1:]:
        kid = int(r[ix["ID"]])
        if kid not in per:
            per[kid] = {"kernel": r[ix["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1]}
            order.append(kid)
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
        per[kid][r[ix["Metric Name"]]] = v * scale
    kernels = []
    for kid in order:
        p = per[kid]
        kernels.append({"kernel": p["kernel"], "us": p.get("gpu__time_duration.sum"), "dram_read": p.get("dram__bytes_read.sum"),
                        "dram_write": p.get("dram__bytes_write.sum")})
    import bench
    wl = bench.WORKLOADS[workload]
    S = min(wl["B"], wl["micro"]) * wl["N"]
    alg = S * (wl["T"] * bench.C_FEATURES * 2 + 4)
    tot = sum((k["dram_read"] or 0) + (k["dram_write"] or 0) for k in kernels)
    front = [k for k in kernels if "front" in k["kernel"]]
    out = {"workload": workload, "how": "ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                                       "--clock-control none, one ELBO step (forward + backward), bf16 panel resident in HBM",
           "algorithmic_bytes_per_step": alg, "dram_bytes_per_step": tot, "ratio_to_algorithmic": tot / alg,
           "sum_kernel_us_under_ncu": sum(k["us"] or 0 for k in kernels),
           "front_forward_dram_bytes_per_launch": ((front[0]["dram_read"] or 0) + (front[0]["dram_write"] or 0)) if front else None,
           "kernels": kernels}
    print(json.dumps(out, indent=1))


def main():
    if sys.argv[1] == "--parse":
        return parse(sys.argv[2], sys.argv[3])
    import torch
    import bench
    from factorvae_b200 import engine
    from factorvae_b200.batched import DateShardedStep
    name = sys.argv[1]
    wl = bench.WORKLOADS[name]
    dev = torch.device("cuda:0")
    B, N, T, H, K, M = min(wl["B"], wl["micro"]), wl["N"], wl["T"], wl["H"], wl["K"], wl["M"]
    layout = engine.ParamLayout(bench.C_FEATURES, H, K, M)
    flat = layout.pack(bench.build_params(H, K, M), dev)
    store = torch.zeros(B * N, T, 160, device=dev, dtype=torch.bfloat16)
    store[:, :, :bench.C_FEATURES] = torch.randn(B * N, T, bench.C_FEATURES, device=dev).clamp_(-3, 3).to(torch.bfloat16)
    x = store[:, :, :bench.C_FEATURES]
    y = torch.randn(B * N, device=dev)
    ptr = engine.uniform_date_ptr(B, N, dev)
    st = DateShardedStep(layout, flat, precision="bf16", seed=42)
    for _ in range(2):
        st.step(x, y, ptr, train=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    st.step(x, y, ptr, train=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", float(st.loss.item()))


if __name__ == "__main__":
    main()
