"""bench.py -- dates x stocks / second per ELBO step (forward + backward [+ gradient all-reduce]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (FactorVAE.forward + backward, reference module.py:250-270 +
train_model.py:29) over one batch of synthetic dates.  Workload at N=1: BASELINE.json configs[1]
(B=256 dates x N=300 stocks x T=20 x C=158, K=H=20, M=128).  For N>1 the per-GPU work is fixed
(weak scaling): every rank processes B dates of the same shape, dates keyed by their global id.

Prints ONE JSON line (rank 0).  `value` = whole-job units/s with the panel resident in HBM;
`e2e` = the same through the host-buffer entry (H2D of the batch + D2H of the loss inside the
timed region); `roofline` / `cpu_baseline` as specified in the task statement.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B dates per GPU, N stocks, T, K=H, M)
    "cfg1": dict(B=1, N=64, T=20, H=20, K=20, M=128),
    "cfg2": dict(B=256, N=300, T=20, H=20, K=20, M=128),
    "cfg3": dict(B=256, N=500, T=60, H=60, K=60, M=128),
    "cfg4": dict(B=64, N=1000, T=20, H=48, K=48, M=128),     # per-GPU share of B=512 over 8 GPUs
    "cfg5": dict(B=128, N=3000, T=60, H=60, K=60, M=128),    # per-GPU share of B=1024 over 8 GPUs
}
C_FEATURES = 158
METRIC = "dates x stocks / sec per ELBO step (fwd+bwd), K=20 C=158"


def f_fe(T, H, C=C_FEATURES):
    """Algorithmic FLOPs per date x stock per step: FeatureExtractor contractions, fwd + 2x bwd."""
    return 3 * 2 * T * (C * C + 3 * H * C + 3 * H * H)


def build_params(H, K, M, seed=42):
    import torch
    import factorvae_b200 as fb
    torch.manual_seed(seed)
    model = fb.FactorVAE(fb.FeatureExtractor(C_FEATURES, H), fb.FactorEncoder(K, M, H),
                         fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)), fb.FactorPredictor(H, K))
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING warm-up + the timed region (one `nvidia-smi -lms 100`
    child process; stopped by its exact PID)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in ln.split(",")] for ln in out.strip().splitlines() if ln.count(",") >= 8]
        busy = [r for r in rows if r[3].replace(".", "").isdigit() and float(r[3]) > 250.0] or rows      # samples under load
        sm = sorted(float(r[1]) for r in busy if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows), "samples_under_load": len(busy)}


def run_reference(args, wl, rank, world):
    """Reference arm: the reference's CPU implementation of the path (oracle/cpu_port.py -- the Python
    reference cannot travel to the GPU box) on the host cores, one date per step as train_model.py does."""
    if rank != 0:
        return
    import torch
    from oracle.cpu_port import CpuPort, pick_threads
    params = build_params(wl["H"], wl["K"], wl["M"])
    port = CpuPort(params)
    g = torch.Generator().manual_seed(1234)
    # bounded sample: each step = `dates_per_step` per-date reference steps of the workload's (N, T, K)
    dates_per_step = 4
    xs = [torch.randn(wl["N"], wl["T"], C_FEATURES, generator=g).clamp_(-3, 3) for _ in range(dates_per_step)]
    ys = [torch.randn(wl["N"], 1, generator=g) for _ in range(dates_per_step)]
    cores = pick_threads(port, xs[0], ys[0])      # fastest intra-op width on this host (of %d cpus)
    for _ in range(args.warmup):
        for x, y in zip(xs, ys):
            port.train_step(x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for x, y in zip(xs, ys):
            port.train_step(x, y)
    dt = time.perf_counter() - t0
    units = args.steps * dates_per_step * wl["N"]
    value = units / dt
    sample = (f"{dates_per_step} dates/step of the workload's per-date shape (N={wl['N']},T={wl['T']},K=H={wl['K']}), one date per "
              f"reference step (zero_grad, forward, loss.item(), backward; no optimizer), fp32 torch CPU, {cores} threads (fastest of 1..{os.cpu_count()} probed)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "date*stocks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload_desc},
            "cpu_baseline": {"value": value, "unit": "date*stocks/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "date*stocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--panel", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-resident", action="store_true", help="skip the resident-panel variant of the end-to-end measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.workload_desc = (f"{args.workload}: B={wl['B']} dates/GPU x N={wl['N']} stocks x T={wl['T']} x C={C_FEATURES}, "
                          f"K=H={wl['K']}, M={wl['M']}; {world} GPU(s), dates sharded, weak scaling")
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)

    import torch
    import torch.distributed as dist
    from factorvae_b200 import _cabi, engine
    from factorvae_b200.batched import DateShardedStep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a stuck rendezvous / collective must end as an error within minutes, never as a hung box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    B, N, T, H, K, M = (wl[k] for k in "BNTHKM")
    S = B * N
    precision = args.precision
    if precision == "auto":
        precision = "bf16" if engine.tc_supported(C_FEATURES, H) else "fp32"
    params = build_params(H, K, M)
    layout = engine.ParamLayout(C_FEATURES, H, K, M)
    flat = layout.pack(params, dev)

    # synthetic panel: N(0,1) clipped to +-3 per GLOBAL date id (identical global batch for any sharding)
    pdt = torch.bfloat16 if args.panel == "bf16" else torch.float32
    x = torch.empty(S, T, C_FEATURES, dtype=pdt, device=dev)
    y = torch.empty(S, dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    for d in range(B):
        gen.manual_seed(1234 + rank * B + d)
        x[d * N:(d + 1) * N] = torch.randn(N, T, C_FEATURES, generator=gen, device=dev).clamp_(-3, 3).to(pdt)
        y[d * N:(d + 1) * N] = torch.randn(N, generator=gen, device=dev)
    date_ptr = engine.uniform_date_ptr(B, N, dev)
    stepper = DateShardedStep(layout, flat, precision=precision, seed=42)
    unit_base = rank * S
    lib = _cabi.lib()

    def one_step():
        stepper.step(x, y, date_ptr, global_dates=B * world, unit_base=unit_base, train=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    t_warm = time.perf_counter()
    for _ in range(args.warmup):
        one_step()
    # keep the GPU under load for ~0.5 s so the clock sampler sees a few samples.  The number of extra steps must be the
    # SAME on every rank (each step carries an all-reduce): agree on it with one collective instead of a per-rank clock.
    torch.cuda.synchronize()
    t_probe = time.perf_counter()
    for _ in range(10):                                            # fixed count: identical on every rank
        one_step()
    torch.cuda.synchronize()
    per_step = max((time.perf_counter() - t_probe) / 10.0, 1e-5)
    extra = torch.tensor([int(min(5000, max(0.0, 0.5 - (time.perf_counter() - t_warm)) / per_step))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(extra, op=dist.ReduceOp.MAX)
    for _ in range(int(extra.item())):
        one_step()
    barrier()
    l0 = lib.fvae_debug_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    barrier()
    launches = lib.fvae_debug_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    loss_val = float(stepper.loss.item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * S / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel, timed ALONE with CUDA events on the launching stream
    # dominant kernel = tc_front_fwd_kernel (K1: LayerNorm -> GEMM 128x160x160 -> LeakyReLU -> GEMM 128xNCx160 per item);
    # algorithmic work per launch: FLOPs = S*T*2*(C^2 + 3HC); bytes = one read of the bf16 panel (S*T*C*2).
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, 6.65 TB/s)"
    flops_step = S * f_fe(T, H)
    step_tf = flops_step / (ms_per_step * 1e-3) / 1e12
    roofline = None
    if rank == 0 and precision == "bf16":
        out_k, st_k = engine.elbo_forward(layout, flat, x, y, date_ptr, train=True, precision="bf16", philox=(42, 1, unit_base))
        torch.cuda.synchronize()
        reps = max(5, args.steps)
        for _ in range(3):
            engine.rerun_front_forward(st_k)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            engine.rerun_front_forward(st_k)
        k1.record()
        torch.cuda.synchronize()
        k_ms = k0.elapsed_time(k1) / reps
        k_flops = S * T * 2.0 * (C_FEATURES * C_FEATURES + 3 * H * C_FEATURES)
        k_bytes = S * T * C_FEATURES * (2 if args.panel == "bf16" else 4)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        ach = k_flops / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tc_front_fwd_kernel", "achieved": ach, "peak": peak_burst, "unit": "TFLOP/s",
                    "frac": ach / peak_burst, "traffic": traffic, "peak_source": peak_src + ", burst figure (kernel timed alone)",
                    "kernel_ms": k_ms, "algorithmic_flops_per_launch": k_flops, "algorithmic_bytes_per_launch": k_bytes,
                    "hbm_algorithmic_gbs": k_bytes / (k_ms * 1e-3) / 1e9, "hbm_frac_algorithmic": k_bytes / (k_ms * 1e-3) / 1e9 / peak_hbm,
                    "step": {"achieved": step_tf, "peak": peak_sust, "frac": step_tf / peak_sust,
                             "note": "whole ELBO step (all kernels): S*3*2T(C^2+3HC+3H^2) algorithmic FLOPs / step time vs sustained bf16 peak"}}
        del out_k, st_k
    elif rank == 0:
        roofline = {"bound": "tensor", "kernel": "whole step (fp32 CUDA-core mode)", "achieved": step_tf, "peak": peak_sust,
                    "unit": "TFLOP/s", "frac": step_tf / peak_sust, "traffic": None, "peak_source": peak_src}

    # ---- end to end through the host-buffer entry: pinned host panel -> H2D -> step -> D2H of the loss, every step.
    # Primary number: fp32 host panel (what the reference's loader yields, train_model.py:23); the H2D copy of step i+1
    # overlaps the compute of step i.  Secondary: the same with the panel kept in bf16 on the host.
    e2e = None
    if not args.no_e2e:
        ph = date_ptr.to("cpu").pin_memory()
        yh = y.to("cpu").pin_memory()
        n_e2e = max(3, args.steps // 2)
        kw = dict(global_dates=B * world, unit_base=unit_base, train=True)

        def run_e2e(xh):
            for _ in stepper.run_from_host([(xh, yh, ph)] * 3, **kw):
                pass
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in stepper.run_from_host([(xh, yh, ph)] * n_e2e, **kw):
                pass
            b1.record()
            barrier()
            t2 = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            return float(t2.item()) / n_e2e

        xh32 = x.float().to("cpu").pin_memory()
        ms32 = run_e2e(xh32)
        h2d32 = xh32.numel() * xh32.element_size() + yh.numel() * 4 + ph.numel() * 4
        del xh32
        xh16 = x.to(torch.bfloat16).to("cpu").pin_memory()
        ms16 = run_e2e(xh16)
        h2d16 = xh16.numel() * xh16.element_size() + yh.numel() * 4 + ph.numel() * 4
        del xh16
        e2e = {"value": world * S / (ms32 * 1e-3), "unit": "date*stocks/s", "ms_per_step": ms32, "h2d_bytes_per_step": h2d32,
               "d2h_bytes_per_step": 4, "host_panel_dtype": "float32", "h2d_gbs": h2d32 / (ms32 * 1e-3) / 1e9,
               "overlap": "H2D of step i+1 on a copy stream under the compute of step i",
               "bf16_host_panel": {"value": world * S / (ms16 * 1e-3), "ms_per_step": ms16, "h2d_bytes_per_step": h2d16}}

    # ---- the same step fed from a RESIDENT row table: the (date, instrument) rows are uploaded once (not
    # timed, like the reference's one-time pickle load); per step the host sends only the batch's date numbers, one kernel
    # builds the look-back row index (TSDataSampler._get_indices) and the ELBO kernels read the rows in place.
    if e2e is not None and not args.no_resident:
        import numpy as np
        from factorvae_b200.panel import PanelIndex, ResidentPanel
        Dn = B + T - 1
        idx_mat = np.arange(Dn * N, dtype=np.int32).reshape(Dn, N)
        sd = np.repeat(np.arange(T - 1, Dn, dtype=np.int32), N)
        sj = np.tile(np.arange(N, dtype=np.int32), B)
        pidx = PanelIndex(idx_mat, sd, sj, np.arange(0, (B + 1) * N, N), Dn * N)
        grow = torch.Generator(device="cpu").manual_seed(99 + rank)
        vals = torch.randn(Dn * N, C_FEATURES + 1, generator=grow).clamp_(-3, 3).numpy()
        rp = ResidentPanel(vals, pidx, C_FEATURES, dev, dtype=pdt)
        dates_h = torch.arange(B, dtype=torch.int32).pin_memory()
        dates_d = torch.empty(B, dtype=torch.int32, device=dev)
        loss_h = torch.empty(1, dtype=torch.float32).pin_memory()

        def resident_step():
            dates_d.copy_(dates_h, non_blocking=True)                      # the step's host input: which dates
            xw, yw, pw = rp.batch(range(B), T)                            # window-index kernel (+ labels)
            stepper.step(xw, yw, pw, global_dates=B * world, unit_base=unit_base, train=True)
            loss_h.copy_(stepper.loss.reshape(1), non_blocking=True)      # D2H of the loss
        for _ in range(3):
            resident_step()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(n_e2e):
            resident_step()
        r1.record()
        barrier()
        t3 = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        msr = float(t3.item()) / n_e2e
        e2e["resident_panel"] = {"value": world * S / (msr * 1e-3), "ms_per_step": msr, "h2d_bytes_per_step": B * 4 + (B + 1) * 4,
                                 "d2h_bytes_per_step": 4, "table_mb": rp.table.numel() * rp.table.element_size() / 1e6,
                                 "note": "row table uploaded once (untimed); per step: date ids H2D, fvae_window_index, "
                                         "ELBO step reading rows through fvae_panel.row_index, loss D2H"}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle.cpu_port import time_cpu_steps
        v, ms_date, nsteps, cores = time_cpu_steps(params, N, T, C_FEATURES, budget_s=10.0)
        cpu_baseline = {"value": v, "unit": "date*stocks/s", "cores": cores, "kind": "port",
                        "sample": f"{nsteps} per-date reference-style steps (N={N},T={T},K=H={K}; zero_grad, forward, "
                                  f"loss.item(), backward) in ~10 s, median {ms_date:.2f} ms/date, fp32 torch CPU"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "date*stocks/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload_desc, "panel_dtype": args.panel, "precision": precision,
                           "l2": "inputs larger than L2 (panel %.0f MB per GPU)" % (x.numel() * x.element_size() / 1e6),
                           "noise": "in-kernel Philox (eps + dropout masks), keyed by global unit id",
                           "parallelism": f"dp{world} over dates"},
                "loss": loss_val, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "e2e": e2e}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
