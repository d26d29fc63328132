"""Where does the end-to-end step (resident row table) lose time against the device-resident step?  Variants, each timed with
CUDA events over 30 steps: dense panel | row index prepared once | + window-index kernel per step | + row upload per step |
+ loss read-back per step (one step deferred)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from factorvae_b200 import engine
from factorvae_b200.batched import DateShardedStep
from factorvae_b200.panel import PanelIndex, ResidentPanel

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["cfg2"]
B, N, T, H, K, M = wl["B"], wl["N"], wl["T"], wl["H"], wl["K"], wl["M"]
C = bench.C_FEATURES
layout = engine.ParamLayout(C, H, K, M)
flat = layout.pack(bench.build_params(H, K, M), dev)
st = DateShardedStep(layout, flat, precision="bf16", seed=42)
store = torch.zeros(B * N, T, 160, device=dev, dtype=torch.bfloat16)
store[:, :, :C] = torch.randn(B * N, T, C, device=dev).clamp_(-3, 3).to(torch.bfloat16)
x = store[:, :, :C]
y = torch.randn(B * N, device=dev)
ptr = engine.uniform_date_ptr(B, N, dev)
Dn = B + T - 1
idx_mat = np.arange(Dn * N, dtype=np.int32).reshape(Dn, N)
pidx = PanelIndex(idx_mat, np.repeat(np.arange(T - 1, Dn, dtype=np.int32), N), np.tile(np.arange(N, dtype=np.int32), B), np.arange(0, (B + 1) * N, N), Dn * N)
vals = torch.randn(Dn * N, C + 1).clamp_(-3, 3).numpy()
rp = ResidentPanel(vals, pidx, C, dev, dtype=torch.bfloat16)
xw, yw, pw = rp.batch(range(B), T)
rows_h = rp.table[(T - 1) * N:(T - 1) * N + B * N].to("cpu").pin_memory()
lab_h = rp.label[(T - 1) * N:(T - 1) * N + B * N].to("cpu").pin_memory()
loss_h = torch.empty(1).pin_memory()
copy_stream = torch.cuda.Stream(dev)

def timeit(name, fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t_issue = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"{name:58s} {e0.elapsed_time(e1) / n:7.3f} ms/step   (host issue {t_issue:6.3f} ms/step)")

timeit("dense panel, device resident", lambda: st.step(x, y, ptr, train=True))
timeit("row table, index prepared once", lambda: st.step(xw, yw, pw, train=True))
def with_index():
    a, b, c = rp.batch(range(B), T)
    st.step(a, b, c, train=True)
timeit("row table + window-index kernel per step", with_index)
def with_upload():
    with torch.cuda.stream(copy_stream):
        rp.upload_rows((T - 1) * N, rows_h, lab_h)
    with_index()
timeit("  + rows uploaded per step (copy stream, unsynchronised)", with_upload)
ev = torch.cuda.Event()
def with_readback():
    with_upload()
    loss_h.copy_(st.loss.reshape(1), non_blocking=True)
    ev.record()
    ev.synchronize()
timeit("  + loss read on the host every step (synchronous)", with_readback)
def upload_only():
    with torch.cuda.stream(copy_stream):
        rp.upload_rows((T - 1) * N, rows_h, lab_h)
timeit("upload alone (24.9 MB H2D)", upload_only)

# ---- the bench's own loop (double-buffered tables, events, one-step-deferred loss read)
tables = [rp, ResidentPanel(vals, pidx, C, dev, dtype=torch.bfloat16)]
first_new = (T - 1) * N
compute = torch.cuda.current_stream(dev)
uploaded = [torch.cuda.Event(), torch.cuda.Event()]
consumed = [torch.cuda.Event(), torch.cuda.Event()]
done = [torch.cuda.Event(), torch.cuda.Event()]
loss_hh = [torch.empty(1).pin_memory() for _ in range(2)]
def upload(slot):
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(consumed[slot])
        tables[slot].upload_rows(first_new, rows_h, lab_h)
        uploaded[slot].record(copy_stream)
def run_resident(nsteps, defer=True, do_upload=True):
    consumed[0].record(compute); consumed[1].record(compute)
    if do_upload: upload(0)
    for i in range(nsteps):
        slot = i & 1
        if do_upload and i + 1 < nsteps: upload(slot ^ 1)
        if do_upload: compute.wait_event(uploaded[slot])
        a, b, c = tables[slot].batch(range(B), T)
        st.step(a, b, c, train=True)
        consumed[slot].record(compute)
        loss_hh[slot].copy_(st.loss.reshape(1), non_blocking=True)
        done[slot].record(compute)
        if defer and i >= 1:
            done[slot ^ 1].synchronize()
    done[(nsteps - 1) & 1].synchronize()
for name, kw in (("bench loop, no upload, deferred read", dict(do_upload=False)), ("bench loop, upload, deferred read", dict()), ("bench loop, upload, no per-step read", dict(defer=False))):
    run_resident(5, **kw)
    torch.cuda.synchronize()
    for n in (10, 40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_resident(n, **kw); e1.record(); torch.cuda.synchronize()
        print(f"{name:44s} n={n:3d}  {e0.elapsed_time(e1) / n:7.3f} ms/step")

# ---- the date-id H2D of the bench's resident loop: on the compute stream vs prefetched on the copy stream
dates_h = torch.arange(B, dtype=torch.int32).pin_memory()
dates_d = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2)]
ids_up = [torch.cuda.Event(), torch.cuda.Event()]
def run_ids(nsteps, mode):
    if mode == "copy_stream":
        with torch.cuda.stream(copy_stream):
            dates_d[0].copy_(dates_h, non_blocking=True); ids_up[0].record(copy_stream)
    for i in range(nsteps):
        b = i & 1
        if mode == "compute":
            dates_d[0].copy_(dates_h, non_blocking=True)
        elif mode == "copy_stream":
            if i + 1 < nsteps:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done[b ^ 1])      # the step that used this id buffer has finished
                    dates_d[b ^ 1].copy_(dates_h, non_blocking=True); ids_up[b ^ 1].record(copy_stream)
            compute.wait_event(ids_up[b])
        a, bb, c = tables[0].batch(range(B), T)
        st.step(a, bb, c, train=True)
        loss_hh[b].copy_(st.loss.reshape(1), non_blocking=True)
        done[b].record(compute)
        if i >= 1:
            done[b ^ 1].synchronize()
    done[(nsteps - 1) & 1].synchronize()
for mode in ("none", "compute", "copy_stream"):
    done[0].record(compute); done[1].record(compute)
    run_ids(8, mode)
    torch.cuda.synchronize()
    for n in (20, 40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_ids(n, mode); e1.record(); torch.cuda.synchronize()
        print(f"resident loop, date ids H2D: {mode:12s} n={n:3d}  {e0.elapsed_time(e1) / n:7.3f} ms/step")
