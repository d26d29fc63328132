"""Worker of tests/test_shard_invariance_gpu.py: launched under torch.distributed.run with G ranks (one GPU each, NCCL).
Every rank runs its contiguous block of the SAME global batch through DateShardedStep (one gradient all-reduce); rank 0 also
runs the whole batch alone (G = 1) and checks: all-reduced gradient == G=1 gradient to fp32 round-off, loss identical, and
the per-unit outputs of its shard bit-identical to the corresponding slice of the G=1 run (noise keyed by global unit id)."""
import datetime
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_batch(dev, counts, T, C=158):
    xs, ys = [], []
    gen = torch.Generator(device=dev)
    for d, n in enumerate(counts):
        gen.manual_seed(4321 + d)                       # keyed by the GLOBAL date id
        xs.append(torch.randn(n, T, C, generator=gen, device=dev).clamp_(-3, 3).to(torch.bfloat16))
        ys.append(torch.randn(n, generator=gen, device=dev))
    return xs, ys


def main():
    from factorvae_b200 import engine
    from factorvae_b200.batched import DateShardedStep, shard_dates
    import factorvae_b200 as fb
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    H = K = 20
    T = 8
    counts = [300, 257, 128, 301, 64, 299, 300, 190][: max(world * 2, 4)]
    if len(counts) % world:
        counts = counts[: len(counts) - len(counts) % world]
    B = len(counts)
    torch.manual_seed(42)
    m = fb.FactorVAE(fb.FeatureExtractor(158, H), fb.FactorEncoder(K, 128, H), fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)),
                     fb.FactorPredictor(H, K))
    L = engine.ParamLayout(158, H, K, 128)
    flat = L.pack(m.state_dict(), dev)
    xs, ys = make_batch(dev, counts, T)
    d0, d1 = shard_dates(B, world, rank)
    ptr = torch.tensor([0] + list(torch.tensor(counts[d0:d1]).cumsum(0)), dtype=torch.int32, device=dev)
    base = sum(counts[:d0])
    st = DateShardedStep(L, flat, precision=precision, seed=11, collective="p2p")     # the one-kernel NVLink all-reduce
    for _ in range(3):                               # several epochs through the double-buffered slots
        st.step_index = 0
        out, _ = st.step(torch.cat(xs[d0:d1]), torch.cat(ys[d0:d1]), ptr, global_dates=B, unit_base=base, train=True)
    torch.cuda.synchronize()
    g_all = st.grad.clone()
    loss_all = float(st.loss.item())
    # the same step through NCCL: the two collectives must agree (both sum the same shard gradients)
    st2 = DateShardedStep(L, flat, precision=precision, seed=11, collective="nccl")
    st2.step(torch.cat(xs[d0:d1]), torch.cat(ys[d0:d1]), ptr, global_dates=B, unit_base=base, train=True)
    torch.cuda.synchronize()
    d_coll = float((st2.grad.double() - g_all.double()).norm() / g_all.double().norm())
    # bit-identical on every rank (rank-ordered sum): compare a checksum of the bytes across ranks
    chk = g_all.view(torch.int32).to(torch.int64).sum().reshape(1)
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    same_everywhere = bool((lo_ == hi_).item())
    ok = True
    if rank == 0:
        solo = DateShardedStep(L, flat, precision=precision, group=None, seed=11, collective="nccl")   # no collective at construction
        solo.world = 1                                   # the whole batch on this GPU, no collective
        pall = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=dev)
        o1, _ = solo.step(torch.cat(xs), torch.cat(ys), pall, global_dates=B, unit_base=0, train=True)
        torch.cuda.synchronize()
        g1, l1 = solo.grad.double(), float(solo.loss.item())
        rel = float((g_all.double() - g1).norm() / g1.norm())
        mx = float((g_all.double() - g1).abs().max() / g1.abs().max())
        lrel = abs(loss_all - l1) / abs(l1)
        n0 = sum(counts[d0:d1])
        # same draws (noise keyed by the global unit id), equal to fp32 round-off: the order of the partial sums over a date's
        # stocks follows the launch geometry, which follows the number of dates in the call
        same = (torch.allclose(out["yhat"], o1["yhat"][base:base + n0], rtol=1e-5, atol=1e-6) and
                torch.allclose(out["mu_y"], o1["mu_y"][base:base + n0], rtol=1e-5, atol=1e-6))
        print(f"G={world} {precision}: grad rel-L2 {rel:.3e} max-rel {mx:.3e}; loss {loss_all:.7f} vs {l1:.7f} (rel {lrel:.2e}); "
              f"per-unit outputs equal to round-off: {same}; p2p vs nccl rel-L2 {d_coll:.2e}; p2p result identical on all ranks: "
              f"{same_everywhere}; p2p active: {st.p2p is not None}", flush=True)
        # fp32 round-off: the weight-gradient sums are float atomics (order free), bf16 mode adds the tensor-core heads' own
        tol = 2e-6 if precision == "fp32" else 1e-5
        ok = rel <= tol and mx <= 2e-5 and lrel <= 1e-6 and same and d_coll <= tol and same_everywhere and st.p2p is not None
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
