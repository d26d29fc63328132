"""Per-section error of the bf16 tensor-core chain against the fp32 CUDA-core chain (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import factorvae_b200 as fb
from factorvae_b200 import engine

dev = torch.device("cuda:0")
def cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))

for (B, N, T, H, K) in [(1, 11, 1, 16, 5), (3, 100, 5, 20, 20), (2, 100, 4, 48, 48), (2, 75, 3, 60, 60), (1, 140, 6, 64, 8), (1, 130, 2, 8, 4), (8, 300, 20, 20, 20)]:
    torch.manual_seed(11 + H)
    m = fb.FactorVAE(fb.FeatureExtractor(158, H), fb.FactorEncoder(K, 128, H), fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)), fb.FactorPredictor(H, K))
    L = engine.ParamLayout(158, H, K, 128)
    flat = L.pack(m.state_dict(), dev)
    S = B * N
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(S, T, 158, device=dev, generator=g).clamp_(-3, 3)
    y = torch.randn(S, device=dev, generator=g)
    ptr = engine.uniform_date_ptr(B, N, dev)
    res = {}
    for prec in ("fp32", "bf16"):
        out, st = engine.elbo_forward(L, flat, x, y, ptr, train=True, precision=prec, philox=(9, 1, 0))
        grad = engine.elbo_backward(L, st).clone()
        res[prec] = (out, grad, engine.latent(st))
    (o32, g32, e32), (o16, g16, e16) = res["fp32"], res["bf16"]
    print(f"--- B={B} N={N} T={T} H={H} K={K}: e maxabs {float((e16-e32).abs().max()):.3e}  loss {float(o32['loss']):.5f} vs {float(o16['loss']):.5f}"
          f"  all-grad rel {float((g16.double()-g32.double()).norm()/g32.double().norm()):.3e} cos {cos(g16,g32):.6f}")
    secs = {}
    for name in L.slices:
        sec = name.split(".")[0] + "." + ".".join(name.split(".")[1:3]) if "attention" not in name else "factor_predictor.attention"
        a, b = L.view(g16, name).double().reshape(-1), L.view(g32, name).double().reshape(-1)
        sa, sb = secs.setdefault(sec, ([], []))
        sa.append(a); sb.append(b)
    for sec, (la, lb) in secs.items():
        a, b = torch.cat(la), torch.cat(lb)
        print(f"   {sec:48s} |ref| {float(b.norm()):.3e} rel {float((a-b).norm()/(b.norm()+1e-30)):.3e} cos {cos(a,b):.5f}")
