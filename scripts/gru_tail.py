"""Tail-effect probe: FeatureExtractor forward+backward at 592 vs 600 tiles (run under an ncu launch list)."""
import sys, torch
sys.path.insert(0, ".")
from factorvae_b200 import engine
import factorvae_b200 as fb
H = 20
m = fb.FactorVAE(fb.FeatureExtractor(158, H), fb.FactorEncoder(20, 128, H), fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, 20)),
                 fb.FactorPredictor(H, 20))
L = engine.ParamLayout(158, H, 20, 128)
dev = torch.device("cuda:0")
flat = L.pack(m.state_dict(), dev)
for tiles in (592, 600, 592, 600):
    S = tiles * 128
    x = torch.randn(S, 20, 158, device=dev).clamp_(-3, 3).to(torch.bfloat16)
    e, st = engine.fe_forward(L, flat, x, "bf16")
    g = engine.fe_backward(L, st, torch.randn_like(e))
    torch.cuda.synchronize()
    print("tiles", tiles, "done")
