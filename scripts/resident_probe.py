"""Resident-panel step alone (for ncu): cfg2-shaped dense row table, three steps through engine.IndexedWindows."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from factorvae_b200 import engine
from factorvae_b200.batched import DateShardedStep
from factorvae_b200.panel import PanelIndex, ResidentPanel
import factorvae_b200 as fb
B, N, T, H, K, C_ = 256, 300, 20, 20, 20, 158
dev = torch.device("cuda:0")
m = fb.FactorVAE(fb.FeatureExtractor(C_, H), fb.FactorEncoder(K, 128, H), fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)),
                 fb.FactorPredictor(H, K))
L = engine.ParamLayout(C_, H, K, 128)
flat = L.pack(m.state_dict(), dev)
Dn = B + T - 1
pidx = PanelIndex(np.arange(Dn * N, dtype=np.int32).reshape(Dn, N), np.repeat(np.arange(T - 1, Dn, dtype=np.int32), N),
                  np.tile(np.arange(N, dtype=np.int32), B), np.arange(0, (B + 1) * N, N), Dn * N)
vals = torch.randn(Dn * N, C_ + 1).clamp_(-3, 3).numpy()
rp = ResidentPanel(vals, pidx, C_, dev, dtype=torch.bfloat16)
stepper = DateShardedStep(L, flat, precision="bf16", seed=1)
for _ in range(3):
    xw, y, ptr = rp.batch(range(B), T)
    stepper.step(xw, y, ptr, global_dates=B, unit_base=0, train=True)
torch.cuda.synchronize()
print("loss", float(stepper.loss))
