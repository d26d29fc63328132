"""The ELBO step captured in a CUDA graph (VERDICT r1 small item 8, SURVEY hard-part 10: BASELINE.json configs[0] -- one date
of 64 stocks -- is ~15 launches of a few microseconds, i.e. launch-bound).  `DateShardedStep.capture` records forward + backward
once; every replay must be the step `DateShardedStep.step` would have run at the same step index: same Philox draws (the counter
lives in device memory, fvae_noise.step_dev, and the graph advances it), same loss, same gradient -- for refreshed inputs and
refreshed parameters too."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(H, K, dev, seed=42):
    import factorvae_b200 as fb
    from factorvae_b200 import engine
    torch.manual_seed(seed)
    m = fb.FactorVAE(fb.FeatureExtractor(158, H), fb.FactorEncoder(K, 128, H), fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)),
                     fb.FactorPredictor(H, K))
    L = engine.ParamLayout(158, H, K, 128)
    return L, L.pack(m.state_dict(), dev)


@pytest.mark.parametrize("precision,counts,T", [("bf16", [64], 20), ("fp32", [64], 20), ("bf16", [300, 257, 128], 8)])
def test_graph_replay_equals_the_eager_step(precision, counts, T, cuda_device):
    from factorvae_b200.batched import DateShardedStep
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _shard_worker import make_batch
    dev = cuda_device
    L, flat = _model(20, 20, dev)
    xs, ys = make_batch(dev, counts, T)
    x, y = torch.cat(xs).contiguous(), torch.cat(ys).contiguous()
    ptr = torch.tensor([0] + counts).cumsum(0).to(torch.int32).to(dev)
    tol = 2e-6 if precision == "fp32" else 1e-5

    graphed = DateShardedStep(L, flat, precision=precision, seed=7)
    g = graphed.capture(x, y, ptr, train=True)
    first = graphed.step_index                       # the warm-up passes consumed step indices; replays continue from here
    eager = DateShardedStep(L, flat, precision=precision, seed=7)
    eager.step_index = first

    def same():
        out_g = g.replay()
        out_e, _ = eager.step(x, y, ptr, train=True)
        assert graphed.step_index == eager.step_index
        assert torch.equal(out_g["yhat"], out_e["yhat"])                                  # same eps, same dropout masks
        le, lg = float(eager.loss.item()), float(graphed.loss.item())
        assert abs(lg - le) <= 1e-6 * abs(le), (lg, le)
        ge, gg = eager.grad.double(), graphed.grad.double()
        assert float((gg - ge).norm() / ge.norm()) <= tol
        return out_g["yhat"].clone(), lg

    y1, l1 = same()
    y2, l2 = same()
    assert not torch.equal(y1, y2)                                                        # the step counter advanced: a fresh draw
    x.copy_(torch.roll(x, 1, dims=0))                                                     # refill the static input in place
    _, l3 = same()
    assert l3 != l2
    flat.mul_(0.97)                                                                       # an optimizer stepping the parameters in place
    _, l4 = same()
    assert l4 != l3 and torch.isfinite(torch.tensor(l4))


def test_graph_capture_is_single_gpu_only_and_counts_one_launch_per_replay(cuda_device):
    from factorvae_b200 import _cabi
    from factorvae_b200.batched import DateShardedStep
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _shard_worker import make_batch
    dev = cuda_device
    L, flat = _model(20, 20, dev)
    xs, ys = make_batch(dev, [64], 20)
    ptr = torch.tensor([0, 64], dtype=torch.int32, device=dev)
    st = DateShardedStep(L, flat, precision="bf16", seed=3)
    g = st.capture(xs[0], ys[0], ptr)
    before = _cabi.lib().fvae_debug_launch_count()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    assert _cabi.lib().fvae_debug_launch_count() == before         # no launch goes through the host-side launch path on replay
    st.world = 2
    with pytest.raises(NotImplementedError):
        st.capture(xs[0], ys[0], ptr)
