"""Resident panel on the device: the window-index kernel and the gather against the reference's
windows (golden, bit-exact), and the FeatureExtractor / ELBO kernels reading rows through the index against the same
kernels on the materialised windows (bit-identical)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "panel_windows.npz")


def _frame(g):
    idx = pd.MultiIndex.from_arrays([pd.to_datetime(g["datetime_ns"]), g["instrument"].astype(str)], names=["datetime", "instrument"])
    cols = [f"F{k}" for k in range(g["values"].shape[1] - 1)] + ["LABEL0"]
    return pd.DataFrame(g["values"], index=idx, columns=cols).sample(frac=1.0, random_state=5)


@pytest.mark.parametrize("fill", ["none", "ffill", "ffill+bfill"])
def test_window_index_and_gather_match_reference_windows(fill, cuda_device):
    from factorvae_b200.panel import ResidentPanel
    g = np.load(GOLD)
    T, Cf = int(g["T"]), g["values"].shape[1] - 1
    rp = ResidentPanel.from_dataframe(_frame(g), Cf, cuda_device, start=pd.Timestamp(int(g["start_ns"])),
                                      end=pd.Timestamp(int(g["end_ns"])), dtype=torch.float32)
    xw, y, date_ptr = rp.batch(range(rp.num_batches), T, fill)
    ref = g["windows_" + fill.replace("+", "_")]
    assert date_ptr.cpu().tolist() == [0] + list(np.cumsum(g["counts_" + fill.replace("+", "_")]))
    # row numbers: device kernel == CPU restatement (itself pinned to the reference in test_panel_cpu.py)
    rows = rp.index.window_rows(np.arange(len(rp.index.sample_date)), T, fill)
    assert np.array_equal(xw.row_index.cpu().numpy().astype(np.int64), rows)
    w = rp.windows(xw, torch.float32).cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isnan(w), np.isnan(ref[:, :, :Cf]))
    assert np.array_equal(np.nan_to_num(w), np.nan_to_num(ref[:, :, :Cf]))       # values are fp32-representable: bit-exact
    yr = ref[:, -1, -1]
    yg = y.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isnan(yg), np.isnan(yr)) and np.array_equal(np.nan_to_num(yg), np.nan_to_num(yr))
    # a shuffled epoch is a permutation of the date batches
    perm = [3, 0, 7, 2]
    xs, ys, ps = rp.batch(perm, T, fill)
    dp = rp.index.date_ptr
    sel = np.concatenate([np.arange(dp[d], dp[d + 1]) for d in perm])
    assert np.array_equal(xs.row_index.cpu().numpy(), rows[sel].astype(np.int32))
    assert ps.cpu().tolist() == [0] + list(np.cumsum([dp[d + 1] - dp[d] for d in perm]))


@pytest.mark.parametrize("prec,dtype", [("fp32", torch.float32), ("bf16", torch.bfloat16), ("bf16", torch.float32)])
def test_kernels_read_rows_through_the_index(prec, dtype, cuda_device):
    """ELBO forward + backward on engine.IndexedWindows == the same kernels on the materialised windows."""
    from factorvae_b200 import engine
    from factorvae_b200.panel import ResidentPanel
    import factorvae_b200 as fb
    rng = np.random.default_rng(3)
    D, I, Cf, T, H, K = 12, 150, 158, 6, 20, 20
    dates = pd.date_range("2021-03-01", periods=D, freq="D")
    rows = [(d, f"S{j:04d}") for d in dates for j in range(I) if rng.random() > 0.15]
    idx = pd.MultiIndex.from_tuples(rows, names=["datetime", "instrument"])
    vals = np.clip(rng.standard_normal((len(rows), Cf + 1)), -3, 3).astype(np.float32)
    df = pd.DataFrame(vals, index=idx)
    rp = ResidentPanel.from_dataframe(df, Cf, cuda_device, start=dates[4], dtype=dtype)
    xw, y, date_ptr = rp.batch(range(rp.num_batches), T)
    x = rp.windows(xw, dtype)                         # what the reference's DataLoader would hand over
    torch.manual_seed(2)
    m = fb.FactorVAE(fb.FeatureExtractor(Cf, H), fb.FactorEncoder(K, 128, H),
                     fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)), fb.FactorPredictor(H, K))
    L = engine.ParamLayout(Cf, H, K, 128)
    flat = L.pack(m.state_dict(), cuda_device)
    # the same windows with a 16-byte row pitch: in bf16 tensor-core mode a bf16 panel then takes the TMA kernels, exactly
    # like the row table (box loads there, gather4 here) -- identical arithmetic; the contiguous pitch-158 tensor takes the
    # cp.async kernels (LayerNorm before the GEMM instead of behind it): same step within the bf16 tolerance
    store = torch.zeros(x.shape[0], T, 160, dtype=dtype, device=cuda_device)
    store[:, :, :Cf] = x
    res = []
    for xin in (store[:, :, :Cf], xw, x):
        out, st = engine.elbo_forward(L, flat, xin, y, date_ptr, train=True, precision=prec, philox=(1, 0, 0))
        res.append((float(out["loss"]), out["yhat"].clone(), engine.elbo_backward(L, st).clone()))
    assert np.isfinite(res[0][0])
    if prec == "fp32":
        assert res[0][0] == res[1][0] == res[2][0]
        assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][1], res[2][1])
    else:   # the tensor-core heads accumulate softmax sums with float atomics: order, hence the last bits, are free
        assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[0][0])
        assert float((res[0][1] - res[1][1]).abs().max()) <= 1e-4
        assert abs(res[0][0] - res[2][0]) <= 2e-3 * abs(res[0][0])
    # gradients: identical inputs to every kernel; float atomics make the accumulation order free
    assert float((res[0][2] - res[1][2]).abs().max()) <= 1e-5 * float(res[0][2].abs().max())
    if prec != "fp32" or dtype == torch.float32:
        assert float((res[0][2] - res[2][2]).abs().max()) <= (1e-5 if prec == "fp32" else 3e-2) * float(res[0][2].abs().max())


def test_batched_scoring_loop_is_chunking_invariant(cuda_device):
    """generate_prediction_scores (utils.py:68-93 replacement): same scores whether 1, 3 or all dates go through one
    fvae_predict call; frame indexed like the reference's; deterministic option returns mu_y."""
    from factorvae_b200 import engine
    from factorvae_b200.inference import generate_prediction_scores
    from factorvae_b200.panel import ResidentPanel
    import factorvae_b200 as fb
    rng = np.random.default_rng(9)
    D, I, Cf, T, H, K = 11, 140, 158, 5, 20, 20
    dates = pd.date_range("2022-01-03", periods=D, freq="D")
    rows = [(d, f"S{j:04d}") for d in dates for j in range(I) if rng.random() > 0.1]
    df = pd.DataFrame(np.clip(rng.standard_normal((len(rows), Cf + 1)), -3, 3).astype(np.float32),
                      index=pd.MultiIndex.from_tuples(rows, names=["datetime", "instrument"]))
    rp = ResidentPanel.from_dataframe(df, Cf, cuda_device, start=dates[4])
    torch.manual_seed(6)
    m = fb.FactorVAE(fb.FeatureExtractor(Cf, H), fb.FactorEncoder(K, 128, H),
                     fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)), fb.FactorPredictor(H, K))
    L = engine.ParamLayout(Cf, H, K, 128)
    flat = L.pack(m.state_dict(), cuda_device)
    frames = [generate_prediction_scores(L, flat, rp, T, dates_per_call=n, seed=3)[0] for n in (1, 3, 64)]
    assert list(frames[0].index.names) == ["datetime", "instrument"] and list(frames[0].columns) == ["score"]
    assert frames[0].index.equals(df.sort_index().loc[dates[4]:].index)
    for f in frames[1:]:
        assert np.abs(f["score"].to_numpy() - frames[0]["score"].to_numpy()).max() <= 1e-4
    det, extras = generate_prediction_scores(L, flat, rp, T, deterministic=True)
    assert np.array_equal(det["score"].to_numpy(), extras["mu_y"])
    assert np.isfinite(det["score"].to_numpy()).all() and (extras["sigma_y"] > 0).all()
