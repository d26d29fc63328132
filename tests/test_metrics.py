"""Device RankIC against the CPU restatement of utils.py:113-129 (pandas rank + scipy spearmanr)."""
import numpy as np
import pytest
import torch


def test_oracle_rank_ic_is_spearman_on_known_cases():
    from oracle.rank_ic import rank_ic
    ptr = np.array([0, 5, 10])
    pred = np.array([1, 2, 3, 4, 5, 5, 4, 3, 2, 1], dtype=np.float64)
    lab = np.array([10, 20, 30, 40, 50, 1, 2, 3, 4, 5], dtype=np.float64)
    v, m, ir = rank_ic(pred, lab, ptr)
    assert np.allclose(v, [1.0, -1.0]) and abs(m) < 1e-12


@pytest.mark.gpu
def test_device_rank_ic_matches_pandas_scipy(cuda_device):
    from factorvae_b200.metrics import rank_ic
    from oracle.rank_ic import rank_ic as ref_rank_ic
    rng = np.random.default_rng(4)
    counts = [300, 2, 17, 1000, 64, 3000, 5]
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    S = int(ptr[-1])
    lab = rng.standard_normal(S).astype(np.float32)
    pred = (0.3 * lab + rng.standard_normal(S)).astype(np.float32)
    pred[:300] = np.round(pred[:300], 1)            # ties -> average ranks
    lab[302:319] = np.round(lab[302:319])           # heavy ties in a small date
    pred[ptr[4]:ptr[5]] = 1.5                        # constant column -> NaN (scipy: correlation undefined)
    pred[ptr[6] + 2] = np.nan                        # NaN -> NaN for that date
    ric, mean, ir = rank_ic(torch.from_numpy(pred).to(cuda_device), torch.from_numpy(lab).to(cuda_device),
                            torch.from_numpy(ptr).to(cuda_device))
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref, rmean, rir = ref_rank_ic(pred.astype(np.float64), lab.astype(np.float64), ptr)
    got = ric.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), (got, ref)
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= 2e-6
    assert ok.sum() == 5
