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


@pytest.mark.gpu
def test_device_rank_ic_nan_and_inf_in_the_largest_non_power_of_two_date(cuda_device):
    """A NaN or a genuine +inf in a date whose size equals max_per_date and is not a power of two (300 -> 512 sort slots):
    the padding slots share the +inf key, so the sort must be total (key, index) or they land among the real entries and
    `rank[idx[p]]` writes past the rank array.  NaN -> NaN for the date; +inf ranks last, like pandas."""
    from factorvae_b200.metrics import rank_ic
    from oracle.rank_ic import rank_ic as ref_rank_ic
    rng = np.random.default_rng(9)
    counts = [300, 300, 300, 77]
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    S = int(ptr[-1])
    lab = rng.standard_normal(S).astype(np.float32)
    pred = (0.5 * lab + rng.standard_normal(S)).astype(np.float32)
    pred[5] = np.nan                                  # date 0: NaN
    pred[300 + 17] = np.inf                           # date 1: one +inf prediction
    pred[300 + 250] = np.inf                          #         and a second one (tie at +inf)
    lab[600 + 3] = np.inf                             # date 2: +inf label
    for _ in range(3):                                # repeat: an out-of-range shared write shows up as a sticky fault
        ric, _, _ = rank_ic(torch.from_numpy(pred).to(cuda_device), torch.from_numpy(lab).to(cuda_device),
                            torch.from_numpy(ptr).to(cuda_device))
        torch.cuda.synchronize()
    import warnings
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, _, _ = ref_rank_ic(pred.astype(np.float64), lab.astype(np.float64), ptr)
    got = ric.cpu().numpy().astype(np.float64)
    assert np.isnan(got[0]) and np.isnan(ref[0])
    assert np.array_equal(np.isnan(got), np.isnan(ref)), (got, ref)
    ok = ~np.isnan(ref)
    assert ok.sum() == 3 and np.abs(got[ok] - ref[ok]).max() <= 2e-6, (got, ref)
