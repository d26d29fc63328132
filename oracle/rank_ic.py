"""CPU restatement of the reference's RankIC (utils.py:113-129): per date pandas `.rank()` (method='average') of both
columns, scipy.stats.spearmanr of the two rank vectors; then np.mean / np.std over the dates.
Test infrastructure only (tests/, smoke, bench baseline): the product path never imports this module."""
import numpy as np
import pandas as pd
from scipy.stats import spearmanr


def rank_ic(pred: np.ndarray, label: np.ndarray, date_ptr: np.ndarray):
    vals = []
    for d in range(len(date_ptr) - 1):
        lo, hi = int(date_ptr[d]), int(date_ptr[d + 1])
        daily = pd.DataFrame({"LABEL0": label[lo:hi], "Pred": pred[lo:hi]})
        if hi - lo < 2:
            vals.append(np.nan)
            continue
        ric, _ = spearmanr(daily["LABEL0"].rank(), daily["Pred"].rank())      # utils.py:118-120
        vals.append(ric)
    vals = np.asarray(vals, dtype=np.float64)
    mean = np.mean(vals)
    std = np.std(vals)
    return vals, mean, (mean / std if std != 0 else np.nan)                     # utils.py:126-128
