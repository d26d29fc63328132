"""Batched scoring loop: the reference walks the test loader one date at a time,
`predictions = model.prediction(char.float())`, concatenates on the CPU and wraps the result in a
(datetime, instrument)-indexed frame with one column 'score' (utils.py:68-93).  Here many dates go through ONE
`fvae_predict` call (any number of dates per call, rows read in place from the resident panel), the sample index comes
from the panel, and a deterministic option returns mu_y instead of the reparameterised sample."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import engine
from .panel import ResidentPanel


@torch.no_grad()
def generate_prediction_scores(layout: engine.ParamLayout, flat: torch.Tensor, panel: ResidentPanel, T: int, *,
                               dates_per_call: int = 64, precision: str = "bf16", seed: int = 0,
                               deterministic: bool = False, fill: str = "ffill+bfill"):
    """Scores for every sample of the panel's [start, end) range, in index order.

    Returns (frame, extras): frame is the reference's DataFrame (index (datetime, instrument), column 'score'; a plain
    dict of numpy arrays if the panel was not built from a pandas frame); extras holds mu_y, sigma_y and the per-date
    prior mu / sigma.  score = yhat = mu_y + eps * sigma_y as in FactorVAE.prediction (module.py:273-278); with
    deterministic=True score = mu_y.  The noise is keyed by the GLOBAL sample number, so the result does not depend on
    dates_per_call."""
    nb = panel.num_batches
    dp = panel.index.date_ptr
    scores, mus, sigmas, pmu, psg = [], [], [], [], []
    workspace: Optional[torch.Tensor] = None
    for b0 in range(0, nb, dates_per_call):
        b1 = min(nb, b0 + dates_per_call)
        xw, _, date_ptr = panel.batch(range(b0, b1), T, fill)
        out, st = engine.elbo_forward(layout, flat, xw, None, date_ptr, train=False, precision=precision, predict=True,
                                      philox=(seed, 0, int(dp[b0])), workspace=workspace)
        workspace = st.workspace
        scores.append((out["mu_y"] if deterministic else out["yhat"]).clone())
        mus.append(out["mu_y"].clone()); sigmas.append(out["sigma_y"].clone())
        pmu.append(out["mu_prior"].clone()); psg.append(out["sigma_prior"].clone())
    score = torch.cat(scores).cpu().numpy()
    extras = dict(mu_y=torch.cat(mus).cpu().numpy(), sigma_y=torch.cat(sigmas).cpu().numpy(),
                  mu_prior=torch.cat(pmu).cpu().numpy(), sigma_prior=torch.cat(psg).cpu().numpy())
    idx = panel.index
    if idx.dates is not None and idx.instruments is not None:
        import pandas as pd
        mi = pd.MultiIndex.from_arrays([idx.dates[idx.sample_date], idx.instruments[idx.sample_inst]], names=["datetime", "instrument"])
        return pd.DataFrame(score.reshape(-1, 1), index=mi, columns=["score"]), extras
    return dict(score=score, sample_date=idx.sample_date, sample_inst=idx.sample_inst), extras
