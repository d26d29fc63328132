"""Resident panel, host side: the vectorised index construction and the CPU restatement of
TSDataSampler._get_indices against golden vectors produced by the reference's own dataset.py
(oracle/gen_panel_golden.py).  No GPU."""
import os

import numpy as np
import pandas as pd
import pytest

from factorvae_b200.panel import PanelIndex

GOLD = os.path.join(os.path.dirname(__file__), "golden", "panel_windows.npz")


def _frame(g, shuffled=True):
    idx = pd.MultiIndex.from_arrays([pd.to_datetime(g["datetime_ns"]), g["instrument"].astype(str)], names=["datetime", "instrument"])
    cols = [f"F{k}" for k in range(g["values"].shape[1] - 1)] + ["LABEL0"]
    df = pd.DataFrame(g["values"], index=idx, columns=cols)
    return df.sample(frac=1.0, random_state=11) if shuffled else df


def test_index_construction_matches_reference_sampler():
    g = np.load(GOLD)
    pi = PanelIndex.from_dataframe(_frame(g), start=pd.Timestamp(int(g["start_ns"])), end=pd.Timestamp(int(g["end_ns"])))
    assert np.array_equal(pi.idx_mat.astype(np.int64), g["ref_idx_arr"])                     # build_index, dataset.py:128-137
    assert len(pi.sample_date) == int(g["ref_end_idx"]) - int(g["ref_start_idx"])            # slice_locs, dataset.py:97-99
    assert np.array_equal(np.diff(pi.date_ptr), g["counts_ffill_bfill"])                     # one batch per date, :219-233
    # every sample names an existing (date, instrument) row, in table order
    rows = pi.idx_mat[pi.sample_date, pi.sample_inst]
    assert np.array_equal(rows, np.arange(int(g["ref_start_idx"]), int(g["ref_end_idx"])))


@pytest.mark.parametrize("fill", ["none", "ffill", "ffill+bfill"])
def test_window_rows_restatement_matches_reference_windows(fill):
    g = np.load(GOLD)
    T = int(g["T"])
    pi = PanelIndex.from_dataframe(_frame(g), start=pd.Timestamp(int(g["start_ns"])), end=pd.Timestamp(int(g["end_ns"])))
    rows = pi.window_rows(np.arange(len(pi.sample_date)), T, fill)
    table = np.concatenate([g["values"], np.full((1, g["values"].shape[1]), np.nan)], axis=0)   # sentinel row, dataset.py:81-84
    got = table[rows]
    ref = g["windows_" + fill.replace("+", "_")]
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))
    if fill == "ffill+bfill":      # a present sample never keeps a missing step after ffill + bfill
        assert (rows < pi.nan_row).all()
