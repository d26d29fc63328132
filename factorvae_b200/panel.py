"""Resident panel: the reference rebuilds every look-back window on the host, per sample,
with pandas / numpy (`TSDataSampler`, dataset.py:41-181; `DateGroupedBatchSampler`, dataset.py:207-249).  Here the
(date, instrument) row table is uploaded ONCE, the per-batch work is one tiny index kernel (`fvae_window_index`), and the
FeatureExtractor kernels read the rows through that index (`engine.IndexedWindows`) -- no window tensor, no H2D per step.

Host-side index construction mirrors the reference line by line (vectorised instead of `iterrows`):
    data.sort_index(); to_numpy + one all-NaN sentinel row          dataset.py:77-84
    idx_df = Series(range(R), index).unstack(), sorted both ways     dataset.py:128-131
    idx_map: table row -> (date row i, instrument column j)          dataset.py:133-137
    start_idx / end_idx = data_index.slice_locs(start, end)           dataset.py:97-99
    one batch per date, in index order                                dataset.py:219-233
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .engine import IndexedWindows, _on, _stream

FILL = {"none": _cabi.FILL_NONE, "ffill": _cabi.FILL_FFILL, "ffill+bfill": _cabi.FILL_FFILL_BFILL}


class PanelIndex:
    """Pure-numpy result of the reference's index construction (no torch, no GPU): testable on CPU."""

    def __init__(self, idx_mat: np.ndarray, sample_date: np.ndarray, sample_inst: np.ndarray, date_ptr: np.ndarray,
                 num_rows: int, dates=None, instruments=None):
        self.idx_mat = np.ascontiguousarray(idx_mat, dtype=np.int32)          # (D, I), -1 = no row
        self.sample_date = np.ascontiguousarray(sample_date, dtype=np.int32)   # (N,) date row of every sample in [start, end)
        self.sample_inst = np.ascontiguousarray(sample_inst, dtype=np.int32)   # (N,) instrument column
        self.date_ptr = np.ascontiguousarray(date_ptr, dtype=np.int64)         # (n_dates + 1,) CSR over the batches (dates)
        self.num_rows = int(num_rows)                                          # table rows without the sentinel
        self.dates, self.instruments = dates, instruments

    @property
    def nan_row(self) -> int:
        return self.num_rows          # the all-NaN sentinel row appended to the table (dataset.py:81-84)

    @staticmethod
    def from_dataframe(df, start=None, end=None) -> "PanelIndex":
        import pandas as pd
        if list(df.index.names) != ["datetime", "instrument"]:
            raise ValueError("the frame must be indexed by (datetime, instrument) like the reference's pickle")
        data = df.sort_index()
        R = data.shape[0]
        dt = data.index.get_level_values("datetime")
        inst = data.index.get_level_values("instrument")
        dates = dt.unique().sort_values()
        instruments = inst.unique().sort_values()
        i_of = dates.get_indexer(dt).astype(np.int64)           # date row of every table row
        j_of = instruments.get_indexer(inst).astype(np.int64)   # instrument column
        idx_mat = np.full((len(dates), len(instruments)), -1, dtype=np.int32)
        idx_mat[i_of, j_of] = np.arange(R, dtype=np.int32)
        lo, hi = data.index.slice_locs(start=None if start is None else pd.Timestamp(start),
                                       end=None if end is None else pd.Timestamp(end))
        sd, sj = i_of[lo:hi], j_of[lo:hi]
        # one batch per date, dates in index order (DateGroupedBatchSampler with shuffle=False)
        change = np.flatnonzero(np.diff(sd)) + 1
        date_ptr = np.concatenate([[0], change, [hi - lo]]).astype(np.int64) if hi > lo else np.zeros(1, np.int64)
        return PanelIndex(idx_mat, sd, sj, date_ptr, R, dates, instruments)

    def window_rows(self, samples: np.ndarray, T: int, fill: str = "ffill+bfill") -> np.ndarray:
        """CPU restatement of TSDataSampler._get_indices + nan_to_num for the given sample numbers -> (n, T) table rows.
        Test oracle for the device kernel (the product path never calls it)."""
        mode = FILL[fill]
        out = np.empty((len(samples), T), dtype=np.int64)
        D = self.idx_mat.shape[0]
        for n, s in enumerate(samples):
            i, j = int(self.sample_date[s]), int(self.sample_inst[s])
            v = np.full(T, -1, dtype=np.int64)
            lo = max(i - T + 1, 0)
            v[T - (i + 1 - lo):] = self.idx_mat[lo:i + 1, j]
            if mode != _cabi.FILL_NONE:
                last = -1
                for t in range(T):
                    if v[t] < 0:
                        v[t] = last
                    else:
                        last = v[t]
                if mode == _cabi.FILL_FFILL_BFILL:
                    nxt = -1
                    for t in range(T - 1, -1, -1):
                        if v[t] < 0:
                            v[t] = nxt
                        else:
                            nxt = v[t]
            v[v < 0] = self.nan_row
            out[n] = v
        return out


class ResidentPanel:
    """The row table and its index on one device.  `batch(...)` is the per-step work: one kernel launch."""

    def __init__(self, values: np.ndarray, index: PanelIndex, num_features: int, device, dtype=torch.bfloat16,
                 label_col: int = -1):
        if values.ndim != 2 or values.shape[0] != index.num_rows:
            raise ValueError("values must be the (rows, features + label) matrix of the sorted frame")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("ResidentPanel lives on a CUDA device: factorvae_b200 has no CPU path")
        self.index, self.C, self.device = index, int(num_features), dev
        vals = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float32))
        feat = torch.cat([vals[:, :num_features], torch.full((1, num_features), float("nan"))], dim=0)
        lab = torch.cat([vals[:, label_col], torch.full((1,), float("nan"))], dim=0)
        # row pitch padded to 16 bytes: every row then starts 16-byte aligned, so the staging kernel sees ONE alignment
        # case for the whole warp (with the natural 316-byte pitch consecutive rows cycle through four, and the realign
        # switch of K1 diverges four ways)
        per16 = 16 // torch.empty((), dtype=dtype).element_size()
        pitch = (num_features + per16 - 1) // per16 * per16
        table = torch.zeros(feat.shape[0], pitch, dtype=dtype, device=dev)
        table[:, :num_features] = feat.to(device=dev, dtype=dtype)
        self.table = table                                                       # (R + 1, pitch >= C), last row = NaN sentinel
        self.label = lab.to(device=dev, dtype=torch.float32).contiguous()       # (R + 1,)
        self.idx_mat = torch.from_numpy(index.idx_mat).to(dev)
        self.sample_date = torch.from_numpy(index.sample_date).to(dev)
        self.sample_inst = torch.from_numpy(index.sample_inst).to(dev)
        self._batch_cache = {}

    @staticmethod
    def from_dataframe(df, num_features: int, device, start=None, end=None, dtype=torch.bfloat16) -> "ResidentPanel":
        idx = PanelIndex.from_dataframe(df, start, end)
        return ResidentPanel(df.sort_index().to_numpy(dtype=np.float32), idx, num_features, device, dtype)

    def upload_rows(self, first_row: int, rows_host: torch.Tensor, labels_host: Optional[torch.Tensor] = None) -> None:
        """Overwrite table rows [first_row, first_row + n) from a pinned HOST buffer (n, pitch) of the table's dtype (and
        their labels (n,) fp32): the streaming form of the resident panel -- every (date, instrument) row crosses PCIe
        exactly once, when its date enters the look-back horizon, instead of T times inside T overlapping windows
        (dataset.py:169-181 materialises every window on the host).  Asynchronous on the current stream."""
        n = rows_host.shape[0]
        if rows_host.dtype != self.table.dtype or rows_host.shape[1] != self.table.shape[1] or first_row < 0 or first_row + n > self.index.num_rows:
            raise ValueError("rows_host must be (n, pitch) in the table's dtype and fit the table")
        self.table[first_row:first_row + n].copy_(rows_host, non_blocking=True)
        if labels_host is not None:
            self.label[first_row:first_row + n].copy_(labels_host, non_blocking=True)

    @property
    def num_batches(self) -> int:
        return len(self.index.date_ptr) - 1

    def batch(self, dates: Sequence[int], T: int, fill: str = "ffill+bfill") -> Tuple[IndexedWindows, torch.Tensor, torch.Tensor]:
        """Windows (as an index), labels y (S,) and the CSR date_ptr (B+1,) of the given batches (positions in date order,
        any order / subset: a shuffled epoch is a permutation of range(num_batches))."""
        # the host side of a batch (sample slices, counts, the CSR date_ptr on the device) depends only on WHICH dates are asked
        # for: cached, so a training loop that revisits a batch pays one kernel launch, no host numpy work and no pageable H2D
        key = (dates.start, dates.stop, dates.step) if isinstance(dates, range) else tuple(int(d) for d in dates)
        hit = self._batch_cache.get(key)
        if hit is None:
            dp = self.index.date_ptr
            dl = [int(d) for d in dates]
            counts = np.array([dp[d + 1] - dp[d] for d in dl], dtype=np.int64)
            contiguous = all(dl[k + 1] == dl[k] + 1 for k in range(len(dl) - 1))
            if contiguous:
                sd = self.sample_date[dp[dl[0]]:dp[dl[-1] + 1]]
                sj = self.sample_inst[dp[dl[0]]:dp[dl[-1] + 1]]
            else:
                sel = torch.from_numpy(np.concatenate([np.arange(dp[d], dp[d + 1]) for d in dl])).to(self.device)
                sd, sj = self.sample_date[sel].contiguous(), self.sample_inst[sel].contiguous()
            date_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).to(self.device)
            if len(self._batch_cache) > 4096:
                self._batch_cache.clear()
            hit = self._batch_cache[key] = (sd, sj, int(counts.sum()), date_ptr)
        sd, sj, S, date_ptr = hit
        row_index = torch.empty(S, T, dtype=torch.int32, device=self.device)
        y = torch.empty(S, dtype=torch.float32, device=self.device)
        D, I = self.index.idx_mat.shape
        with _on(self.device):
            rc = _cabi.lib().fvae_window_index(self.idx_mat.data_ptr(), D, I, sd.data_ptr(), sj.data_ptr(), S, T, FILL[fill],
                                               self.index.nan_row, row_index.data_ptr(), self.label.data_ptr(), y.data_ptr(),
                                               _stream(self.device))
        _cabi.check(rc, "fvae_window_index")
        return IndexedWindows(self.table, row_index, self.C), y, date_ptr

    def windows(self, xw: IndexedWindows, dtype=torch.float32) -> torch.Tensor:
        """Materialise the (S, T, C) window tensor the reference's DataLoader would have produced (train_model.py:17-19)."""
        S, T, Cf = xw.shape
        out = torch.empty(S, T, Cf, dtype=dtype, device=self.device)
        from .engine import _panel
        _, panel = _panel(xw)
        with _on(self.device):
            rc = _cabi.lib().fvae_gather_windows(C.byref(panel), S, T, Cf, out.data_ptr(),
                                                 _cabi.F32 if dtype == torch.float32 else _cabi.BF16, _stream(self.device))
        _cabi.check(rc, "fvae_gather_windows")
        return out
