"""Runs the UNMODIFIED reference drivers -- main.main (main.py:19-87: model build :27-33, loaders :36-51, Adam + cosine
:60-61, epoch loop with train / validate :66-80, best-val checkpoint :73-80), train_model.train / validate
(train_model.py:11-60) and utils.load_model / generate_prediction_scores (utils.py:57-93) -- against a `module` package found in
<module_dir>.  With <module_dir> = dropin/ that is the drop-in proof (tests/test_dropin_reference_drivers_gpu.py); with
<module_dir> = the reference itself it validates this harness on CPU.

    python tests/_dropin_driver.py <module_dir> <reference_dir> <workdir>

The harness only supplies what the environment lacks, never model code: a synthetic (datetime, instrument) pickle shaped like
data/make_dataset.py's export, a stub `matplotlib` (imported but unused by dataset.py:8 / main.py:8), and a
pandas-3 shim that makes DateGroupedBatchSampler's group array writable for np.random.shuffle (dataset.py:224,230).
Prints one JSON line: per-epoch train / validation losses, checkpoint path, reload check, prediction-score frame shape."""
import argparse
import contextlib
import io
import json
import os
import re
import sys


def make_pickle(path, n_dates=44, n_inst=36, seed=0):
    import numpy as np
    import pandas as pd
    rng = np.random.default_rng(seed)
    dates = pd.bdate_range("2020-01-01", periods=n_dates)
    inst = [f"SH{600000 + i}" for i in range(n_inst)]
    rows = [(d, s) for d in dates for s in inst if rng.random() > 0.08]            # ragged membership
    idx = pd.MultiIndex.from_tuples(rows, names=["datetime", "instrument"])
    R = len(rows)
    feat = np.clip(rng.standard_normal((R, 158)), -3, 3).astype(np.float32)
    w = rng.standard_normal(158).astype(np.float32) / 12.0
    label = (feat @ w + 0.3 * rng.standard_normal(R)).astype(np.float32)           # learnable signal
    cols = [f"F{i}" for i in range(158)] + ["LABEL_RAW", "MKT0", "MKT1"]             # main.py:36 keeps the first 159 columns
    df = pd.DataFrame(np.concatenate([feat, label[:, None], rng.standard_normal((R, 2)).astype(np.float32)], axis=1),
                      index=idx, columns=cols).sort_index()
    df.to_pickle(path)
    return dates


def main():
    module_dir, ref_dir, work = (os.path.abspath(p) for p in sys.argv[1:4])
    os.makedirs(work, exist_ok=True)
    stubs = os.path.join(work, "stubs", "matplotlib")
    os.makedirs(stubs, exist_ok=True)
    open(os.path.join(stubs, "__init__.py"), "w").write("")
    open(os.path.join(stubs, "pyplot.py"), "w").write("")
    pkl = os.path.join(work, "synthetic_csi.pkl")
    dates = make_pickle(pkl)
    sys.path[:0] = [os.path.dirname(stubs), module_dir, ref_dir]
    os.environ.setdefault("WANDB_MODE", "disabled")

    import numpy as np
    import torch
    import dataset as ref_dataset                    # the reference's dataset.py, unmodified
    import main as ref_main                          # the reference's main.py, unmodified
    import module as used_module
    import train_model as ref_train
    import utils as ref_utils

    orig_group = ref_dataset.DateGroupedBatchSampler._group_indices_by_date

    def writable_groups(self):                       # pandas 3: `.values` is read-only, np.random.shuffle needs to write
        return np.array(orig_group(self), dtype=object, copy=True)
    ref_dataset.DateGroupedBatchSampler._group_indices_by_date = writable_groups

    fmt = lambda d: d.strftime("%Y-%m-%d")
    args = argparse.Namespace(num_epochs=3, lr=2e-3, num_latent=158, num_portfolio=16, seq_len=5, num_factor=8, hidden_size=16,
                              dataset=pkl, start_time=fmt(dates[0]), fit_end_time=fmt(dates[31]), val_start_time=fmt(dates[32]),
                              val_end_time=fmt(dates[-1]), end_time=fmt(dates[-1]), seed=42, run_name="dropin",
                              save_dir=os.path.join(work, "best_models"), num_workers=0, wandb=False)
    data_args = ref_utils.DataArgument(start_time=args.start_time, end_time=args.end_time, fit_end_time=args.fit_end_time,
                                       val_start_time=args.val_start_time, val_end_time=args.val_end_time, seq_len=args.seq_len)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref_main.main(args, data_args)               # main.py:19-87
    log = buf.getvalue()
    epochs = [(float(a), float(b)) for a, b in re.findall(r"Train Loss: ([-\d.einf]+), Validation Loss: ([-\d.einf]+)", log)]
    saved = re.findall(r"Model saved at (\S+)", log)
    ckpt = saved[-1] if saved else None

    # utils.load_model (utils.py:57-67) + the checkpoint main.py wrote (main.py:79) + validate (train_model.py:40-60)
    model = ref_utils.load_model(args)
    state = torch.load(ckpt, map_location="cpu")
    missing = model.load_state_dict(state)
    import pandas as pd
    df = pd.read_pickle(pkl).iloc[:, :159]
    df.rename(columns={df.columns[-1]: "LABEL0"}, inplace=True)
    loader = ref_dataset.init_data_loader(df, shuffle=False, step_len=args.seq_len, start=args.val_start_time, end=args.val_end_time)
    with contextlib.redirect_stdout(io.StringIO()):
        v1 = ref_train.validate(model, loader, args)
    # utils.generate_prediction_scores (utils.py:70-93) needs args.seq_length and a dataset with get_index()
    args.seq_length = args.seq_len
    score_shape = None
    try:
        ds = loader.dataset.sampler                      # TSDataSampler.get_index (dataset.py:124)
        loader_x = [(cw[:, :, :-1], None) for cw, _ in loader]
        with contextlib.redirect_stdout(io.StringIO()):
            frame = ref_utils.generate_prediction_scores(model, loader_x, ds, args)
        score_shape = list(frame.shape)
        scores_finite = bool(np.isfinite(frame["score"].to_numpy()).all())
    except Exception as e:                            # reported, not hidden
        import traceback
        score_shape, scores_finite = f"{type(e).__name__}: {e} | " + traceback.format_exc()[-600:], False
    print(json.dumps({"module_file": used_module.__file__, "device": str(next(model.parameters()).device), "epochs": epochs,
                      "checkpoint": ckpt, "state_keys": len(state), "load_state_dict": str(missing), "reloaded_val_loss": v1,
                      "score_frame_shape": score_shape, "scores_finite": scores_finite}))


if __name__ == "__main__":
    main()
