"""Trainer-level goldens: run the UNMODIFIED reference trainer `weargait_train.run_cv` (train/weargait_train.py:533-642) on
CPU on a synthetic WearGait directory and record what it prints -- best fold accuracies and the 7-mask table.  The GPU test
(tests/test_gpu_trainer.py) runs the very same unmodified `run_cv` with gaitk shadowing the encoder / loss / CAGrad / data
modules and must land within 0.5 accuracy points (north_star's end-to-end bar).

    python oracle/make_trainer_golden.py        # rewrites tests/golden/run_cv_*.json (needs oracle/_ref or /root/reference)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
OUT = HERE.parent / "tests" / "golden"

CASES = {
    "run_cv_sync": dict(async_loading=False, wm="gcl"),
    "run_cv_async": dict(async_loading=True, wm="gcl"),
}


def trainer_args(**kw):
    """argparse defaults of weargait_train.py:646-689, shrunk to a minutes-scale run"""
    a = dict(n_folds=1, test_per_class=2, win_len=64, hop_len=64, batch_size=64, num_workers=0, epochs=3, patience=50,
             num_classes=2, lr=1e-3, seed=43, async_loading=False, single_mod=None, proj_ch=16, enc_out_ch=12, backbone_dim=8,
             shared_out_ch=16, use_norm=False, use_cosine=False, baseline=None, wm="gcl", gcl_m=0.2, gcl_s=25, noise_mul=0.0,
             drw_warmup=0, alpha=0.5)
    a.update(kw)
    return SimpleNamespace(**a)


def parse_run_cv_output(text: str) -> dict:
    """what run_cv prints: per-mask accuracy of the restored best model and the best macro / per-stream accuracies"""
    out = {"masks": {}, "epochs": []}
    for m in re.finditer(r"\[SYNC\]\[mask=([WIM+]+)\] acc=\s*([0-9.]+)%", text):
        out["masks"][m.group(1)] = float(m.group(2))
    for m in re.finditer(r"\[ASYNC\]\[mask=([WIM+]+)\] (\{.*\})", text):
        out["masks"][m.group(1)] = {k: float(v) for k, v in re.findall(r"'(\w+)': (?:np\.float64\()?([0-9.]+)", m.group(2))}
    m = re.search(r"Best macro acc: ([0-9.]+)% \(W=([0-9.]+) I=([0-9.]+) M=([0-9.]+)\)", text)
    if m:
        out["best"] = [float(x) for x in m.groups()]
    for m in re.finditer(r"Ep (\d+) \| L=\[([0-9.,nan]+)\] acc=\[([0-9., ]+)\] \| L=\[([0-9.,nan]+)\] acc=\[([0-9., ]+)\]", text):
        out["epochs"].append({"train_loss": [float(x) for x in m.group(2).split(",")], "train_acc": [float(x) for x in m.group(3).split(",")],
                              "val_loss": [float(x) for x in m.group(4).split(",")], "val_acc": [float(x) for x in m.group(5).split(",")]})
    return out


def run_reference_cv(WT, root: Path, **kw) -> dict:
    buf = io.StringIO()
    cwd = os.getcwd()
    os.chdir(root)
    try:
        with contextlib.redirect_stdout(buf):
            WT.run_cv(trainer_args(**kw))
    finally:
        os.chdir(cwd)
    return parse_run_cv_output(buf.getvalue())


# ---------------------------------------------------------------------------------------------- fbg_fog_train.main
FOG_CASES = {
    "fog_main_sync": dict(synchronized_loading=True),
    "fog_main_async": dict(synchronized_loading=False),
}
FOG_READER = dict(dataset="fog", seed=11, n_subjects=12, scalar_labels=False, informative=True)
FOG_OVERRIDES = dict(epochs=4, batch_size=16)                       # configs.FBG_FOG_PARAMS["fog"] (a dict the trainer reads at run time)


def fog_args(**kw):
    """argparse defaults of fbg_fog_train.py:447-466"""
    a = dict(dataset="fog", modality="multimodal", consistency_lambda=1, seed=43, wm="gcl", synchronized_loading=False, alpha=0.1,
             max_norm=1.0, ldam_s=30, ldam_m=0.5, gcl_m=0.2, gcl_s=25, noise_mul=0, drw_warmup=0, use_norm_and_cos=False,
             save_loss_plots=False, rebuild_cache=False)
    a.update(kw)
    return SimpleNamespace(**a)


def parse_fog_main_output(text: str) -> dict:
    out = {"epochs": [], "best": [], "mean": None}
    for m in re.finditer(r"\[Fold (\d+)\]\[Ep (\d+)/\d+\] Train loss=([0-9.nan]+)\s+acc=([0-9.]+)% \| Eval loss=([0-9.nan]+)\s+ens_acc=([0-9.]+)%", text):
        out["epochs"].append(dict(fold=int(m.group(1)), ep=int(m.group(2)), train_loss=float(m.group(3)), train_acc=[float(m.group(4))],
                                  val_loss=float(m.group(5)), val_acc=[float(m.group(6))]))
    for m in re.finditer(r"\[Fold (\d+)\]\[Ep (\d+)/\d+\] Train loss=([0-9.nan]+) skel=([0-9.]+)% sen=([0-9.]+)% \| Eval loss=([0-9.nan]+) skel=([0-9.]+)% sen=([0-9.]+)% avg=([0-9.]+)%", text):
        out["epochs"].append(dict(fold=int(m.group(1)), ep=int(m.group(2)), train_loss=float(m.group(3)), train_acc=[float(m.group(4)), float(m.group(5))],
                                  val_loss=float(m.group(6)), val_acc=[float(m.group(7)), float(m.group(8)), float(m.group(9))]))
    for m in re.finditer(r"\*\*\* Fold (\d+) Best Ensemble Acc: ([0-9.]+)% \*\*\*", text):
        out["best"].append([float(m.group(2))])
    for m in re.finditer(r"\*\*\* Fold (\d+) Best skel=([0-9.]+)%\s+sens=([0-9.]+)%, avg=([0-9.]+)% \*\*\*", text):
        out["best"].append([float(m.group(2)), float(m.group(3)), float(m.group(4))])
    m = re.search(r"mean Ensemble Acc: ([0-9.]+)%", text)
    if m:
        out["mean"] = [float(m.group(1))]
    m = re.search(r"mean skel=([0-9.]+)%, sensor=([0-9.]+)%, avg=([0-9.]+)%", text)
    if m:
        out["mean"] = [float(x) for x in m.groups()]
    return out


def run_reference_fog_main(FT, reader, max_folds: int = 2, **kw) -> dict:
    """the UNMODIFIED fbg_fog_train.main (:410-438) on an injected reader (load_reader reads a pickle under the reference tree,
    which is read-only here) with the run-time config dict shortened to a minutes-scale run and the fold list cut to max_folds"""
    import configs
    saved = dict(FT.FBG_FOG_PARAMS["fog"]); load_reader = FT.load_reader; gen = FT.generate_class_stratified_folds
    FT.FBG_FOG_PARAMS["fog"].update(FOG_OVERRIDES)
    FT.load_reader = lambda *a, **k: reader
    FT.generate_class_stratified_folds = lambda r, d: gen(r, d)[:max_folds]
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            FT.main(fog_args(**kw))
    finally:
        FT.FBG_FOG_PARAMS["fog"].clear(); FT.FBG_FOG_PARAMS["fog"].update(saved)
        FT.load_reader = load_reader; FT.generate_class_stratified_folds = gen
    return parse_fog_main_output(buf.getvalue())


def main():
    import torch
    import ref_harness as H
    R = H.load_reference()
    assert R is not None, "no reference copy"
    R.WT.DEVICE = torch.device("cpu")
    torch.set_num_threads(8)
    for name, kw in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            H.write_synthetic_weargait(Path(td), n_per_class=6, seed=0, frames=(260, 520))
            with H._float_tensor_on(torch.device("cpu")):
                res = run_reference_cv(R.WT, Path(td), **kw)
        res["meta"] = dict(kw, n_per_class=6, data_seed=0, frames=[260, 520], torch=torch.__version__)
        (OUT / f"{name}.json").write_text(json.dumps(res, indent=1))
        print(name, json.dumps(res)[:400])
    R.FT.DEVICE = torch.device("cpu")
    for name, kw in FOG_CASES.items():
        reader, _ = H.synthetic_fog_reader(**FOG_READER)
        with H._float_tensor_on(torch.device("cpu")):
            res = run_reference_fog_main(R.FT, reader, **kw)
        res["meta"] = dict(kw, reader=FOG_READER, overrides=FOG_OVERRIDES, torch=torch.__version__)
        (OUT / f"{name}.json").write_text(json.dumps(res, indent=1))
        print(name, json.dumps(res)[:600])


if __name__ == "__main__":
    main()
