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


if __name__ == "__main__":
    main()
