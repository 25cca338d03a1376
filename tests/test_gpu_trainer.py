"""Trainer-level drop-in proof: the UNMODIFIED reference trainer `weargait_train.run_cv` (train/weargait_train.py:533-642)
runs on gaitk through module shadowing (`gaitk.integration.install_shadow`) -- device-resident
data path, fused kernels behind the autograd bridge, CUDA losses, on-device CAGrad -- and reproduces what the reference
itself printed on CPU (tests/golden/run_cv_*.json, made by oracle/make_trainer_golden.py): per-epoch losses, accuracies and
the 7-mask table of the restored best model, within north_star's 0.5 accuracy points.

The reference copy travels in oracle/_ref (oracle/build_ref.py); the test skips when it is absent."""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))


@pytest.fixture(scope="module")
def harness():
    import ref_harness as H
    if H.load_reference() is None:
        pytest.skip("no reference copy (oracle/_ref): run python oracle/build_ref.py where /root/reference is mounted")
    return H


@pytest.mark.parametrize("name,dtype", [("run_cv_sync", "f32"), ("run_cv_async", "f32"), ("run_cv_sync", "bf16x3")])
def test_unmodified_run_cv_on_the_dropin(harness, name, dtype):
    import gaitk
    from make_trainer_golden import run_reference_cv
    gold = json.loads((ROOT / "tests" / "golden" / f"{name}.json").read_text())
    meta = gold["meta"]
    WT = gaitk.integration.install_shadow(harness.reference_root())
    assert WT.WearGaitThreeModal is gaitk.WearGaitThreeModal and WT.CAGrad is gaitk.CAGrad and WT.GCLLoss is gaitk.GCLLoss
    assert WT.prepare_split is gaitk.dataloader_weargait.prepare_split
    WT.DEVICE = torch.device("cuda")
    old = gaitk.WearGaitThreeModal.compute_dtype
    gaitk.WearGaitThreeModal.compute_dtype = {"f32": gaitk.DTYPE_F32, "bf16x3": gaitk.DTYPE_BF16X3}[dtype]
    try:
        with tempfile.TemporaryDirectory() as td:
            harness.write_synthetic_weargait(Path(td), n_per_class=meta["n_per_class"], seed=meta["data_seed"], frames=tuple(meta["frames"]))
            res = run_reference_cv(WT, Path(td), async_loading=meta["async_loading"], wm=meta["wm"])
    finally:
        gaitk.WearGaitThreeModal.compute_dtype = old
    assert len(res["epochs"]) == len(gold["epochs"]) > 0
    tol_l = 2e-3 if dtype == "f32" else 5e-3                           # printed with three decimals
    for e, (a, b) in enumerate(zip(res["epochs"], gold["epochs"])):
        assert np.allclose(a["train_loss"], b["train_loss"], atol=tol_l), (e, a, b)
        assert np.allclose(a["val_loss"], b["val_loss"], atol=tol_l), (e, a, b)
        assert np.abs(np.array(a["train_acc"]) - np.array(b["train_acc"])).max() <= 0.5, (e, a, b)
        assert np.abs(np.array(a["val_acc"]) - np.array(b["val_acc"])).max() <= 0.5, (e, a, b)
    assert np.abs(np.array(res["best"]) - np.array(gold["best"])).max() <= 0.5, (res["best"], gold["best"])
    assert set(res["masks"]) == set(gold["masks"]) and len(gold["masks"]) == 7
    for k, v in gold["masks"].items():
        if isinstance(v, dict):
            assert set(res["masks"][k]) == set(v)
            for kk in v:
                assert abs(res["masks"][k][kk] - v[kk]) <= 0.5, (k, kk, res["masks"][k], v)
        else:
            assert abs(res["masks"][k] - v) <= 0.5, (k, res["masks"][k], v)


@pytest.mark.parametrize("name", ["fog_main_sync", "fog_main_async"])
def test_unmodified_fbg_fog_main_on_the_dropin(harness, name):
    """The UNMODIFIED FoG trainer `fbg_fog_train.main` (train/fbg_fog_train.py:410-438: folds, create_fusion_loaders, choose_model,
    class counts, GCL branch losses, CAGrad(n_tasks=2), process_batch incl. the symmetric-KL consistency term in the synchronised
    case, run_epoch, reports) on gaitk through module shadowing -- device-resident clip stores and loaders (dataloader_fbg_fog),
    fused kernels behind the autograd bridge, CUDA losses, on-device CAGrad -- against what the reference printed on CPU
    (tests/golden/fog_main_*.json, oracle/make_trainer_golden.py)."""
    import importlib
    import gaitk
    from make_trainer_golden import run_reference_fog_main
    gold = json.loads((ROOT / "tests" / "golden" / f"{name}.json").read_text())
    meta = gold["meta"]
    gaitk.integration.install_shadow(harness.reference_root())
    FT = importlib.import_module("fbg_fog_train")
    assert FT.create_fusion_loaders is gaitk.dataloader_fbg_fog.create_fusion_loaders and FT.CAGrad is gaitk.CAGrad
    import utilities
    assert utilities.MultiModalMultiTaskModel is gaitk.MultiModalMultiTaskModel
    FT.DEVICE = torch.device("cuda")
    reader, _ = harness.synthetic_fog_reader(**meta["reader"])
    res = run_reference_fog_main(FT, reader, synchronized_loading=meta["synchronized_loading"])
    assert len(res["epochs"]) == len(gold["epochs"]) > 0
    for a, b in zip(res["epochs"], gold["epochs"]):
        assert (a["fold"], a["ep"]) == (b["fold"], b["ep"])
        assert abs(a["train_loss"] - b["train_loss"]) <= 2e-3 and abs(a["val_loss"] - b["val_loss"]) <= 2e-3, (a, b)
        assert np.abs(np.array(a["train_acc"]) - np.array(b["train_acc"])).max() <= 0.5, (a, b)
        assert np.abs(np.array(a["val_acc"]) - np.array(b["val_acc"])).max() <= 0.5, (a, b)
    assert len(res["best"]) == len(gold["best"]) > 0
    assert np.abs(np.array(res["best"]) - np.array(gold["best"])).max() <= 0.5, (res["best"], gold["best"])
    assert np.abs(np.array(res["mean"]) - np.array(gold["mean"])).max() <= 0.5, (res["mean"], gold["mean"])
