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
