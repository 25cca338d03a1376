"""Module shadowing: run the UNMODIFIED reference trainers on gaitk.

The trainers import their building blocks by bare module name (``weargait_train.py:23-39``, ``fbg_fog_train.py:19-38``,
``utilities.py:9-10``); ``install_shadow`` registers gaitk's drop-in modules under those names BEFORE the trainer is imported,
and swaps the loss / CAGrad classes inside the reference's own ``learning.optimizers`` modules (the trainers do
``from learning.optimizers.multitask_weighting import CAGrad``).  Nothing under the reference tree is edited.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path


def install_shadow(reference_root, *, data_path: bool = True):
    """reference_root: directory that holds ``train/`` and ``data/WearGait/`` of the reference.  Returns the imported
    (unmodified) ``weargait_train`` module.  data_path=False keeps the reference's own host DataLoaders."""
    from . import classification_losses, dataloader_fbg_fog, dataloader_weargait, feature_encoder, multitask_weighting, weargait_encoders
    root = Path(reference_root)
    for p in (root, root / "data" / "WearGait", root / "train"):
        if str(p) not in sys.path:
            sys.path.insert(0, str(p))
    # the reference's encoder module also defines baselines gaitk does not run on the GPU yet (EarlyFusion3, CheapXAttn3, ...):
    # names gaitk lacks fall through to the reference module, so `from weargait_encoders import ...` keeps working
    ref_enc = importlib.import_module("weargait_encoders") if "weargait_encoders" not in sys.modules else sys.modules["weargait_encoders"]
    if ref_enc is not weargait_encoders:
        for name in dir(ref_enc):
            if not name.startswith("_") and not hasattr(weargait_encoders, name):
                setattr(weargait_encoders, name, getattr(ref_enc, name))
    sys.modules["weargait_encoders"] = weargait_encoders
    ref_fe = importlib.import_module("feature_encoder") if "feature_encoder" not in sys.modules else sys.modules["feature_encoder"]
    if ref_fe is not feature_encoder:
        for name in dir(ref_fe):
            if not name.startswith("_") and not hasattr(feature_encoder, name):
                setattr(feature_encoder, name, getattr(ref_fe, name))
    sys.modules["feature_encoder"] = feature_encoder
    cl = importlib.import_module("learning.optimizers.classification_losses")
    mw = importlib.import_module("learning.optimizers.multitask_weighting")
    cl.GCLLoss, cl.LDAMLoss = classification_losses.GCLLoss, classification_losses.LDAMLoss
    mw.CAGrad = multitask_weighting.CAGrad
    if data_path:
        import data_processing                                           # the reference package (for its other members)
        sys.modules["data_processing.dataloader_weargait"] = dataloader_weargait
        data_processing.dataloader_weargait = dataloader_weargait
        sys.modules["data_processing.dataloader_fbg_fog"] = dataloader_fbg_fog
        data_processing.dataloader_fbg_fog = dataloader_fbg_fog
    for mod in ("weargait_train", "fbg_fog_train", "utilities"):            # (re)import the trainers against the shadowed names
        sys.modules.pop(mod, None)
    return importlib.import_module("weargait_train")
