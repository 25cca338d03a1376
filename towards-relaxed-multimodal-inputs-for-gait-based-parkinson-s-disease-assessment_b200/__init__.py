"""gaitk -- B200 (sm_100a) implementation of the gait training hot path behind the reference's own
module API.  See DESIGN.md / INTEGRATION.md at the repository root."""
from . import _lib
from ._lib import GaitkError, lib, DTYPE_F32, DTYPE_TF32, DTYPE_BF16X3, SOLVER_SLSQP, SOLVER_EXACT, SOLVER_MEAN
from .plan import Plan, FlatParamModule
from .weargait_encoders import WearGaitThreeModal, LateFusion3, SharedLatent3, EarlyFusion3, CheapXAttn3
from .feature_encoder import (MultiModalMultiTaskModel, SensorModalityModel, SkelModalityModel, EarlyFusionModel, LateFusionModel,
                              ShareLatentModel, CheapXAttnModel)
from . import staged
from .staged import FusedAdam
from .classification_losses import GCLLoss, LDAMLoss, CrossEntropyLoss, make_loss_desc, criterion_spec
from .multitask_weighting import CAGrad
from .fused_step import FusedTrainStep
from . import dataloader_weargait
from . import dataloader_fbg_fog
from . import evaluation
from .evaluation import MASK_COMBOS, eval_all_masks, eval_with_mask, eval_one_epoch
from . import dist
from . import integration

__all__ = ["GaitkError", "lib", "Plan", "FlatParamModule", "WearGaitThreeModal", "LateFusion3", "SharedLatent3", "EarlyFusion3", "CheapXAttn3", "EarlyFusionModel", "LateFusionModel", "ShareLatentModel", "CheapXAttnModel", "FusedAdam", "MultiModalMultiTaskModel", "SensorModalityModel", "SkelModalityModel",
           "GCLLoss", "LDAMLoss", "CrossEntropyLoss", "make_loss_desc", "criterion_spec", "CAGrad", "FusedTrainStep"]
