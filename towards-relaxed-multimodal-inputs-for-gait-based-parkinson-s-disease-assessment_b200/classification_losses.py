"""Drop-in for ``train/learning/optimizers/classification_losses.py`` (GCLLoss :79-109, LDAMLoss :54-76)
and for ``nn.CrossEntropyLoss(weight)`` as the trainers use it (weargait_train.py:121-130).  One CUDA
kernel (gaitk_loss) computes the weighted-mean margin/scale cross-entropy, the argmax-correct count and
d loss / d logits; autograd only scales the saved gradient."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
from torch.distributions import normal

from . import _lib
from ._lib import LossDesc, check, lib, stream_handle


def make_loss_desc(K: int, scale: float = 1.0, margin: Optional[Sequence[float]] = None,
                   weight: Optional[Sequence[float]] = None, nan_if_degenerate: bool = False) -> LossDesc:
    d = LossDesc()
    d.scale = float(scale)
    for k in range(_lib.MAX_CLASSES):
        d.margin[k] = float(margin[k]) if (margin is not None and k < K) else 0.0
        d.cls_weight[k] = (float(weight[k]) if weight is not None else 1.0) if k < K else 0.0
    d.nan_if_degenerate = int(bool(nan_if_degenerate))
    return d


class _CELossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, desc, logit_off, stats):
        if not logits.is_cuda:
            raise _lib.GaitkError("gaitk losses run on CUDA only (no CPU path)")
        lg = logits.contiguous().float(); y = target.contiguous().long()
        B, K = lg.shape
        # a 0-dim tensor of its own (NOT a view of a buffer): the trainers update losses in place (`l_skel += ...`,
        # fbg_fog_train.py:123), which autograd forbids on a view returned by a custom Function
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        correct = torch.empty(1, dtype=torch.int32, device=lg.device)
        dlog = torch.empty_like(lg)
        off = None if logit_off is None else logit_off.contiguous().float()
        check(lib().gaitk_loss(lg.data_ptr(), y.data_ptr(), B, K, C.byref(desc), 0 if off is None else off.data_ptr(),
                               loss.data_ptr(), correct.data_ptr(), dlog.data_ptr(), stream_handle()), "gaitk_loss")
        ctx.save_for_backward(dlog)
        if stats is not None:
            stats["correct"] = correct
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        return dlog * g, None, None, None, None


def _weights_to_host(weight, K):
    if weight is None:
        return None
    if torch.is_tensor(weight):
        weight = weight.detach().float().cpu().tolist()
    return [float(w) for w in weight][:K]


class _MarginCE(nn.Module):
    """Shared machinery: cached host copy of the (mutable, DRW) class weights."""
    def __init__(self):
        super().__init__()
        self._wkey = None; self._whost = None
        self.last_stats = {}

    def _host_weight(self, K):
        w = self.weight
        key = None if w is None else ((id(w), w._version) if torch.is_tensor(w) else id(w))
        if key != self._wkey:
            self._whost = _weights_to_host(w, K); self._wkey = key
        return self._whost

    def loss_desc(self, K: int) -> LossDesc:  # pragma: no cover - abstract
        raise NotImplementedError

    def logit_offset(self, logits):
        return None

    def forward(self, logits, target):
        desc = self.loss_desc(logits.shape[1])
        return _CELossFn.apply(logits, target, desc, self.logit_offset(logits), self.last_stats)


class CrossEntropyLoss(_MarginCE):
    """nn.CrossEntropyLoss(weight=...) (mean reduction)."""
    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight

    def loss_desc(self, K):
        return make_loss_desc(K, 1.0, None, self._host_weight(K))


class GCLLoss(_MarginCE):
    """classification_losses.py:79-109.  cls_num_list -> m_list = max(log n) - log n (only its max and
    ratios enter through the noise term); margin ``m`` on the true class, scale ``s``; clamped N(0,1/3)
    noise is drawn from the CPU RNG every call exactly as the reference does (:101), so the global RNG
    stream stays aligned even when noise_mul == 0.  Equal class counts give 0/0 = NaN (:104), reproduced."""

    def __init__(self, cls_num_list, m=0.5, weight=None, s=30, train_cls=False, noise_mul=1., gamma=0.):
        super().__init__()
        if train_cls:
            raise _lib.GaitkError("train_cls=True (focal re-weighting) is not used by the trainers and not implemented")
        cl = np.asarray(cls_num_list, dtype=np.float32)
        ml = np.log(cl); ml = ml.max() - ml
        self._m_list_host = ml.astype(np.float32)
        self.m_list = torch.tensor(ml, dtype=torch.float32, device="cuda" if torch.cuda.is_available() else "cpu")
        assert s > 0
        self.m = m; self.s = s; self.weight = weight
        self.simpler = normal.Normal(0, 1 / 3)
        self.train_cls = train_cls; self.noise_mul = noise_mul; self.gamma = gamma
        self.consume_rng = True

    def degenerate(self) -> bool:
        return float(self._m_list_host.max()) == 0.0

    def loss_desc(self, K):
        return make_loss_desc(K, self.s, [self.m] * K, self._host_weight(K), nan_if_degenerate=self.degenerate())

    def logit_offset(self, logits):
        if not self.consume_rng and self.noise_mul == 0:
            return None
        noise = self.simpler.sample(logits.shape).clamp(-1, 1)
        if self.noise_mul == 0:
            return None
        ml = self.m_list.to(logits.device)
        return self.noise_mul * noise.to(logits.device).abs() / ml.max() * ml


class LDAMLoss(_MarginCE):
    """classification_losses.py:54-76: per-class margin max_m * n^-1/4 / max(n^-1/4), scale s."""
    def __init__(self, cls_num_list, max_m=0.5, weight=None, s=30):
        super().__init__()
        ml = 1.0 / np.sqrt(np.sqrt(np.asarray(cls_num_list, dtype=np.float64)))
        ml = ml * (max_m / np.max(ml))
        self._m_list_host = ml.astype(np.float32)
        self.m_list = torch.tensor(ml, dtype=torch.float32, device="cuda" if torch.cuda.is_available() else "cpu")
        assert s > 0
        self.s = s; self.weight = weight

    def loss_desc(self, K):
        return make_loss_desc(K, self.s, self._m_list_host.tolist(), self._host_weight(K))


def criterion_spec(crit, K: int):
    """(LossDesc, offset_fn) for our criteria, torch's CrossEntropyLoss and duck-typed reference GCL/LDAM."""
    if isinstance(crit, _MarginCE):
        return crit.loss_desc(K), crit.logit_offset
    if isinstance(crit, nn.CrossEntropyLoss):
        if crit.label_smoothing or crit.reduction != "mean" or crit.ignore_index != -100:
            raise _lib.GaitkError("only plain mean-reduced CrossEntropyLoss(weight) is supported")
        return make_loss_desc(K, 1.0, None, _weights_to_host(crit.weight, K)), (lambda lg: None)
    if hasattr(crit, "m_list") and hasattr(crit, "s"):
        ml = crit.m_list.detach().float().cpu().numpy()
        w = _weights_to_host(getattr(crit, "weight", None), K)
        if hasattr(crit, "noise_mul"):     # reference GCLLoss
            if crit.noise_mul != 0:
                raise _lib.GaitkError("reference GCLLoss with noise_mul != 0: use gaitk's GCLLoss")
            return make_loss_desc(K, crit.s, [crit.m] * K, w, nan_if_degenerate=float(ml.max()) == 0.0), (lambda lg: None)
        return make_loss_desc(K, crit.s, ml.tolist(), w), (lambda lg: None)
    raise _lib.GaitkError(f"unsupported criterion {type(crit).__name__}")
