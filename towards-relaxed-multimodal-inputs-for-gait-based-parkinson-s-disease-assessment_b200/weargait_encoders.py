"""Drop-in for the reference's ``data/WearGait/weargait_encoders.py`` (:19-189): same class names,
constructor signatures, attribute names, ``state_dict`` keys/shapes and -- because the same torch
modules are instantiated in the same order -- the same seeded initial weights.  The arithmetic does not
run in these modules: ``WearGaitThreeModal.forward`` hands the three (B,T,D) streams to the fused
sm_100a kernels of libgaitk.so through ``plan.run_streams``.

Sub-modules stay callable the way ``weargait_train._single_logits_and_labels`` (:262-270) calls them
(``model.head_w(model.backbone(model.enc_w(x)).flatten(1))``): an encoder returns a deferred token, the
head resolves it with one fused single-stream launch.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import _lib
from .plan import FlatParamModule, run_streams


class _Deferred:
    """Token for 'stream s of `owner` applied to x' travelling enc -> backbone -> .flatten(1) -> head."""
    def __init__(self, owner, stream, x, stage="enc"):
        self.owner, self.stream, self.x, self.stage = owner, stream, x, stage

    def flatten(self, start_dim=1):
        return self


class _Part(nn.Module):
    _owner = None
    _stream = None

    def _bind(self, owner, stream):
        object.__setattr__(self, "_owner", weakref.ref(owner)); object.__setattr__(self, "_stream", stream)


class CosineLinear(nn.Module):
    """weargait_encoders.py:19-28."""
    def __init__(self, in_features: int, out_features: int, eps: float = 1e-8):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)
        self.eps = eps


class TaskHead(_Part):
    """weargait_encoders.py:30-37."""
    def __init__(self, input_dim: int, num_classes: int, use_norm: bool = False, use_cosine: bool = False):
        super().__init__()
        self.norm = nn.LayerNorm(input_dim) if (use_norm or use_cosine) else None
        self.fc = CosineLinear(input_dim, num_classes) if use_cosine else nn.Linear(input_dim, num_classes)

    def forward(self, d):
        if not isinstance(d, _Deferred) or d.stage != "backbone":
            raise _lib.GaitkError("TaskHead consumes the deferred output of backbone(enc(x)).flatten(1); "
                                  "call the model (or model.forward_stream) instead of isolated sub-modules")
        owner = d.owner
        if not owner.synchronized and owner._heads[d.stream] is not self:
            raise _lib.GaitkError("this head belongs to another stream")
        return owner.forward_stream(d.stream, d.x)


class WalkwayEncoder(_Part):
    """weargait_encoders.py:40-52."""
    def __init__(self, out_ch: int):
        super().__init__()
        self.conv = nn.Conv1d(2, out_ch, kernel_size=3, padding=1)
        self.act = nn.GELU()
        self.ln = nn.LayerNorm(out_ch)

    def forward(self, x):
        return _Deferred(self._owner(), self._stream, x)


class IMUEncoderShallow(_Part):
    """weargait_encoders.py:54-69."""
    def __init__(self, in_ch: int, out_ch: int, pool_len=None):
        super().__init__()
        if pool_len:
            raise _lib.GaitkError("pool_len is not supported (the reference trainer always passes None, weargait_train.py:468)")
        self.pool_len = pool_len
        self.conv = nn.Conv1d(in_ch, out_ch, kernel_size=3, padding=1)
        self.act = nn.GELU()
        self.ln = nn.LayerNorm(out_ch)
        self.pool = None

    def forward(self, x):
        return _Deferred(self._owner(), self._stream, x)


class InsoleEncoderDeep(_Part):
    """weargait_encoders.py:71-101 (``ln1`` is constructed and never applied, as in the reference)."""
    def __init__(self, in_ch: int, out_ch: int, hidden_ch=None, pool_len=None):
        super().__init__()
        if pool_len:
            raise _lib.GaitkError("pool_len is not supported (the reference trainer always passes None)")
        self.pool_len = pool_len
        h = hidden_ch or max(out_ch, 2 * out_ch)
        self.conv1 = nn.Conv1d(in_ch, h, kernel_size=5, padding=2)
        self.act1 = nn.GELU()
        self.ln1 = nn.LayerNorm(h)
        self.conv2 = nn.Conv1d(h, out_ch, kernel_size=3, padding=1)
        self.act2 = nn.GELU()
        self.ln2 = nn.LayerNorm(out_ch)
        self.skip = nn.Conv1d(h, out_ch, kernel_size=1) if h != out_ch else nn.Identity()
        self.pool = None

    def forward(self, x):
        return _Deferred(self._owner(), self._stream, x)


class SharedBackbone(_Part):
    """weargait_encoders.py:103-113."""
    def __init__(self, in_ch: int, out_ch: int = 16, bdim: int = 8):
        super().__init__()
        self.conv = nn.Conv1d(in_ch, out_ch, kernel_size=3, padding=1)
        self.act = nn.ReLU()
        self.pool = nn.AdaptiveAvgPool1d(bdim)

    def forward(self, d):
        if not isinstance(d, _Deferred) or d.stage != "enc":
            raise _lib.GaitkError("SharedBackbone consumes the deferred output of an encoder of the same model")
        return _Deferred(d.owner, d.stream, d.x, "backbone")


class WearGaitThreeModal(FlatParamModule):
    """weargait_encoders.py:116-189.  forward(x_walk (B,T,2), x_insole (B,T,13), x_imu (B,T,24)) ->
    (lw, li, lm), each (B, K) fp32 and autograd-connected to every parameter."""

    def __init__(self, *, enc_out_ch=12, backbone_dim=8, shared_out_ch=16, num_classes=2, use_norm=False,
                 use_cosine=False, synchronized=True, pool_len=None):
        super().__init__()
        self.enc_w = WalkwayEncoder(out_ch=enc_out_ch)
        self.enc_i = InsoleEncoderDeep(in_ch=13, out_ch=enc_out_ch, hidden_ch=enc_out_ch * 2, pool_len=pool_len)
        self.enc_m = IMUEncoderShallow(in_ch=24, out_ch=enc_out_ch, pool_len=pool_len)
        self.backbone = SharedBackbone(in_ch=enc_out_ch, out_ch=shared_out_ch, bdim=backbone_dim)
        feat_dim = shared_out_ch * backbone_dim
        self.synchronized = synchronized
        if synchronized:
            shared = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
            self.head_w = self.head_i = self.head_m = shared
            self._shared_head = shared
        else:
            self.head_w = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
            self.head_i = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
            self.head_m = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
            self._shared_head = None
        self._cfg = dict(enc_out_ch=enc_out_ch, backbone_dim=backbone_dim, shared_out_ch=shared_out_ch,
                         num_classes=num_classes, use_norm=bool(use_norm), use_cosine=bool(use_cosine))
        self._T = None
        for s, e in enumerate((self.enc_w, self.enc_i, self.enc_m)):
            e._bind(self, s)
        self.backbone._bind(self, None)
        for s, h in enumerate((self.head_w, self.head_i, self.head_m)):
            h._bind(self, s)
        object.__setattr__(self, "_heads", (self.head_w, self.head_i, self.head_m))

    # ---- plan plumbing
    def _plan_kwargs(self):
        if self._T is None:
            raise _lib.GaitkError("window length unknown: call the model on a batch first")
        c = self._cfg
        return dict(family=_lib.FAMILY_WEARGAIT, T=self._T, enc_out_ch=c["enc_out_ch"], shared_out_ch=c["shared_out_ch"],
                    backbone_dim=c["backbone_dim"], num_classes=c["num_classes"], use_norm=c["use_norm"],
                    use_cosine=c["use_cosine"], synchronized=self.synchronized)

    def set_window(self, T: int):
        if self._T != T:
            self._T = int(T)
            object.__setattr__(self, "_plan", None); object.__setattr__(self, "_flat", None)
        return self

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        object.__setattr__(self, "_flat", None)      # .to()/.cuda()/.float() re-materialise parameters
        return out

    # ---- forward
    def forward(self, x_walk, x_insole, x_imu, enabled=(True, True, True)):
        self.set_window(x_walk.shape[1])
        mask = sum(1 << s for s, e in enumerate(enabled) if e)
        lw, li, lm = run_streams(self, [x_walk, x_insole, x_imu], mask)
        return lw, li, lm

    def forward_stream(self, stream: int, x):
        """One branch only: enc_s -> backbone -> head_s (weargait_train.py:262-270)."""
        self.set_window(x.shape[1])
        xs = [None, None, None]; xs[stream] = x
        return run_streams(self, xs, 0b111)[stream]

    # ---- parameter groups (weargait_encoders.py:159-189)
    def walkway_parameters(self):
        params = list(self.enc_w.parameters())
        if not self.synchronized:
            params += list(self.head_w.parameters())
        return params

    def insole_parameters(self):
        params = list(self.enc_i.parameters())
        if not self.synchronized:
            params += list(self.head_i.parameters())
        return params

    def imu_parameters(self):
        params = list(self.enc_m.parameters())
        if not self.synchronized:
            params += list(self.head_m.parameters())
        return params

    def get_shared_parameters(self):
        params = list(self.backbone.parameters())
        if self._shared_head is not None:
            params += list(self._shared_head.parameters())
        return params


# ---------------------------------------------------------------------------------------------------------------
# Fusion baselines reachable through ``weargait_train.py --baseline`` (weargait_encoders.py:199-322)
def _shared_or_three_heads(feat_dim, num_classes, synchronized, use_norm, use_cosine):
    """weargait_encoders.py:199-207."""
    if synchronized:
        shared = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
        return shared, shared, shared, shared
    hw = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
    hi = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
    hm = TaskHead(feat_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
    return None, hw, hi, hm


class _ThreeStreamBase(FlatParamModule):
    """Common plumbing of the 3-stream fusion baselines (same plan family as WearGaitThreeModal)."""
    _proj_ch = 0

    def _build(self, enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm, use_cosine, synchronized, proj_ch=0):
        self.synchronized = synchronized
        self.enc_w = WalkwayEncoder(out_ch=enc_out_ch)
        self.enc_i = InsoleEncoderDeep(in_ch=13, out_ch=enc_out_ch)
        self.enc_m = IMUEncoderShallow(in_ch=24, out_ch=enc_out_ch)
        if proj_ch:
            self.proj_w = nn.Linear(enc_out_ch, proj_ch)
            self.proj_i = nn.Linear(enc_out_ch, proj_ch)
            self.proj_m = nn.Linear(enc_out_ch, proj_ch)
        self.backbone = SharedBackbone(in_ch=proj_ch or enc_out_ch, out_ch=shared_out_ch, bdim=backbone_dim)
        feat_dim = shared_out_ch * backbone_dim
        self._shared_head, self.head_w, self.head_i, self.head_m = _shared_or_three_heads(
            feat_dim, num_classes, synchronized, use_norm, use_cosine)
        self._cfg = dict(enc_out_ch=enc_out_ch, backbone_dim=backbone_dim, shared_out_ch=shared_out_ch,
                         num_classes=num_classes, use_norm=bool(use_norm), use_cosine=bool(use_cosine))
        self._proj_ch = int(proj_ch)
        self._T = None

    def _plan_kwargs(self):
        if self._T is None:
            raise _lib.GaitkError("window length unknown: call the model on a batch first")
        c = self._cfg
        return dict(family=_lib.FAMILY_WEARGAIT, T=self._T, enc_out_ch=c["enc_out_ch"], shared_out_ch=c["shared_out_ch"],
                    backbone_dim=c["backbone_dim"], num_classes=c["num_classes"], use_norm=c["use_norm"],
                    use_cosine=c["use_cosine"], synchronized=self.synchronized, proj_ch=self._proj_ch)

    def _plan_name_map(self):
        if not self.synchronized or self._proj_ch:
            return None                       # names coincide (the plan already calls the sync head _shared_head. with proj)
        mp = {pi.name: pi.name for pi in self.plan().params}
        for k in list(mp):
            if k.startswith("head_w."):
                mp[k] = "_shared_head." + k[len("head_w."):]
        return mp

    def set_window(self, T: int):
        if self._T != T:
            self._T = int(T)
            object.__setattr__(self, "_plan", None); object.__setattr__(self, "_flat", None)
        return self

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        object.__setattr__(self, "_flat", None)
        return out

    def _streams(self, xw, xi, xm):
        self.set_window(xw.shape[1])
        return run_streams(self, [xw, xi, xm], 0b111)


class LateFusion3(_ThreeStreamBase):
    """weargait_encoders.py:247-282.  Sync: the shared head sees the mean of the three latent vectors; the head is
    linear, so W * mean_s(r_s) + b == mean_s(W r_s + b): the fused logits are the mean of the per-stream logits of the
    shared-head model, and the backward is the fused stream kernels with dlogits / 3.  Async: per-stream heads on
    per-stream latents (no fusion), i.e. the three-stream model itself."""

    def __init__(self, enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm=False, use_cosine=False,
                 synchronized=True):
        super().__init__()
        if synchronized and (use_norm or use_cosine):
            raise _lib.GaitkError("LateFusion3 with a normalising head is not linear in the latent; the trainer never builds it "
                                  "(weargait_train.py:513-524 passes neither use_norm nor use_cosine)")
        self._build(enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm, use_cosine, synchronized)

    @property
    def fuses_streams(self) -> bool:
        """sync: every returned logit is the shared head on the MEAN latent, so masking one input changes all three
        (weargait_encoders.py:272-279) -- the one-pass mask evaluation does not apply (evaluation.streams_are_independent)"""
        return bool(self.synchronized)

    def forward(self, xw, xi, xm):
        lw, li, lm = self._streams(xw, xi, xm)
        if self.synchronized:
            logits = (lw + li + lm) / 3.0
            return logits, logits, logits
        return lw, li, lm


class SharedLatent3(_ThreeStreamBase):
    """weargait_encoders.py:284-322: encoder -> per-stream Linear(enc_out_ch -> proj_ch) -> shared backbone -> head(s).
    The projection is an extra stage inside the fused stream kernel (fp32 path)."""

    def __init__(self, enc_out_ch, proj_ch, backbone_dim, shared_out_ch, num_classes, use_norm=False, use_cosine=False,
                 synchronized=True):
        super().__init__()
        self._build(enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm, use_cosine, synchronized, proj_ch=proj_ch)

    def forward(self, xw, xi, xm):
        return tuple(self._streams(xw, xi, xm))


# ---------------------------------------------------------------------------------------------------------------
# Baselines that couple the streams between encoder and backbone: staged execution (staged.py)
class _StagedThreeStream(nn.Module):
    """EarlyFusion3 / CheapXAttn3: same sub-modules, construction order and ``state_dict`` keys as the reference classes; plain
    ``nn.Parameter``s (any optimizer works), every forward / backward stage a libgaitk.so kernel (staged.py)."""

    def _build(self, enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm, use_cosine, synchronized, backbone_in):
        if use_norm or use_cosine:
            raise _lib.GaitkError("the staged fusion baselines use plain linear heads (weargait_train.py:513-524 builds them without "
                                  "use_norm / use_cosine)")
        self.synchronized = synchronized
        self.enc_w = WalkwayEncoder(out_ch=enc_out_ch)
        self.enc_i = InsoleEncoderDeep(in_ch=13, out_ch=enc_out_ch)
        self.enc_m = IMUEncoderShallow(in_ch=24, out_ch=enc_out_ch)
        self._dims = (enc_out_ch, backbone_dim, shared_out_ch)
        return backbone_in

    def _heads(self, backbone_dim, shared_out_ch, num_classes, synchronized):
        self._shared_head, self.head_w, self.head_i, self.head_m = _shared_or_three_heads(
            shared_out_ch * backbone_dim, num_classes, synchronized, False, False)

    def _encode(self, xw, xi, xm):
        from . import staged
        C = self._dims[0]
        ew, ei, em = self.enc_w, self.enc_i, self.enc_m
        fw = staged.encode(_lib.STAGE_CONV_GELU_LN, xw, [ew.conv.weight, ew.conv.bias, ew.ln.weight, ew.ln.bias], C_out=C)
        if isinstance(ei.skip, nn.Identity):
            raise _lib.GaitkError("insole hidden width == output width (identity skip) is not built by the reference defaults")
        fi = staged.encode(_lib.STAGE_INSOLE, xi, [ei.conv1.weight, ei.conv1.bias, ei.conv2.weight, ei.conv2.bias, ei.ln2.weight, ei.ln2.bias,
                                                   ei.skip.weight, ei.skip.bias], C_out=C, H=ei.conv1.out_channels)
        fm = staged.encode(_lib.STAGE_CONV_GELU_LN, xm, [em.conv.weight, em.conv.bias, em.ln.weight, em.ln.bias], C_out=C)
        return fw, fi, fm

    def _repr(self, X):
        from . import staged
        return staged.trunk(X, self.backbone.conv.weight, self.backbone.conv.bias, self._dims[1])

    @staticmethod
    def _head(head, r):
        from . import staged
        return staged.linear(r, head.fc.weight, head.fc.bias)

    @property
    def fuses_streams(self) -> bool:
        return True


class EarlyFusion3(_StagedThreeStream):
    """weargait_encoders.py:209-245: the three encoder outputs concatenated along channels feed ONE backbone (3 C -> S)."""

    def __init__(self, enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm=False, use_cosine=False, synchronized=True):
        super().__init__()
        self._build(enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm, use_cosine, synchronized, 3 * enc_out_ch)
        self.backbone = SharedBackbone(in_ch=3 * enc_out_ch, out_ch=shared_out_ch, bdim=backbone_dim)
        self._heads(backbone_dim, shared_out_ch, num_classes, synchronized)

    def forward(self, xw, xi, xm):
        r = self._repr(torch.cat(self._encode(xw, xi, xm), dim=-1))
        if self.synchronized:
            logits = self._head(self._shared_head, r)
            return logits, logits, logits
        return self._head(self.head_w, r), self._head(self.head_i, r), self._head(self.head_m, r)


class CheapCrossAttention(nn.Module):
    """weargait_encoders.py:324-336 (zero parameters): softmax(A B^T / sqrt(d)) B."""

    def __init__(self, dim: int):
        super().__init__()
        self.scale = dim ** -0.5

    def forward(self, A, B):
        from . import staged
        return staged.cheap_xattn(A, B)


class CheapXAttn3(_StagedThreeStream):
    """weargait_encoders.py:338-387: six pairwise cross attentions, each stream's two attended sequences averaged, shared
    backbone per stream, per-stream (or shared) heads."""

    def __init__(self, enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm=False, use_cosine=False, synchronized=True):
        super().__init__()
        self._build(enc_out_ch, backbone_dim, shared_out_ch, num_classes, use_norm, use_cosine, synchronized, enc_out_ch)
        self.xattn = CheapCrossAttention(dim=enc_out_ch)
        self.backbone = SharedBackbone(in_ch=enc_out_ch, out_ch=shared_out_ch, bdim=backbone_dim)
        self._heads(backbone_dim, shared_out_ch, num_classes, synchronized)

    def forward(self, xw, xi, xm):
        W, I, M = self._encode(xw, xi, xm)
        xa = self.xattn
        W_star = (xa(W, I) + xa(W, M)) * 0.5
        I_star = (xa(I, W) + xa(I, M)) * 0.5
        M_star = (xa(M, W) + xa(M, I)) * 0.5
        return (self._head(self.head_w, self._repr(W_star)), self._head(self.head_i, self._repr(I_star)),
                self._head(self.head_m, self._repr(M_star)))
