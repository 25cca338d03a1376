"""Staged execution of the fusion baselines (EarlyFusion3 / CheapXAttn3, ``data/WearGait/weargait_encoders.py:209-245, 338-387``;
EarlyFusionModel / LateFusionModel / ShareLatentModel / CheapXAttnModel, ``train/feature_encoder.py:346-596``).

These models couple the streams BETWEEN encoder and backbone (channel concat, zero-parameter cross attention, concatenated
latents), so one fused kernel per stream cannot express them.  They run as a chain of CUDA stages with the intermediate
tensors in HBM, every stage an autograd node backed by libgaitk.so:

    encode(...)       one launch of the stream kernel cut after the encoder           gaitk_stage_forward / _backward
    cheap_xattn(A, B) softmax(A B^T / sqrt(d)) B per window                           gaitk_xattn_forward / _backward
    trunk(x, w, b)    SharedBackbone: conv k3 + ReLU + adaptive pooling + flatten     gaitk_stage_forward / _backward (ENC_NONE)
    linear(x, W, b)   heads and per-stream projections                                gaitk_linear_forward / _backward
    FusedAdam         torch.optim.Adam.step as one launch (baselines/fusion_train.py:202)   gaitk_adam

Backward passes recompute the stage's forward inside the kernel (nothing but the stage inputs is saved).  No eager fallback:
every function raises when the inputs are not CUDA tensors or libgaitk.so is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, lib, stream_handle

_STAGES: Dict[Tuple, "Stage"] = {}


class Stage:
    """One stage plan (geometry + kernel lookup live in the library); cached per (description, device)."""

    def __init__(self, enc: int, CIN: int, H: int, Cout: int, T_in: int, T: int, pool: bool, S: int, bdim: int, device: torch.device):
        d = _lib.StageDesc(enc=enc, CIN=CIN, H=H, C=Cout, T_in=T_in, T=T, pool_sensor=1 if pool else 0, S=S, bdim=bdim)
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().gaitk_stage_create(C.byref(d), device.index if device.index is not None else torch.cuda.current_device(), C.byref(h)),
                  "gaitk_stage_create")
        self.handle = h; self.device = device
        self.NP = int(lib().gaitk_param_total(h))
        self.enc = enc; self.CIN = CIN; self.C = Cout; self.T_in = T_in; self.T = T; self.NF = S * bdim
        self._ws: Dict[int, torch.Tensor] = {}

    def workspace(self, B: int) -> torch.Tensor:
        n = int(lib().gaitk_workspace_bytes(self.handle, B))
        w = self._ws.get(B)
        if w is None or w.numel() < n:
            w = torch.empty(n, dtype=torch.uint8, device=self.device); self._ws = {B: w}
        return w

    def __del__(self):
        try:
            lib().gaitk_plan_destroy(self.handle)
        except Exception:
            pass


def get_stage(enc, CIN, H, Cout, T_in, T, pool, S, bdim, device) -> Stage:
    key = (enc, CIN, H, Cout, T_in, T, bool(pool), S, bdim, str(device))
    st = _STAGES.get(key)
    if st is None:
        st = _STAGES[key] = Stage(enc, CIN, H, Cout, T_in, T, pool, S, bdim, device)
    return st


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.dtype == torch.float32):
            raise _lib.GaitkError("gaitk stages take CUDA fp32 tensors (there is no CPU path)")


class _StageFn(torch.autograd.Function):
    """out = stage(x; flat parameters).  Saves only (x, flat): the backward kernel recomputes the forward pass."""

    @staticmethod
    def forward(ctx, stage: Stage, x: torch.Tensor, flat: torch.Tensor):
        _need_cuda(x, flat)
        x = x.contiguous(); flat = flat.contiguous()
        B = x.shape[0]
        out = torch.empty((B, stage.NF) if stage.enc == _lib.STAGE_TRUNK else (B, stage.T, stage.C), dtype=torch.float32, device=x.device)
        check(lib().gaitk_stage_forward(stage.handle, flat.data_ptr(), x.data_ptr(), None, B, 0, out.data_ptr(), stream_handle()),
              "gaitk_stage_forward")
        ctx.stage = stage
        ctx.save_for_backward(x, flat)
        return out

    @staticmethod
    def backward(ctx, dout):
        stage: Stage = ctx.stage
        x, flat = ctx.saved_tensors
        B = x.shape[0]
        dout = dout.contiguous()
        g = torch.zeros(stage.NP + 8, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x) if (stage.enc == _lib.STAGE_TRUNK and ctx.needs_input_grad[1]) else None
        ws = stage.workspace(B)
        check(lib().gaitk_stage_backward(stage.handle, flat.data_ptr(), x.data_ptr(), None, B, 0, dout.data_ptr(),
                                         None if dx is None else dx.data_ptr(), g.data_ptr(), ws.data_ptr(), ws.numel(), stream_handle()),
              "gaitk_stage_backward")
        return None, dx, g[:stage.NP]


def _flat(params: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([p.reshape(-1) for p in params])


def encode(kind: int, x: torch.Tensor, params: Sequence[torch.Tensor], *, C_out: int, H: int = 0, T_out: Optional[int] = None,
           pool: bool = False) -> torch.Tensor:
    """Encoder stage: x (B, T_in, CIN) -> (B, T_out, C_out).  params in the stage's flat order (see gaitk.h: conv encoders
    w1 b1 lng lnb; insole w1 b1 w2 b2 lng lnb wsk bsk; SkeletonMLP w1 b1 lng lnb; SensorEncoder w1 b1)."""
    _need_cuda(x)
    T_in = x.shape[1]; T = T_out or T_in
    st = get_stage(kind, x.shape[2], H, C_out, T_in, T, pool, 16, min(8, T), x.device)
    return _StageFn.apply(st, x, _flat(params))


def trunk(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, bdim: int) -> torch.Tensor:
    """SharedBackbone + flatten: x (B, T, CIN), conv weight (S, CIN, 3) -> (B, bdim * S) with feature index b * S + s
    (weargait_encoders.py:103-113 / feature_encoder.py:80-109)."""
    _need_cuda(x, weight)
    st = get_stage(_lib.STAGE_TRUNK, x.shape[2], 0, x.shape[2], x.shape[1], x.shape[1], False, weight.shape[0], bdim, x.device)
    return _StageFn.apply(st, x, _flat([weight, bias]))


class _XAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A: torch.Tensor, Bm: torch.Tensor):
        _need_cuda(A, Bm)
        A = A.contiguous(); Bm = Bm.contiguous()
        n, T, d = A.shape
        out = torch.empty_like(A)
        check(lib().gaitk_xattn_forward(A.data_ptr(), Bm.data_ptr(), out.data_ptr(), n, T, d, stream_handle()), "gaitk_xattn_forward")
        ctx.save_for_backward(A, Bm)
        return out

    @staticmethod
    def backward(ctx, dout):
        A, Bm = ctx.saved_tensors
        n, T, d = A.shape
        dA = torch.empty_like(A); dB = torch.empty_like(Bm)
        check(lib().gaitk_xattn_backward(A.data_ptr(), Bm.data_ptr(), dout.contiguous().data_ptr(), dA.data_ptr(), dB.data_ptr(), n, T, d,
                                         stream_handle()), "gaitk_xattn_backward")
        return dA, dB


def cheap_xattn(A: torch.Tensor, Bm: torch.Tensor) -> torch.Tensor:
    """CheapCrossAttention.forward (weargait_encoders.py:332-336): softmax(A B^T * d^-0.5) B, (B, T, d) each."""
    return _XAttnFn.apply(A, Bm)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        _need_cuda(x, W, b)
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous(); W = W.contiguous()
        R, I = x2.shape; O = W.shape[0]
        y = torch.empty(R, O, dtype=torch.float32, device=x.device)
        check(lib().gaitk_linear_forward(x2.data_ptr(), W.data_ptr(), None if b is None else b.contiguous().data_ptr(), y.data_ptr(), R, I, O,
                                         stream_handle()), "gaitk_linear_forward")
        ctx.save_for_backward(x2, W); ctx.has_bias = b is not None; ctx.shape = shape
        return y.reshape(shape[:-1] + (O,))

    @staticmethod
    def backward(ctx, dy):
        x2, W = ctx.saved_tensors
        R, I = x2.shape; O = W.shape[0]
        dy2 = dy.reshape(R, O).contiguous()
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(W); db = torch.empty(O, dtype=torch.float32, device=W.device) if ctx.has_bias else None
        n = int(lib().gaitk_linear_workspace_bytes(R, I, O))
        ws = torch.empty(n, dtype=torch.uint8, device=W.device)
        check(lib().gaitk_linear_backward(x2.data_ptr(), W.data_ptr(), dy2.data_ptr(), None if dx is None else dx.data_ptr(), dW.data_ptr(),
                                          None if db is None else db.data_ptr(), R, I, O, ws.data_ptr(), n, stream_handle()),
              "gaitk_linear_backward")
        return (None if dx is None else dx.reshape(ctx.shape)), dW, db


def linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear on the last dimension (in <= 256, out <= 32)."""
    return _LinearFn.apply(x, W, b)


class FusedAdam:
    """torch.optim.Adam(params, lr) as ONE kernel launch per step over all parameter tensors (gaitk_adam); same defaults, same
    arithmetic order (lerp for exp_avg, sqrt(v) / sqrt(1 - b2^t) + eps)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        self.params: List[torch.Tensor] = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.state_step = 0
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        idx = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not idx:
            return
        self.state_step += 1
        ps = [self.params[i] for i in idx]
        _need_cuda(*ps)
        n = len(ps)
        grads = [p.grad.contiguous() for p in ps]
        numel = (C.c_int64 * n)(*[p.numel() for p in ps])
        check(lib().gaitk_adam(_lib.ptr_array([p.data_ptr() for p in ps]), _lib.ptr_array([g.data_ptr() for g in grads]),
                               _lib.ptr_array([self.exp_avg[i].data_ptr() for i in idx]), _lib.ptr_array([self.exp_avg_sq[i].data_ptr() for i in idx]),
                               numel, n, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.state_step, stream_handle()),
              "gaitk_adam")
