"""Drop-in for the reference's ``train/feature_encoder.py`` :7-265 (2-stream FoG / FBG model): same class
names, constructor signatures, attribute names, ``state_dict`` keys and seeded initial weights; the
arithmetic runs in the fused sm_100a kernels of libgaitk.so."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .plan import FlatParamModule, run_streams


class CosineLinear(nn.Module):
    """feature_encoder.py:7-24."""
    def __init__(self, in_features: int, out_features: int, eps: float = 1e-8):
        super().__init__()
        self.weight = nn.Parameter(torch.Tensor(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)
        self.eps = eps


class SensorEncoder(nn.Module):
    """feature_encoder.py:27-58 (parameter container)."""
    def __init__(self, in_channels: int, out_channels: int, sensor_length: int = None, output_length: int = 101):
        super().__init__()
        self.d2_sensor_length = sensor_length
        self.output_length = output_length
        self.conv1d = nn.Conv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        self.pool = nn.AdaptiveAvgPool1d(output_length)


class SkeletonMLP(nn.Module):
    """feature_encoder.py:61-77."""
    def __init__(self, input_dim: int, output_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, output_dim)
        self.ln1 = nn.LayerNorm(output_dim)
        self.relu = nn.ReLU()


class SharedBackbone(nn.Module):
    """feature_encoder.py:80-109."""
    def __init__(self, in_channels: int, shared_out_channels: int = 16, backbone_dim: int = 8):
        super().__init__()
        self.conv1d = nn.Conv1d(in_channels=in_channels, out_channels=shared_out_channels, kernel_size=3, stride=1, padding=1)
        self.relu = nn.ReLU()
        self.pool = nn.AdaptiveAvgPool1d(backbone_dim)


class TaskHead(nn.Module):
    """feature_encoder.py:112-146."""
    def __init__(self, input_dim: int, num_classes: int, use_norm: bool = False, use_cosine: bool = False):
        super().__init__()
        self.use_cosine = use_cosine
        if use_cosine:
            self.norm = nn.LayerNorm(input_dim); self.fc = CosineLinear(input_dim, num_classes)
        elif use_norm:
            self.norm = nn.LayerNorm(input_dim); self.fc = nn.Linear(input_dim, num_classes)
        else:
            self.norm = None; self.fc = nn.Linear(input_dim, num_classes)


class MultiModalMultiTaskModel(FlatParamModule):
    """feature_encoder.py:149-265.  forward(x_skel (B,T,Dskel), x_sensor (B,Tsens,Csens)) ->
    (logits_skel, logits_sensor)."""

    def __init__(self, skeleton_input_dim: int, skeleton_output_dim: int, sensor_in_channels: int,
                 sensor_out_channels: int, sensor_length: int, shared_out_channels: int, backbone_dim: int,
                 taskhead_input_dim: int, num_classes: int, use_norm: bool = False, use_cosine: bool = False,
                 synchronized_loading: bool = False):
        super().__init__()
        if skeleton_output_dim != sensor_out_channels:
            raise _lib.GaitkError("skeleton_output_dim must equal sensor_out_channels (one shared backbone)")
        if taskhead_input_dim != shared_out_channels * backbone_dim:
            raise _lib.GaitkError("taskhead_input_dim must equal shared_out_channels * backbone_dim")
        self.skeleton_encoder = SkeletonMLP(input_dim=skeleton_input_dim, output_dim=skeleton_output_dim)
        self.sensor_encoder = SensorEncoder(in_channels=sensor_in_channels, out_channels=sensor_out_channels,
                                            sensor_length=sensor_length)
        self.backbone = SharedBackbone(in_channels=sensor_out_channels, shared_out_channels=shared_out_channels,
                                       backbone_dim=backbone_dim)
        self.synchronized_loading = synchronized_loading
        if synchronized_loading:
            self.task_head_shared = TaskHead(taskhead_input_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
        else:
            self.task_head_skel = TaskHead(taskhead_input_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
            self.task_head_sensor = TaskHead(taskhead_input_dim, num_classes, use_norm=use_norm, use_cosine=use_cosine)
        self.use_skeleton_only = False
        self.use_sensor_only = False
        self._cfg = dict(skel_in_dim=skeleton_input_dim, enc_out_ch=skeleton_output_dim, sensor_in_ch=sensor_in_channels,
                         sensor_len=sensor_length, shared_out_ch=shared_out_channels, backbone_dim=backbone_dim,
                         num_classes=num_classes, use_norm=bool(use_norm), use_cosine=bool(use_cosine))
        self._T = None

    def _plan_kwargs(self):
        if self._T is None:
            raise _lib.GaitkError("pose length unknown: call the model on a batch first")
        c = self._cfg
        return dict(family=_lib.FAMILY_FOG, T=self._T, enc_out_ch=c["enc_out_ch"], shared_out_ch=c["shared_out_ch"],
                    backbone_dim=c["backbone_dim"], num_classes=c["num_classes"], use_norm=c["use_norm"],
                    use_cosine=c["use_cosine"], synchronized=self.synchronized_loading, skel_in_dim=c["skel_in_dim"],
                    sensor_in_ch=c["sensor_in_ch"], sensor_len=c["sensor_len"],
                    sensor_out_len=self.sensor_encoder.output_length)

    def set_window(self, T: int):
        if self._T != T:
            self._T = int(T)
            object.__setattr__(self, "_plan", None); object.__setattr__(self, "_flat", None)
        return self

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        object.__setattr__(self, "_flat", None)
        return out

    def forward(self, x_skel, x_sensor):
        self.set_window(x_skel.shape[1])
        if x_sensor.shape[1] != self._cfg["sensor_len"]:
            raise _lib.GaitkError(f"sensor clips must have sensor_length={self._cfg['sensor_len']} frames "
                                  "(create_fusion_loaders pads/trims to it, dataloader_fbg_fog.py:135,158)")
        if x_skel.shape[1] != self.sensor_encoder.output_length:
            raise _lib.GaitkError("pose length must equal the sensor encoder's pooled length (101)")
        if self.use_skeleton_only:
            ls, _ = run_streams(self, [x_skel, None], 0b11)
            return ls, None
        if self.use_sensor_only:
            _, lt = run_streams(self, [None, x_sensor], 0b11)
            return None, lt
        ls, lt = run_streams(self, [x_skel, x_sensor], 0b11)
        return ls, lt

    def get_shared_parameters(self):
        shared = list(self.backbone.parameters())
        if self.synchronized_loading:
            shared += list(self.task_head_shared.parameters())
        return shared


class _SingleStreamModel(FlatParamModule):
    """One stream of the 2-stream FoG/FBG plan (the other stream's parameters are zero dummies, never launched)."""
    _stream = 0

    def _plan_kwargs(self):
        c = self._cfg
        return dict(family=_lib.FAMILY_FOG, T=c["T"], enc_out_ch=c["enc_out_ch"], shared_out_ch=c["shared_out_ch"],
                    backbone_dim=c["backbone_dim"], num_classes=c["num_classes"], use_norm=c["use_norm"], use_cosine=False,
                    synchronized=False, skel_in_dim=c["skel_in_dim"], sensor_in_ch=c["sensor_in_ch"],
                    sensor_len=c["sensor_len"], sensor_out_len=c["T"])

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        object.__setattr__(self, "_flat", None)
        return out

    def forward(self, x):
        xs = [None, None]; xs[self._stream] = x
        return run_streams(self, xs, 0b11)[self._stream]


class SensorModalityModel(_SingleStreamModel):
    """feature_encoder.py:268-305: SensorEncoder -> SharedBackbone -> TaskHead (use_norm=True by default)."""
    _stream = 1

    def __init__(self, sensor_in_channels: int, sensor_out_channels: int, sensor_length: int, shared_out_channels: int,
                 backbone_dim: int, taskhead_input_dim: int, num_classes: int, use_norm: bool = True):
        super().__init__()
        self.encoder = SensorEncoder(in_channels=sensor_in_channels, out_channels=sensor_out_channels, sensor_length=sensor_length)
        self.backbone = SharedBackbone(in_channels=sensor_out_channels, shared_out_channels=shared_out_channels, backbone_dim=backbone_dim)
        self.task_head = TaskHead(input_dim=taskhead_input_dim, num_classes=num_classes, use_norm=use_norm, use_cosine=False)
        self._cfg = dict(T=self.encoder.output_length, enc_out_ch=sensor_out_channels, shared_out_ch=shared_out_channels,
                         backbone_dim=backbone_dim, num_classes=num_classes, use_norm=bool(use_norm),
                         skel_in_dim=21 if sensor_in_channels == 6 else 51, sensor_in_ch=sensor_in_channels, sensor_len=sensor_length)

    def _plan_name_map(self):
        mp = {"sensor_encoder.conv1d.weight": "encoder.conv1d.weight", "sensor_encoder.conv1d.bias": "encoder.conv1d.bias",
              "backbone.conv1d.weight": "backbone.conv1d.weight", "backbone.conv1d.bias": "backbone.conv1d.bias"}
        for k in ("norm.weight", "norm.bias", "fc.weight", "fc.bias"):
            mp["task_head_sensor." + k] = "task_head." + k
        return mp

    def forward(self, x):
        if x.shape[1] != self._cfg["sensor_len"]:
            raise _lib.GaitkError(f"sensor clips must have sensor_length={self._cfg['sensor_len']} frames")
        return super().forward(x)


class SkelModalityModel(_SingleStreamModel):
    """feature_encoder.py:308-344: SkeletonMLP -> SharedBackbone -> TaskHead (use_norm=True by default)."""
    _stream = 0

    def __init__(self, skeleton_input_dim: int, skeleton_output_dim: int, sensor_out_channels: int, shared_out_channels: int,
                 backbone_dim: int, taskhead_input_dim: int, num_classes: int, use_norm: bool = True):
        super().__init__()
        self.encoder = SkeletonMLP(input_dim=skeleton_input_dim, output_dim=skeleton_output_dim)
        self.backbone = SharedBackbone(in_channels=sensor_out_channels, shared_out_channels=shared_out_channels, backbone_dim=backbone_dim)
        self.task_head = TaskHead(input_dim=taskhead_input_dim, num_classes=num_classes, use_norm=use_norm, use_cosine=False)
        fog = skeleton_input_dim == 21
        self._cfg = dict(T=101, enc_out_ch=skeleton_output_dim, shared_out_ch=shared_out_channels, backbone_dim=backbone_dim,
                         num_classes=num_classes, use_norm=bool(use_norm), skel_in_dim=skeleton_input_dim,
                         sensor_in_ch=6 if fog else 3, sensor_len=426 if fog else 65)

    def _plan_name_map(self):
        mp = {"skeleton_encoder.fc1.weight": "encoder.fc1.weight", "skeleton_encoder.fc1.bias": "encoder.fc1.bias",
              "skeleton_encoder.ln1.weight": "encoder.ln1.weight", "skeleton_encoder.ln1.bias": "encoder.ln1.bias",
              "backbone.conv1d.weight": "backbone.conv1d.weight", "backbone.conv1d.bias": "backbone.conv1d.bias"}
        for k in ("norm.weight", "norm.bias", "fc.weight", "fc.bias"):
            mp["task_head_skel." + k] = "task_head." + k
        return mp

    def forward(self, x):
        self._cfg["T"] = int(x.shape[1])
        return super().forward(x)


# ---------------------------------------------------------------------------------------------------------------
# 2-stream fusion baselines (feature_encoder.py:346-596, trained by baselines/fusion_train.py): staged execution (staged.py)
class _StagedTwoStream(nn.Module):
    """Same sub-modules, construction order and ``state_dict`` keys as the reference classes; every stage a libgaitk.so kernel."""

    def _encoders(self, skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length):
        self.skel_enc = SkeletonMLP(skeleton_input_dim, skeleton_output_dim)
        self.sens_enc = SensorEncoder(in_channels=sensor_in_channels, out_channels=sensor_out_channels, sensor_length=sensor_length)

    def _encode(self, x_skel, x_sens):
        from . import staged
        k, e = self.skel_enc, self.sens_enc
        sk = staged.encode(_lib.STAGE_LINEAR_LN_RELU, x_skel, [k.fc1.weight, k.fc1.bias, k.ln1.weight, k.ln1.bias], C_out=k.fc1.out_features)
        pool = x_sens.shape[1] == e.d2_sensor_length                        # SensorEncoder pools only then (feature_encoder.py:55)
        se = staged.encode(_lib.STAGE_CONV_POOL, x_sens, [e.conv1d.weight, e.conv1d.bias], C_out=e.conv1d.out_channels,
                           T_out=e.output_length if pool else None, pool=pool)
        return sk, se

    def _repr(self, X):
        from . import staged
        return staged.trunk(X, self.backbone.conv1d.weight, self.backbone.conv1d.bias, self._bdim)

    def _out(self, r_skel, r_sens=None):
        from . import staged
        r_sens = r_skel if r_sens is None else r_sens
        if self.synchronized_loading:
            return staged.linear(r_skel, self.head.weight, self.head.bias)
        return staged.linear(r_skel, self.head_skel.weight, self.head_skel.bias), staged.linear(r_sens, self.head_sens.weight, self.head_sens.bias)

    def _make_heads(self, feature_dim, num_classes, synchronized_loading):
        if synchronized_loading:
            self.head = nn.Linear(feature_dim, num_classes)
        else:
            self.head_skel = nn.Linear(feature_dim, num_classes)
            self.head_sens = nn.Linear(feature_dim, num_classes)


class EarlyFusionModel(_StagedTwoStream):
    """feature_encoder.py:347-396: channel concat of the two encoder outputs -> one backbone -> head(s)."""

    def __init__(self, skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length, shared_out_channels,
                 backbone_dim, num_classes, synchronized_loading=False):
        super().__init__()
        self.synchronized_loading = synchronized_loading; self._bdim = backbone_dim
        self._encoders(skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length)
        self.backbone = SharedBackbone(in_channels=skeleton_output_dim + sensor_out_channels, shared_out_channels=shared_out_channels,
                                       backbone_dim=backbone_dim)
        self._make_heads(backbone_dim * shared_out_channels, num_classes, synchronized_loading)

    def forward(self, x_skel, x_sens):
        return self._out(self._repr(torch.cat(self._encode(x_skel, x_sens), dim=-1)))


class LateFusionModel(_StagedTwoStream):
    """feature_encoder.py:399-446: one backbone applied to each stream, the two latent vectors concatenated (2 x 128) -> head(s)."""

    def __init__(self, skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length, shared_out_channels,
                 backbone_dim, num_classes, synchronized_loading=False):
        super().__init__()
        self.synchronized_loading = synchronized_loading; self._bdim = backbone_dim
        self._encoders(skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length)
        self.backbone = SharedBackbone(in_channels=skeleton_output_dim, shared_out_channels=shared_out_channels, backbone_dim=backbone_dim)
        self._make_heads(2 * backbone_dim * shared_out_channels, num_classes, synchronized_loading)

    def forward(self, x_skel, x_sens):
        sk, se = self._encode(x_skel, x_sens)
        return self._out(torch.cat([self._repr(sk), self._repr(se)], dim=1))


class ShareLatentModel(_StagedTwoStream):
    """feature_encoder.py:449-494: per-stream Linear projection into a common width, shared backbone, ONE head on each latent."""

    def __init__(self, skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length, shared_out_channels,
                 backbone_dim, taskhead_input_dim, num_classes, synchronized_loading=False):
        super().__init__()
        self.synchronized_loading = synchronized_loading; self._bdim = backbone_dim
        self._encoders(skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length)
        self.proj_skel = nn.Linear(skeleton_output_dim, shared_out_channels)
        self.proj_sens = nn.Linear(sensor_out_channels, shared_out_channels)
        self.backbone = SharedBackbone(in_channels=shared_out_channels, shared_out_channels=shared_out_channels, backbone_dim=backbone_dim)
        self.head = nn.Linear(backbone_dim * shared_out_channels, num_classes)

    def forward(self, x_skel, x_sens):
        from . import staged
        sk, se = self._encode(x_skel, x_sens)
        r_sk = self._repr(staged.linear(sk, self.proj_skel.weight, self.proj_skel.bias))
        r_se = self._repr(staged.linear(se, self.proj_sens.weight, self.proj_sens.bias))
        return staged.linear(r_sk, self.head.weight, self.head.bias), staged.linear(r_se, self.head.weight, self.head.bias)


class CheapCrossAttention(nn.Module):
    """feature_encoder.py:497-528 (symmetric, zero parameters): 0.5 (softmax(S G^T / sqrt d) G + softmax(G S^T / sqrt d) S)."""

    def __init__(self, dim: int):
        super().__init__()
        self.scale = dim ** -0.5

    def forward(self, S, G):
        from . import staged
        return (staged.cheap_xattn(S, G) + staged.cheap_xattn(G, S)) * 0.5


class CheapXAttnModel(_StagedTwoStream):
    """feature_encoder.py:531-596: symmetric cross attention fuses the two encoded sequences -> backbone -> head(s)."""

    def __init__(self, skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length, shared_out_channels,
                 backbone_dim, num_classes, synchronized_loading=False):
        super().__init__()
        if skeleton_output_dim != sensor_out_channels:
            raise _lib.GaitkError("cross attention needs the same feature width on both modalities")
        self.synchronized_loading = synchronized_loading; self._bdim = backbone_dim
        self._encoders(skeleton_input_dim, skeleton_output_dim, sensor_in_channels, sensor_out_channels, sensor_length)
        self.cross_attn = CheapCrossAttention(dim=skeleton_output_dim)
        self.backbone = SharedBackbone(in_channels=skeleton_output_dim, shared_out_channels=shared_out_channels, backbone_dim=backbone_dim)
        self._make_heads(backbone_dim * shared_out_channels, num_classes, synchronized_loading)

    def forward(self, x_skel, x_sens):
        sk, se = self._encode(x_skel, x_sens)
        return self._out(self._repr(self.cross_attn(sk, se)))
