"""Plan = one model's shapes, canonical flat parameter layout and launch geometry (gaitk_plan_*), plus the
autograd bridge that lets unmodified trainer code call ``.backward()`` / ``autograd.grad`` on the logits."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, lib, ptr_array, stream_handle


@dataclass
class ParamInfo:
    name: str
    offset: int
    numel: int
    group: int            # 0 shared (CAGrad), 1+s private to stream s, -1 never gets a gradient
    shape: Tuple[int, ...]


class Plan:
    def __init__(self, *, family: int, T: int, enc_out_ch: int, shared_out_ch: int, backbone_dim: int,
                 num_classes: int, use_norm=False, use_cosine=False, synchronized=True, skel_in_dim=0,
                 sensor_in_ch=0, sensor_len=0, sensor_out_len=0, proj_ch=0, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.GaitkError("gaitk needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = _lib.ModelDesc(family=family, T=T, enc_out_ch=enc_out_ch, shared_out_ch=shared_out_ch,
                           backbone_dim=backbone_dim, num_classes=num_classes, use_norm=int(bool(use_norm)),
                           use_cosine=int(bool(use_cosine)), synchronized=int(bool(synchronized)),
                           skel_in_dim=skel_in_dim, sensor_in_ch=sensor_in_ch, sensor_len=sensor_len,
                           sensor_out_len=sensor_out_len)
        d.reserved[0] = int(proj_ch)
        self.desc = d
        h = C.c_void_p()
        check(lib().gaitk_plan_create(C.byref(d), self.device.index or 0, C.byref(h)), "gaitk_plan_create")
        self._h = h
        L = lib()
        self.params: List[ParamInfo] = []
        name = C.create_string_buffer(128)
        off, num, grp, dims = C.c_int64(), C.c_int64(), C.c_int32(), (C.c_int32 * 4)()
        for i in range(L.gaitk_param_count(h)):
            check(L.gaitk_param_info(h, i, name, 128, C.byref(off), C.byref(num), C.byref(grp), dims))
            shape = tuple(int(x) for x in dims if x > 0)
            self.params.append(ParamInfo(name.value.decode(), off.value, num.value, grp.value, shape))
        self.NP = int(L.gaitk_param_total(h)); self.P = int(L.gaitk_shared_total(h))
        self.n_streams = int(L.gaitk_num_streams(h)); self.gbuf_floats = int(L.gaitk_gbuf_floats(h))
        self.K = num_classes
        self.in_dims = [int(L.gaitk_stream_in_dim(h, s)) for s in range(self.n_streams)]
        self.in_lens = [int(L.gaitk_stream_in_len(h, s)) for s in range(self.n_streams)]
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().gaitk_plan_destroy(self._h); self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def geometry(self, stream: int, dtype: int = _lib.DTYPE_F32):
        """(CTAs per SM, windows per tile, dynamic shared memory bytes) of one stream kernel."""
        c, w, sm = C.c_int(), C.c_int(), C.c_size_t()
        check(lib().gaitk_stream_geometry(self._h, stream, int(dtype), C.byref(c), C.byref(w), C.byref(sm)))
        return c.value, w.value, sm.value

    def workspace_bytes(self, B: int) -> int:
        return int(lib().gaitk_workspace_bytes(self._h, int(B)))

    def workspace(self, B: int, extra_floats: int = 0) -> torch.Tensor:
        need = self.workspace_bytes(B) + 4 * extra_floats + 512
        if self._ws is None or self._ws.numel() < need:
            # grow-only, and every buffer ever handed out stays alive: captured CUDA graphs keep the ADDRESS of the workspace
            # they were captured with (a smaller batch captured earlier must not find its workspace freed and reused)
            self.__dict__.setdefault("_ws_keep", []).append(self._ws)
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------ raw calls
    def _check_inputs(self, xs: Sequence[Optional[torch.Tensor]], win_start=None):
        B = None
        for s, x in enumerate(xs):
            if x is None:
                continue
            if x.dtype != torch.float32 or not x.is_cuda or not x.is_contiguous():
                raise _lib.GaitkError(f"stream {s}: input must be a contiguous CUDA float32 tensor")
            if win_start is None or win_start[s] is None:
                if x.dim() != 3 or x.shape[1] != self.in_lens[s] or x.shape[2] != self.in_dims[s]:
                    raise _lib.GaitkError(f"stream {s}: expected (B,{self.in_lens[s]},{self.in_dims[s]}), got {tuple(x.shape)}")
                b = x.shape[0]
            else:
                b = win_start[s].numel()
            if B is not None and b != B:
                raise _lib.GaitkError("streams disagree on batch size")
            B = b
        return B

    def forward(self, flat_params: torch.Tensor, xs, enabled_mask: int = 0b111, win_start=None, want=None,
                dtype: int = _lib.DTYPE_F32):
        B = self._check_inputs(xs, win_start)
        logits = [torch.empty(B, self.K, dtype=torch.float32, device=self.device)
                  if (xs[s] is not None and (want is None or want[s])) else None for s in range(self.n_streams)]
        check(lib().gaitk_forward(self._h, flat_params.data_ptr(),
                                  ptr_array([0 if x is None else x.data_ptr() for x in xs]),
                                  None if win_start is None else ptr_array([0 if w is None else w.data_ptr() for w in win_start]),
                                  B, enabled_mask, ptr_array([0 if l is None else l.data_ptr() for l in logits]),
                                  int(dtype), stream_handle()), "gaitk_forward")
        return logits

    def backward(self, flat_params, xs, dlogits, grads_flat: torch.Tensor, enabled_mask: int = 0b111, win_start=None,
                 dtype: int = _lib.DTYPE_F32):
        B = self._check_inputs(xs, win_start)
        ws = self.workspace(B, self.gbuf_floats + 64)
        check(lib().gaitk_backward(self._h, flat_params.data_ptr(),
                                   ptr_array([0 if x is None else x.data_ptr() for x in xs]),
                                   None if win_start is None else ptr_array([0 if w is None else w.data_ptr() for w in win_start]),
                                   B, enabled_mask, ptr_array([0 if d is None else d.data_ptr() for d in dlogits]),
                                   grads_flat.data_ptr(), ws.data_ptr(), ws.numel(), int(dtype), stream_handle()),
              "gaitk_backward")


class FlatParamModule(torch.nn.Module):
    """nn.Module whose parameters are views into ONE flat fp32 buffer laid out in the plan's canonical
    order (== the reference's named_parameters() order), so the kernels take a single pointer and SGD
    is one pass.  state_dict keys/shapes are untouched."""

    _plan: Optional[Plan] = None
    _flat: Optional[torch.Tensor] = None
    compute_dtype: int = _lib.DTYPE_F32      # _lib.DTYPE_TF32 selects the tcgen05 / mma.sync tensor-core kernels

    def _plan_kwargs(self) -> dict:  # pragma: no cover - abstract
        raise NotImplementedError

    def _plan_name_map(self) -> Optional[dict]:
        """plan parameter name -> this module's parameter name; a plan parameter that is missing here (a stream the
        module does not have) is backed by a zero dummy.  None = names are identical."""
        return None

    def _named_for_plan(self):
        named = dict(self.named_parameters())
        mp = self._plan_name_map()
        if mp is None:
            return named
        out = {}
        for pi in self.plan().params:
            src = mp.get(pi.name)
            out[pi.name] = named[src] if src is not None else None
        return out

    def plan(self) -> Plan:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.GaitkError("gaitk modules run on CUDA (sm_100a) only: call .to('cuda') first; there is no CPU path")
        if self._plan is None or self._plan.device != dev:
            object.__setattr__(self, "_plan", Plan(device=dev, **self._plan_kwargs()))
            object.__setattr__(self, "_flat", None)
        return self._plan

    def flat_params(self) -> torch.Tensor:
        """The flat buffer, (re)built whenever some parameter no longer aliases it (after .to(),
        load_state_dict(assign=True), manual .data swaps ...)."""
        plan = self.plan()
        named = self._named_for_plan()
        flat = self._flat
        ok = flat is not None
        if ok:
            base = flat.data_ptr()
            for pi in plan.params:
                p = named[pi.name]
                if p is None:
                    continue
                if p.data_ptr() != base + 4 * pi.offset or p.dtype != torch.float32 or not p.is_contiguous():
                    ok = False; break
        if not ok:
            flat = torch.zeros(plan.NP, dtype=torch.float32, device=plan.device)
            with torch.no_grad():
                for pi in plan.params:
                    p = named[pi.name]
                    if p is None:
                        continue
                    if tuple(p.shape) != pi.shape:
                        raise _lib.GaitkError(f"parameter {pi.name} has shape {tuple(p.shape)}, plan expects {pi.shape}")
                    v = flat[pi.offset:pi.offset + pi.numel].view(pi.shape)
                    v.copy_(p.data)
                    p.data = v
            object.__setattr__(self, "_flat", flat)
        return flat

    def plan_parameters(self) -> List[Optional[torch.nn.Parameter]]:
        named = self._named_for_plan()
        return [named[pi.name] for pi in self.plan().params]


class _StreamsFn(torch.autograd.Function):
    """logits = model(x_0..x_{n-1}); backward runs the fused recompute+backward kernel only for the
    streams that actually received a gradient, so the reference's ``losses[i].backward(retain_graph=True)``
    loop costs one stream pass per loss."""

    @staticmethod
    def forward(ctx, module, enabled_mask, n_streams, *args):
        xs = list(args[:n_streams])
        plan = module.plan()
        flat = module.flat_params()
        dtype = int(getattr(module, 'compute_dtype', _lib.DTYPE_F32))
        logits = plan.forward(flat, xs, enabled_mask, dtype=dtype)
        ctx.dtype = dtype
        ctx.module = module; ctx.enabled_mask = enabled_mask; ctx.n_streams = n_streams
        ctx.xs = xs
        ctx.present = [a is not None for a in args[n_streams:]]
        ctx.set_materialize_grads(False)
        outs = tuple(l if l is not None else torch.zeros(0, device=plan.device) for l in logits)
        ctx.mark_non_differentiable(*[o for o, l in zip(outs, logits) if l is None])
        return outs

    @staticmethod
    def backward(ctx, *dls):
        module = ctx.module; plan = module.plan(); flat = module.flat_params()
        dls = [None if (d is None or ctx.xs[s] is None) else d.contiguous().float() for s, d in enumerate(dls)]
        grads = torch.zeros(plan.NP, dtype=torch.float32, device=plan.device)
        if any(d is not None for d in dls):
            plan.backward(flat, ctx.xs, dls, grads, ctx.enabled_mask, dtype=ctx.dtype)
        touched = set()
        for s, d in enumerate(dls):
            if d is not None:
                touched.add(0); touched.add(1 + s)
        out = [None, None, None] + [None] * ctx.n_streams          # no input gradients (raw sensor data)
        for pi, present in zip(plan.params, ctx.present):
            # parameters no live stream reaches get None, exactly like autograd (enc_i.ln1 never gets a grad)
            reach = pi.group == 0 or pi.group in touched
            out.append(grads[pi.offset:pi.offset + pi.numel].view(pi.shape) if (reach and pi.group >= 0 and present) else None)
        return tuple(out)


def run_streams(module: FlatParamModule, xs: Sequence[Optional[torch.Tensor]], enabled_mask: int):
    xs = [None if x is None else x.contiguous().float() for x in xs]
    params = module.plan_parameters()
    module.flat_params()
    outs = _StreamsFn.apply(module, enabled_mask, len(xs), *xs, *params)
    return [None if xs[s] is None else outs[s] for s in range(len(xs))]
