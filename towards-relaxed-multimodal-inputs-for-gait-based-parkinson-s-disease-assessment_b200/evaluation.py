"""Relaxed-input evaluation: the mask table and the eval passes of ``train/weargait_train.py`` on the device.

``eval_with_mask`` (:391-433) only reads the logits of the ENABLED streams, and in the three-stream model a stream's
logits depend on its own input alone (weargait_encoders.py:148-156), so the seven zero-filled forward passes the
reference runs per fold (``eval_all_masks`` :384-389) collapse into ONE forward pass plus ``gaitk_mask_eval``, which
counts the hits of all seven softmax-mean ensembles and the per-stream hits in one launch per batch.  Counts stay on the
device for the whole pass (the reference synchronises with ``.item()`` several times per batch); one read at the end.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .dataloader_weargait import DeviceLoader, IndexBatch

# weargait_train.py:49-57 (order matters: it is the column order of the summary table)
MASK_COMBOS = {
    "W": (True, False, False), "I": (False, True, False), "M": (False, False, True),
    "W+I": (True, True, False), "W+M": (True, False, True), "I+M": (False, True, True), "W+I+M": (True, True, True),
}
_STREAMS = ("walkway", "insole", "imu")


def _ptr_array(ts: Sequence[torch.Tensor]):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def streams_are_independent(model) -> bool:
    """True when stream s's logits depend on input s alone (WearGaitThreeModal, SharedLatent3, LateFusion3 async): only
    then do the seven zero-filled passes of the reference collapse into one.  Fusion models whose logits couple the
    streams (LateFusion3 sync feeds the MEAN latent -- zero-input latents included -- through the shared head,
    weargait_encoders.py:272-279; EarlyFusion3; CheapXAttn3) declare `fuses_streams = True` and are evaluated with one
    masked pass per mask, exactly as eval_with_mask does."""
    return not bool(getattr(model, "fuses_streams", False))


def _indexed_forward(model, ib: IndexBatch):
    return model.plan().forward(model.flat_params(), list(ib.frames), win_start=list(ib.win_start))


def _dense(ib: IndexBatch, win: int):
    """materialise the (B, T, D) windows of an index batch (fused models go through model.forward)"""
    out = []
    for fr, ws in zip(ib.frames, ib.win_start):
        idx = ws.view(-1, 1) + torch.arange(win, device=ws.device).view(1, -1)
        out.append(fr[idx].contiguous())
    return out


def _batches(model, loader, async_mode: bool, mask=None):
    """yield (logits[3], ys[3]) per batch.  mask = None: all streams enabled (one-pass evaluation of independent
    streams); mask = 3 booleans: disabled inputs are zero-filled and the batch goes through model.forward
    (forward_batch_masked :360-382).  DeviceLoader batches are read from the resident stores by index."""
    independent = streams_are_independent(model)

    def run(xs):
        if mask is not None:
            xs = [x if u else torch.zeros_like(x) for x, u in zip(xs, mask)]
        return model(*xs)

    if isinstance(loader, DeviceLoader):
        ds = loader.dataset
        st0 = ds.stores[0] if isinstance(ds.stores, tuple) else ds.stores[ds.modalities[0]]
        if hasattr(model, "set_window"):
            model.set_window(st0.win)
        for ib in loader.index_batches():
            if independent and mask is None:
                yield _indexed_forward(model, ib), ib.ys
            else:
                yield run(_dense(ib, st0.win)), ib.ys
        return
    for b in loader:
        if isinstance(b, IndexBatch):
            if independent and mask is None:
                yield _indexed_forward(model, b), b.ys
            else:
                win = int(getattr(model, "_T", 0) or 0)
                if not win:
                    raise _lib.GaitkError("window length unknown for an index batch: call model.set_window(T) first")
                yield run(_dense(b, win)), b.ys
        elif async_mode:                                    # forward_batch :163-184
            xs = [b[m].cuda().float() for m in _STREAMS]; ys = [b["y"][m].cuda().long() for m in _STREAMS]
            yield run(xs), ys
        else:
            xs = [t.cuda().float() for t in b["xs"]]; y = b["y"].cuda().long()
            yield run(xs), [y, y, y]


@torch.no_grad()
def mask_counts(model, loader, async_mode: bool, criterions: Optional[Sequence] = None, mask=None):
    """One pass over ``loader``: int32 (n_batches, 10) hit counts (see gaitk_mask_eval), batch sizes, and -- when
    criterions are given -- the (n_batches, 3) per-stream losses.  Single device->host read at the end.
    mask: see _batches (only the counter of that mask is meaningful then)."""
    was_training = model.training
    model.eval()
    L = _lib.lib(); rows, sizes, losses = [], [], []
    for lg, ys in _batches(model, loader, async_mode, mask):
        lg = [l.contiguous() for l in lg]; ys = [y.contiguous() for y in ys]
        B, K = lg[0].shape
        cnt = torch.zeros(10, dtype=torch.int32, device=lg[0].device)
        _lib.check(L.gaitk_mask_eval(_ptr_array(lg), _ptr_array(ys), B, K, cnt.data_ptr(), _lib.stream_handle()), "gaitk_mask_eval")
        rows.append(cnt); sizes.append(B)
        if criterions is not None:
            losses.append(torch.stack([c(l, y).detach().reshape(()) for c, l, y in zip(criterions, lg, ys)]))
    model.train(was_training)
    counts = torch.stack(rows).cpu().numpy() if rows else np.zeros((0, 10), dtype=np.int32)
    loss = torch.stack(losses).cpu().numpy() if losses else None
    return counts, np.asarray(sizes, dtype=np.int64), loss


def _batch_acc(c: np.ndarray, n: np.ndarray) -> np.ndarray:
    """(pred == y).float().mean().item() * 100 per batch: fp32 mean, then double"""
    return (c.astype(np.float32) / n.astype(np.float32)).astype(np.float64) * 100.0


def _mask_result(counts, sizes, async_mode, idx, mask):
    use = [bool(u) for u in mask]
    if not async_mode:
        return 100.0 * float(counts[:, idx].sum()) / max(1, int(sizes.sum()))
    k = max(1, len(sizes)); res = {}
    for s, nm in enumerate(_STREAMS):
        if use[s]:
            res[nm] = float(_batch_acc(counts[:, 7 + s], sizes).sum()) / k if len(sizes) else 0.0
    res["macro_enabled"] = sum(res.values()) / max(1, len(res)) if res else 0.0
    return res


def eval_all_masks(model, loader, async_mode: bool) -> Dict[str, object]:
    """:384-389 -- the whole seven-mask table from one pass (sync: accuracy in %, async: dict per enabled stream +
    ``macro_enabled``), same values as calling the reference's eval_with_mask per mask."""
    if streams_are_independent(model):
        counts, sizes, _ = mask_counts(model, loader, async_mode)
        return {k: _mask_result(counts, sizes, async_mode, i, m) for i, (k, m) in enumerate(MASK_COMBOS.items())}
    out = {}
    for i, (k, m) in enumerate(MASK_COMBOS.items()):         # coupled streams: one zero-filled pass per mask (:384-389)
        counts, sizes, _ = mask_counts(model, loader, async_mode, mask=m)
        out[k] = _mask_result(counts, sizes, async_mode, i, m)
    return out


def eval_with_mask(model, loader, async_mode: bool, mask, verbose: bool = False):
    """:391-433 for one mask (name or 3 booleans)."""
    if isinstance(mask, str):
        mask = MASK_COMBOS[mask]
    mask = tuple(bool(u) for u in mask)
    names = [k for k, v in MASK_COMBOS.items() if v == mask]
    if not names:                                            # (False, False, False): nothing enabled
        return 0.0 if not async_mode else {"macro_enabled": 0.0}
    counts, sizes, _ = mask_counts(model, loader, async_mode, mask=None if streams_are_independent(model) else mask)
    return _mask_result(counts, sizes, async_mode, list(MASK_COMBOS).index(names[0]), mask)


def eval_one_epoch(model, loader, async_mode: bool, criterions):
    """:322-350: per-stream mean batch loss, per-stream mean batch accuracy, micro ensemble accuracy (sync only)."""
    counts, sizes, loss = mask_counts(model, loader, async_mode, criterions)
    n = max(1, len(sizes))
    per_mod_loss = (loss.astype(np.float64).sum(0) / n).tolist() if loss is not None and len(loss) else [0.0, 0.0, 0.0]
    per_mod_acc = [float(_batch_acc(counts[:, 7 + s], sizes).sum()) / n if len(sizes) else 0.0 for s in range(3)]
    ens = None if async_mode else 100.0 * float(counts[:, 6].sum()) / max(1, int(sizes.sum()))
    return per_mod_loss, per_mod_acc, ens
