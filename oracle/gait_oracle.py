"""CPU oracle for the gait hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this module.  The product path
(``gaitk``) never routes through it and fails loudly without the CUDA library.

It is a from-scratch restatement (functional torch-CPU ops + numpy + SciPy, the
same third-party arithmetic the reference dispatches to) of the reference's
algorithm for the path.  Every function cites the reference ``file:line`` it
follows (paths relative to the reference root).  The reference holds no tests
or golden vectors of its own, so the oracle is pinned
against outputs of the reference itself, generated here by
``oracle/make_golden.py`` (which imports the reference read-only) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every one.

Version stamp of the third-party arithmetic: see ``tests/golden/VERSIONS.json``
(SciPy's SLSQP differs between releases; CAGrad weights are pinned to the
installed SciPy only).
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# A1  windowing                                     dataloader_weargait.py:230-237
# --------------------------------------------------------------------------

def window_indices(n_frames: int, win: int, hop: int) -> List[Tuple[int, int, int]]:
    """Strict full windows ``(wid, start, stop)`` while ``start+win <= n``."""
    out: List[Tuple[int, int, int]] = []
    if n_frames <= 0 or n_frames < win:
        return out
    start, wid = 0, 0
    while start + win <= n_frames:
        out.append((wid, start, start + win))
        start += hop
        wid += 1
    return out


# --------------------------------------------------------------------------
# A2  z-score statistics                            dataloader_weargait.py:181-227
# --------------------------------------------------------------------------
MIN_STD = 1e-6


def fit_channel_stats(frames: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per-channel (mean, std, seen) over the *finite* values of a list of
    (N_i, D) float arrays; float64 sum / sum-of-squares accumulators, subject by
    subject, as fit_stats_on_train :183-209 does.  ``seen[d]`` is False for a
    channel with no finite value (the reference then has no stats entry)."""
    D = frames[0].shape[1]
    s = np.zeros(D); ss = np.zeros(D); n = np.zeros(D, dtype=np.int64)
    for x in frames:
        x = np.asarray(x, dtype=float)
        for d in range(D):
            col = x[:, d]
            m = np.isfinite(col)
            if not m.any():
                continue
            v = col[m].astype(float)
            s[d] += float(v.sum()); ss[d] += float(np.dot(v, v)); n[d] += int(v.size)
    seen = n > 0
    mean = np.zeros(D); std = np.full(D, MIN_STD)
    for d in range(D):
        if seen[d]:
            mean[d] = s[d] / n[d]
            var = max(ss[d] / n[d] - mean[d] ** 2, 0.0)
            std[d] = max(math.sqrt(var), MIN_STD)
    return mean, std, seen


def apply_channel_stats(x: np.ndarray, mean: np.ndarray, std: np.ndarray,
                        seen: Optional[np.ndarray] = None) -> np.ndarray:
    """apply_stats :212-227 on a dense (N, D) float64 array: non-finite -> mean,
    ``(x-m)/max(s,MIN_STD)``, nan_to_num -> 0.  Channels without stats pass
    through untouched.  (ensure_cols :76-91 -- a channel that is entirely
    non-finite is mean-filled first -- is covered by the NaN->mean rule.)"""
    x = np.array(x, dtype=float, copy=True)
    D = x.shape[1]
    for d in range(D):
        if seen is not None and not seen[d]:
            continue
        m, s = float(mean[d]), float(std[d])
        col = x[:, d]
        col[~np.isfinite(col)] = m if np.isfinite(m) else 0.0
        s_eff = s if (np.isfinite(s) and s > MIN_STD) else MIN_STD
        z = (col - (m if np.isfinite(m) else 0.0)) / s_eff
        x[:, d] = np.nan_to_num(z, nan=0.0, posinf=0.0, neginf=0.0)
    return x


# --------------------------------------------------------------------------
# A3  index maps                                    dataloader_weargait.py:278-348
# --------------------------------------------------------------------------

def sync_index_map(n_windows: Dict[str, Sequence[int]]) -> List[Tuple[str, int]]:
    """``n_windows[sid] = (n_walkway, n_insole, n_imu)`` window counts.  Returns
    the stream-aligned list ``(sid, wid)`` in reference order: subjects in dict
    order, window ids = intersection over modalities, numeric sort
    (_build_index_maps :287-298; ids are ``0..n-1`` per modality so the
    intersection is ``range(min n)`` and a subject with an empty modality is
    skipped)."""
    out: List[Tuple[str, int]] = []
    for sid, counts in n_windows.items():
        if not all(c > 0 for c in counts):
            continue
        for wid in range(min(counts)):
            out.append((sid, wid))
    return out


def async_key_order(keys: Sequence[str]) -> List[str]:
    """Async datasets address windows through ``sorted(keys)`` -- a *string*
    sort of ``"SID|mod|wid"`` (:318), so wid 10 sorts before wid 2."""
    return sorted(keys)


def async_permutations(lens: Sequence[int], seed: int) -> List[List[int]]:
    """Per-modality permutations, no replacement, truncated to the shortest
    modality (WearGaitMultiAsyncDataset.__init__/reseed :316-334).  One
    ``random.Random(seed)`` is shared by the modalities in order."""
    rng = random.Random(seed)
    mn = min(lens)
    perms = []
    for n in lens:
        idx = list(range(n))
        rng.shuffle(idx)
        perms.append(idx[:mn])
    return perms


# --------------------------------------------------------------------------
# A4  modality masks                                weargait_train.py:49-57,355-358
# --------------------------------------------------------------------------
MASK_COMBOS = {
    "W": (True, False, False), "I": (False, True, False), "M": (False, False, True),
    "W+I": (True, True, False), "W+M": (True, False, True), "I+M": (False, True, True),
    "W+I+M": (True, True, True),
}


def apply_mask(xs: Sequence[torch.Tensor], mask: Sequence[bool]) -> List[torch.Tensor]:
    """Disabled stream -> zeros_like (still encoded downstream)."""
    return [x if m else torch.zeros_like(x) for x, m in zip(xs, mask)]


# --------------------------------------------------------------------------
# A5  FoG / FBG clip preparation                    dataloader_fbg_fog.py:24-121
# --------------------------------------------------------------------------

def pad_or_trim(seq: np.ndarray, target_len: int, pad_value: float = 0.0) -> np.ndarray:
    L = seq.shape[0]
    if L == target_len:
        return seq
    if L > target_len:
        return seq[:target_len]
    pad = np.full((target_len - L, *seq.shape[1:]), pad_value, dtype=seq.dtype)
    return np.concatenate([seq, pad], axis=0)


def center_pose(arr: np.ndarray) -> np.ndarray:
    """Subtract joint 0 from every joint (:93-99).  arr (T, J, 3)."""
    return arr - arr[:, 0:1, :]


def minmax_pose(arr: np.ndarray) -> np.ndarray:
    """Per-clip, per-coordinate min-max over (T, J) (:107-113)."""
    mins = arr.min(axis=(0, 1)); maxs = arr.max(axis=(0, 1))
    return (arr - mins) / (maxs - mins + 1e-6)


def prepare_pose_clip(arr: np.ndarray, target_len: int) -> np.ndarray:
    """center -> minmax -> pad_or_trim -> float32 -> flatten joints
    (create_fusion_loaders :316-318, SkeletonDataset :135,144, flatten_skel
    utilities.py:28-32)."""
    a = pad_or_trim(minmax_pose(center_pose(np.asarray(arr, dtype=float))), target_len)
    return a.astype(np.float32).reshape(target_len, -1)


def prepare_sensor_clip(arr: np.ndarray, target_len: int) -> np.ndarray:
    return pad_or_trim(np.asarray(arr, dtype=float), target_len).astype(np.float32)


# --------------------------------------------------------------------------
# adaptive average pooling bins (ATen adaptive_avg_pool1d)
# --------------------------------------------------------------------------

def adaptive_bins(L: int, O: int) -> List[Tuple[int, int]]:
    """``[floor(i*L/O), ceil((i+1)*L/O))`` -- bins overlap when O does not divide L."""
    return [((i * L) // O, -((-(i + 1) * L) // O)) for i in range(O)]


def adaptive_avg_pool_time(x: torch.Tensor, O: int) -> torch.Tensor:
    """x (B, T, C) -> (B, O, C), averaging over the time bins above."""
    T = x.shape[1]
    return torch.stack([x[:, a:b].mean(1) for a, b in adaptive_bins(T, O)], 1)


# --------------------------------------------------------------------------
# A6-A10  model forward (channels-last functional restatement)
# --------------------------------------------------------------------------
Params = Dict[str, torch.Tensor]


def conv_time(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.Conv1d with zero 'same' padding applied along time of a (B,T,Cin)
    tensor; weight (Cout, Cin, k) as in the reference state_dict."""
    k = w.shape[2]
    return F.conv1d(x.transpose(1, 2), w, b, padding=k // 2).transpose(1, 2)


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def enc_walkway(p: Params, x, pre="enc_w."):
    """WalkwayEncoder weargait_encoders.py:40-52."""
    return layer_norm(gelu(conv_time(x, p[pre + "conv.weight"], p[pre + "conv.bias"])),
                      p[pre + "ln.weight"], p[pre + "ln.bias"])


def enc_imu(p: Params, x, pre="enc_m."):
    """IMUEncoderShallow :54-69 (pool_len=None as build_model passes)."""
    return layer_norm(gelu(conv_time(x, p[pre + "conv.weight"], p[pre + "conv.bias"])),
                      p[pre + "ln.weight"], p[pre + "ln.bias"])


def enc_insole(p: Params, x, pre="enc_i."):
    """InsoleEncoderDeep :71-101.  ``ln1`` exists in the state_dict but is never
    applied (:93-101)."""
    h = gelu(conv_time(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"]))
    y = conv_time(h, p[pre + "conv2.weight"], p[pre + "conv2.bias"])
    if (pre + "skip.weight") in p:
        y = y + conv_time(h, p[pre + "skip.weight"], p[pre + "skip.bias"])
    else:
        y = y + h
    return layer_norm(gelu(y), p[pre + "ln2.weight"], p[pre + "ln2.bias"])


def backbone(p: Params, f, bdim: int, wkey="backbone.conv.weight", bkey="backbone.conv.bias"):
    """SharedBackbone :103-113 / feature_encoder.py:80-109 followed by
    ``.flatten(1)``: feature index ``bin*S + s``."""
    z = torch.relu(conv_time(f, p[wkey], p[bkey]))
    return adaptive_avg_pool_time(z, bdim).flatten(1)


def task_head(p: Params, x, pre: str):
    """TaskHead :30-37 (+CosineLinear :19-28); norm present iff ``pre+'norm.weight'``
    is in the state_dict, cosine iff there is no fc.bias."""
    if (pre + "norm.weight") in p:
        x = layer_norm(x, p[pre + "norm.weight"], p[pre + "norm.bias"])
    w = p[pre + "fc.weight"]
    if (pre + "fc.bias") in p:
        return x @ w.t() + p[pre + "fc.bias"]
    eps = 1e-8
    xn = x / x.norm(dim=1, keepdim=True).clamp_min(eps)
    wn = w / w.norm(dim=1, keepdim=True).clamp_min(eps)
    return (xn @ wn.t()).clamp(-1.0 + eps, 1.0 - eps)


def weargait_forward(p: Params, xw, xi, xm, bdim: int = 8):
    """WearGaitThreeModal.forward :148-156 -> (lw, li, lm)."""
    fw = backbone(p, enc_walkway(p, xw), bdim)
    fi = backbone(p, enc_insole(p, xi), bdim)
    fm = backbone(p, enc_imu(p, xm), bdim)
    return task_head(p, fw, "head_w."), task_head(p, fi, "head_i."), task_head(p, fm, "head_m.")


def fog_forward(p: Params, x_skel, x_sens, sensor_length: int, bdim: int = 8,
                synchronized: bool = False, out_len: int = 101):
    """MultiModalMultiTaskModel.forward feature_encoder.py:222-254."""
    # SkeletonMLP :73-77
    h = x_skel @ p["skeleton_encoder.fc1.weight"].t() + p["skeleton_encoder.fc1.bias"]
    sk = torch.relu(layer_norm(h, p["skeleton_encoder.ln1.weight"], p["skeleton_encoder.ln1.bias"]))
    # SensorEncoder :51-58 (pool only when T == sensor_length)
    se = conv_time(x_sens, p["sensor_encoder.conv1d.weight"], p["sensor_encoder.conv1d.bias"])
    if se.shape[1] == sensor_length:
        se = adaptive_avg_pool_time(se, out_len)
    rs = backbone(p, sk, bdim, "backbone.conv1d.weight", "backbone.conv1d.bias")
    rt = backbone(p, se, bdim, "backbone.conv1d.weight", "backbone.conv1d.bias")
    if synchronized:
        return task_head(p, rs, "task_head_shared."), task_head(p, rt, "task_head_shared.")
    return task_head(p, rs, "task_head_skel."), task_head(p, rt, "task_head_sensor.")


def fog_single_forward(p: Params, x, kind: str, sensor_length: int = 426, bdim: int = 8, out_len: int = 101):
    """SkelModalityModel / SensorModalityModel.forward feature_encoder.py:301-305,340-344
    (state_dict names encoder.*, backbone.*, task_head.*)."""
    if kind == "skeleton":
        h = x @ p["encoder.fc1.weight"].t() + p["encoder.fc1.bias"]
        f = torch.relu(layer_norm(h, p["encoder.ln1.weight"], p["encoder.ln1.bias"]))
    else:
        f = conv_time(x, p["encoder.conv1d.weight"], p["encoder.conv1d.bias"])
        if f.shape[1] == sensor_length:
            f = adaptive_avg_pool_time(f, out_len)
    r = backbone(p, f, bdim, "backbone.conv1d.weight", "backbone.conv1d.bias")
    return task_head(p, r, "task_head.")


def weargait_single_forward(p: Params, x, mod: str, bdim: int = 8):
    """_single_logits_and_labels weargait_train.py:252-271: one branch of the 3-stream model."""
    enc = {"walkway": enc_walkway, "insole": enc_insole, "imu": enc_imu}[mod]
    head = {"walkway": "head_w.", "insole": "head_i.", "imu": "head_m."}[mod]
    return task_head(p, backbone(p, enc(p, x), bdim), head)


# --------------------------------------------------------------------------
# A10  losses                          classification_losses.py:54-109 etc.
# --------------------------------------------------------------------------

def inv_freq_weights(counts: Sequence[int]) -> torch.Tensor:
    """weargait_train.py:107-109 / utilities.py:130-132."""
    w = 1.0 / (torch.tensor(counts, dtype=torch.float32) + 1e-8)
    return w / w.sum() * len(counts)


def weighted_ce(logits, y, weight=None):
    """F.cross_entropy(reduction='mean'): sum(w_y * nll) / sum(w_y)."""
    logp = torch.log_softmax(logits, dim=1)
    nll = -logp.gather(1, y.view(-1, 1)).squeeze(1)
    if weight is None:
        return nll.mean()
    wy = weight[y]
    return (wy * nll).sum() / wy.sum()


def gcl_loss(logits, y, cls_num_list, m=0.5, s=30.0, weight=None, noise_mul=1.0,
             noise: Optional[torch.Tensor] = None):
    """GCLLoss.forward :97-109 (train_cls=False).  ``noise`` is the clamped
    N(0, 1/3) sample the reference draws from the CPU RNG every call; pass None
    with noise_mul == 0.  NaN when all class counts are equal (0/0 at :104)."""
    cl = torch.tensor(cls_num_list, dtype=torch.float32)
    m_list = torch.log(cl); m_list = m_list.max() - m_list
    if noise is None:
        noise = torch.zeros_like(logits)
    z = logits - noise_mul * noise.abs() / m_list.max() * m_list
    onehot = F.one_hot(y, logits.shape[1]).bool()
    z = torch.where(onehot, z - m, z)
    return weighted_ce(s * z, y, weight)


def ldam_loss(logits, y, cls_num_list, max_m=0.5, s=30.0, weight=None):
    """LDAMLoss :54-76."""
    ml = 1.0 / np.sqrt(np.sqrt(np.asarray(cls_num_list, dtype=float)))
    ml = torch.tensor(ml * (max_m / ml.max()), dtype=torch.float32)
    onehot = F.one_hot(y, logits.shape[1]).bool()
    z = torch.where(onehot, logits - ml[y].view(-1, 1), logits)
    return weighted_ce(s * z, y, weight)


def sym_kl(la, lb):
    """fbg_fog_train.py:81-89 (kl_div batchmean both ways)."""
    lpa, lpb = torch.log_softmax(la, 1), torch.log_softmax(lb, 1)
    pa, pb = lpa.exp(), lpb.exp()
    B = la.shape[0]
    return (pb * (lpb - lpa)).sum() / B + (pa * (lpa - lpb)).sum() / B


# --------------------------------------------------------------------------
# A11  CAGrad                                   multitask_weighting.py:653-776
# --------------------------------------------------------------------------

def cagrad_combine(G: torch.Tensor, alpha: float):
    """``cagrad`` :694-729 with rescale=1.  G (P, n) float32.  Returns
    (g (P,), GG float32 ndarray (n,n), w float64 (n,))."""
    from scipy.optimize import minimize
    n = G.shape[1]
    GG = G.t().mm(G)
    g0 = (GG.mean() + 1e-8).sqrt()
    A = GG.numpy()
    b = np.ones(n) / n
    c = (alpha * g0 + 1e-8).item()

    def obj(x):
        return float(x @ A @ b + c * np.sqrt(x @ A @ x + 1e-8))

    res = minimize(obj, b.copy(), bounds=tuple((0, 1) for _ in range(n)),
                   constraints={"type": "eq", "fun": lambda x: 1 - sum(x)})
    w = res.x
    ww = torch.tensor(w, dtype=torch.float32)
    gw = (G * ww.view(1, -1)).sum(1)
    lam = c / (gw.norm() + 1e-8)
    g = (G.mean(1) + lam * gw) / (1 + alpha ** 2)
    return g, A, w


def cagrad_objective(A: np.ndarray, w: np.ndarray, alpha: float) -> float:
    """Objective SLSQP minimises, for comparing solvers by value."""
    A = np.asarray(A, dtype=float); n = A.shape[0]
    c = alpha * math.sqrt(A.mean() + 1e-8) + 1e-8
    return float(w @ A @ (np.ones(n) / n) + c * math.sqrt(w @ A @ w + 1e-8))


def clip_coef(total_norm: float, max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_: min(1, max_norm / (norm + 1e-6))."""
    return min(1.0, max_norm / (total_norm + 1e-6))


def flat_grads(loss, params: Sequence[torch.Tensor]) -> torch.Tensor:
    gs = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
    return torch.cat([(torch.zeros_like(p) if g is None else g).reshape(-1)
                      for p, g in zip(params, gs)])


def step_gradients(p: Params, losses: Sequence[torch.Tensor], shared_keys: Sequence[str],
                   private_keys: Sequence[Sequence[str]], alpha: float, max_norm: float,
                   private_twice: bool):
    """Gradients left in ``.grad`` by one reference step.

    * shared (CAGrad.backward :761-776): G column i = dL_i/d(shared);
      grad = n * cagrad(G), then clip_grad_norm_(shared, max_norm).
    * private: CAGrad's n full-graph backwards accumulate sum_i dL_i/dtheta into
      every non-shared leaf (:680-688).  weargait_train.step_cagrad_three
      :218-242 then adds dL_k/dtheta_k again (``private_twice``); fbg_fog_train
      :146-152 does not.
    Returns dict name -> grad tensor (None for parameters no loss reaches),
    plus extras (G, GG, w, pre-clip norm).
    """
    n = len(losses)
    keys = list(p.keys())
    leaves = [p[k] for k in keys]
    # n full-graph sweeps (losses[i].backward(retain_graph=True) :680-688): shared columns AND the
    # accumulation into every other leaf come out of the same sweep
    sweeps = [torch.autograd.grad(L, leaves, retain_graph=True, allow_unused=True) for L in losses]
    sh = set(shared_keys)
    idx = {k: i for i, k in enumerate(keys)}
    G = torch.stack([torch.cat([(torch.zeros_like(p[k]) if sw[idx[k]] is None else sw[idx[k]]).reshape(-1)
                                for k in shared_keys]) for sw in sweeps], 1)
    g, GG, w = cagrad_combine(G, alpha)
    g = g * n
    norm = float(g.norm())
    if max_norm > 0:
        g = g * clip_coef(norm, max_norm)
    grads: Dict[str, Optional[torch.Tensor]] = {}
    off = 0
    for k in shared_keys:
        cnt = p[k].numel(); grads[k] = g[off:off + cnt].view_as(p[k]).clone(); off += cnt
    for k in keys:
        if k in sh:
            continue
        acc = None
        for sw in sweeps:
            gk = sw[idx[k]]
            if gk is not None:
                acc = gk.clone() if acc is None else acc + gk
        grads[k] = acc
    if private_twice:
        # torch.autograd.grad(L_k, private_k) once more per stream (weargait_train.py:218-242)
        for L, pk in zip(losses, private_keys):
            if not pk:
                continue
            gs = torch.autograd.grad(L, [p[k] for k in pk], retain_graph=True, allow_unused=True)
            for k, gk in zip(pk, gs):
                if gk is not None:
                    grads[k] = gk.clone() if grads[k] is None else grads[k] + gk
    return grads, {"G": G, "GTG": GG, "weights": w, "norm": norm}


# --------------------------------------------------------------------------
# A12  SGD                                       torch.optim.SGD (momentum, wd)
# --------------------------------------------------------------------------

def sgd_update(p: Params, grads, bufs: Dict[str, torch.Tensor], lr=1e-3, momentum=0.9, wd=1e-4):
    """In-place; parameters whose grad is None are skipped entirely (no weight
    decay either) -- this is why ``enc_i.ln1`` never moves."""
    with torch.no_grad():
        for k, v in p.items():
            g = grads.get(k)
            if g is None:
                continue
            g = g + wd * v
            if k not in bufs:
                bufs[k] = g.clone()
            else:
                bufs[k].mul_(momentum).add_(g)
            v.sub_(lr * bufs[k])


# --------------------------------------------------------------------------
# whole training steps
# --------------------------------------------------------------------------
WG_PRIVATE_PREFIX = ("enc_w.", "enc_i.", "enc_m.")
WG_HEADS = ("head_w.", "head_i.", "head_m.")


def canonical_params(state: Dict[str, np.ndarray], synchronized: bool) -> Params:
    """Reference state_dict -> leaf tensors, de-aliasing the sync head: in sync
    mode head_w/head_i/head_m/_shared_head are ONE module
    (weargait_encoders.py:133-136), so the oracle keeps ``head_w.*`` only and
    maps the other names onto it."""
    p: Params = {}
    for k, v in state.items():
        if k.startswith("_shared_head."):
            continue
        if synchronized and (k.startswith("head_i.") or k.startswith("head_m.")):
            continue
        p[k] = torch.tensor(np.asarray(v), dtype=torch.float32).requires_grad_(True)
    return p


class _HeadAlias(dict):
    """Param view where head_i./head_m. resolve to head_w. (sync mode)."""
    def __init__(self, base): super().__init__(base)
    def _k(self, k):
        for h in ("head_i.", "head_m."):
            if k.startswith(h):
                return "head_w." + k[len(h):]
        return k
    def __getitem__(self, k): return dict.__getitem__(self, self._k(k))
    def __contains__(self, k): return dict.__contains__(self, self._k(k))


def weargait_losses(p: Params, xs, ys, *, synchronized, wm, counts, gcl_m=0.2, gcl_s=25.0,
                    noise_mul=0.0, weights=None, bdim=8, noises=None):
    """forward_batch :163-184 + criteria :111-130 -> (logits, losses)."""
    view = _HeadAlias(p) if synchronized else p
    logits = weargait_forward(view, *xs, bdim=bdim)
    losses = []
    for i, (lg, y) in enumerate(zip(logits, ys)):
        wt = None if weights is None else weights[i]
        if wm == "gcl":
            losses.append(gcl_loss(lg, y, counts[i], m=gcl_m, s=gcl_s, weight=wt, noise_mul=noise_mul,
                                   noise=None if noises is None else noises[i]))
        else:
            losses.append(weighted_ce(lg, y, wt))
    return logits, losses


def weargait_shared_keys(p: Params, synchronized: bool) -> List[str]:
    """get_shared_parameters :185-189: backbone params, then the shared head's
    (norm before fc, module registration order)."""
    keys = ["backbone.conv.weight", "backbone.conv.bias"]
    if synchronized:
        keys += [k for k in ("head_w.norm.weight", "head_w.norm.bias", "head_w.fc.weight", "head_w.fc.bias")
                 if k in p]
    return keys


def weargait_private_keys(p: Params, synchronized: bool) -> List[List[str]]:
    out = []
    for enc, head in zip(WG_PRIVATE_PREFIX, WG_HEADS):
        ks = [k for k in p if k.startswith(enc)]
        if not synchronized:
            ks += [k for k in p if k.startswith(head)]
        out.append(ks)
    return out


def weargait_train_step(p: Params, bufs, xs, ys, *, synchronized=True, wm="gcl", counts=None,
                        alpha=0.5, max_norm=1.0, lr=1e-3, momentum=0.9, wd=1e-4, tasks=None, **loss_kw):
    """One train_one_epoch iteration :305-311: forward, 3 losses, step_cagrad_three, SGD.  Mutates p/bufs;
    returns diagnostics.

    ``tasks`` (3 bools) is the relaxed-input training case of SURVEY 8(d) cfg 3(ii): a disabled stream is
    zero-filled exactly as _maybe_zero (weargait_train.py:355-358) and its loss is passed as None, which
    step_cagrad_three filters out (:200-203); CAGrad then runs with n_tasks = number of live losses (the
    harness holds one CAGrad(n_tasks=k) per k, since get_weighted_loss indexes range(self.n_tasks))."""
    if tasks is None:
        tasks = (True, True, True)
    xs = apply_mask(xs, tasks)
    logits, losses = weargait_losses(p, xs, ys, synchronized=synchronized, wm=wm, counts=counts, **loss_kw)
    live = [i for i, t in enumerate(tasks) if t]
    priv = weargait_private_keys(p, synchronized)
    dead_keys = set(k for i in range(3) if i not in live for k in priv[i])
    sub = {k: v for k, v in p.items() if k not in dead_keys}          # leaves no live loss reaches keep grad None
    grads, extra = step_gradients(sub, [losses[i] for i in live], weargait_shared_keys(p, synchronized),
                                  [priv[i] for i in live], alpha, max_norm, private_twice=True)
    sgd_update(p, grads, bufs, lr, momentum, wd)
    extra.update(logits=[l.detach() for l in logits], losses=[float(l.detach()) for l in losses], grads=grads, live=live)
    return extra


def fog_losses(p: Params, x_skel, x_sens, ys, yt, *, sensor_length, synchronized, wm, counts,
               gcl_m=0.2, gcl_s=25.0, ldam_m=0.5, ldam_s=30.0, consistency_lambda=1.0,
               weights=None, bdim=8):
    """process_batch :66-144 (multimodal)."""
    ls, lt = fog_forward(p, x_skel, x_sens, sensor_length, bdim, synchronized)
    w_s, w_t = (None, None) if weights is None else weights
    if wm == "gcl":
        l1 = gcl_loss(ls, ys, counts[0], m=gcl_m, s=gcl_s, weight=w_s, noise_mul=0.0)
        l2 = gcl_loss(lt, yt, counts[1], m=gcl_m, s=gcl_s, weight=w_t, noise_mul=0.0)
        if synchronized:
            cons = sym_kl(ls, lt)
            l1 = l1 + 0.5 * consistency_lambda * cons
            l2 = l2 + 0.5 * consistency_lambda * cons
    elif wm == "ldam":
        l1 = ldam_loss(ls, ys, counts[0], max_m=ldam_m, s=ldam_s, weight=inv_freq_weights(counts[0]))
        l2 = ldam_loss(lt, yt, counts[1], max_m=ldam_m, s=ldam_s, weight=inv_freq_weights(counts[1]))
    elif wm == "class_wt":
        l1 = weighted_ce(ls, ys, inv_freq_weights(counts[0]))
        l2 = weighted_ce(lt, yt, inv_freq_weights(counts[1]))
    else:
        l1 = weighted_ce(ls, ys); l2 = weighted_ce(lt, yt)
    return (ls, lt), [l1, l2]


def fog_shared_keys(p: Params, synchronized: bool) -> List[str]:
    """feature_encoder.py:256-265."""
    keys = ["backbone.conv1d.weight", "backbone.conv1d.bias"]
    if synchronized:
        keys += [k for k in ("task_head_shared.norm.weight", "task_head_shared.norm.bias",
                             "task_head_shared.fc.weight", "task_head_shared.fc.bias") if k in p]
    return keys


def fog_train_step(p: Params, bufs, x_skel, x_sens, ys, yt, *, sensor_length, synchronized=False,
                   wm="gcl", counts=None, alpha=0.1, max_norm=1.0, lr=1e-3, momentum=0.9, wd=1e-4,
                   **loss_kw):
    """process_batch(train=True) :146-152 with CAGrad(n_tasks=2)."""
    logits, losses = fog_losses(p, x_skel, x_sens, ys, yt, sensor_length=sensor_length,
                                synchronized=synchronized, wm=wm, counts=counts, **loss_kw)
    grads, extra = step_gradients(p, losses, fog_shared_keys(p, synchronized), [[], []], alpha,
                                  max_norm, private_twice=False)
    sgd_update(p, grads, bufs, lr, momentum, wd)
    extra.update(logits=[l.detach() for l in logits], losses=[float(l.detach()) for l in losses], grads=grads)
    return extra


# --------------------------------------------------------------------------
# evaluation under masks                         weargait_train.py:391-433
# --------------------------------------------------------------------------

def cheap_xattn(a, b):
    """CheapCrossAttention weargait_encoders.py:324-337: softmax(A B^T / sqrt(d)) B, no parameters."""
    sim = (a @ b.transpose(1, 2)) * (a.shape[-1] ** -0.5)
    return torch.softmax(sim, dim=-1) @ b


def baseline_forward(p: Params, xs, kind: str, synchronized: bool, bdim: int = 8):
    """Fusion baselines weargait_encoders.py:209-387 (``kind`` = "early_fusion" | "late_fusion" | "shared_latent" |
    "cheap_xattn").
    early_fusion (:209-245): ONE backbone over the channel concat of the three encoder outputs; sync: one shared head,
        async: three heads on the SAME fused representation.
    late_fusion  (:247-282): sync: shared head on the MEAN latent; async: per-stream heads on per-stream latents.
    shared_latent(:284-322): encoder -> Linear(enc_out_ch->proj_ch) per stream -> shared backbone -> head(s).
    cheap_xattn  (:339-387): six pairwise zero-parameter cross-attentions, X* = mean of X attended to the other two,
        shared backbone, per-branch heads (shared module in sync)."""
    view = _HeadAlias(p) if synchronized else p
    feats = [enc_walkway(p, xs[0]), enc_insole(p, xs[1]), enc_imu(p, xs[2])]
    if kind == "early_fusion":
        rep = backbone(p, torch.cat(feats, dim=-1), bdim)
        if synchronized:
            lg = task_head(view, rep, "head_w.")
            return lg, lg, lg
        return tuple(task_head(view, rep, h) for h in WG_HEADS)
    if kind == "cheap_xattn":
        W, I, M = feats
        feats = [(cheap_xattn(W, I) + cheap_xattn(W, M)) * 0.5, (cheap_xattn(I, W) + cheap_xattn(I, M)) * 0.5,
                 (cheap_xattn(M, W) + cheap_xattn(M, I)) * 0.5]
    elif kind == "shared_latent":
        feats = [f @ p[f"proj_{m}.weight"].t() + p[f"proj_{m}.bias"] for f, m in zip(feats, "wim")]
    elif kind != "late_fusion":
        raise ValueError(kind)
    reps = [backbone(p, f, bdim) for f in feats]
    if kind == "late_fusion" and synchronized:
        lg = task_head(view, (reps[0] + reps[1] + reps[2]) / 3.0, "head_w.")
        return lg, lg, lg
    return tuple(task_head(view, r, h) for r, h in zip(reps, WG_HEADS))


def baseline_train_step(p: Params, bufs, xs, ys, *, kind, synchronized, lr=1e-3, momentum=0.9, wd=1e-4, bdim=8):
    """step_cagrad_three with cagrad=None (weargait_train.py:244-248): mean of the three CE losses, one backward,
    no clipping, SGD."""
    logits = baseline_forward(p, xs, kind, synchronized, bdim)
    losses = [weighted_ce(lg, y) for lg, y in zip(logits, ys)]
    keys = list(p)
    gs = torch.autograd.grad(torch.stack(losses).mean(), [p[k] for k in keys], allow_unused=True)
    grads = {k: g for k, g in zip(keys, gs)}
    sgd_update(p, grads, bufs, lr, momentum, wd)
    return dict(logits=[l.detach() for l in logits], losses=[float(l.detach()) for l in losses], grads=grads)


def fog_baseline_forward(p: Params, x_skel, x_sens, kind: str, *, sensor_length: int, synchronized: bool, bdim: int = 8,
                         out_len: int = 101):
    """2-stream fusion baselines feature_encoder.py:346-596 behind baselines/fusion_train.py (``kind`` = "early" | "late" |
    "share_latent" | "cheap_xattn").  Returns one logits tensor (sync, except share_latent) or (logits_skel, logits_sens).
    early  (:347-394): backbone over the channel concat of the two encoder outputs; head / head_skel + head_sens
    late   (:397-444): shared backbone per stream, head(s) on the CONCAT of the two latents (2 * feature_dim inputs)
    share_latent (:447-491): Linear(enc -> S) per stream, shared backbone, ONE head applied to each latent
    cheap_xattn  (:494-596): symmetric zero-parameter cross-attention, fused = mean of the two attended sequences"""
    h = x_skel @ p["skel_enc.fc1.weight"].t() + p["skel_enc.fc1.bias"]
    sk = torch.relu(layer_norm(h, p["skel_enc.ln1.weight"], p["skel_enc.ln1.bias"]))
    se = conv_time(x_sens, p["sens_enc.conv1d.weight"], p["sens_enc.conv1d.bias"])
    if se.shape[1] == sensor_length:
        se = adaptive_avg_pool_time(se, out_len)
    bb = lambda f: backbone(p, f, bdim, "backbone.conv1d.weight", "backbone.conv1d.bias")
    lin = lambda x, pre: x @ p[pre + "weight"].t() + p[pre + "bias"]
    if kind == "share_latent":
        rs = bb(lin(sk, "proj_skel.")); rt = bb(lin(se, "proj_sens."))
        return lin(rs, "head."), lin(rt, "head.")
    if kind == "early":
        rep = bb(torch.cat([sk, se], dim=-1))
    elif kind == "late":
        rep = torch.cat([bb(sk), bb(se)], dim=1)
    elif kind == "cheap_xattn":
        sim = (sk @ se.transpose(1, 2)) * (sk.shape[-1] ** -0.5)
        s_star = torch.softmax(sim, dim=-1) @ se
        g_star = torch.softmax(sim.transpose(1, 2), dim=-1) @ sk
        rep = bb((s_star + g_star) * 0.5)
    else:
        raise ValueError(kind)
    if synchronized:
        return lin(rep, "head.")
    return lin(rep, "head_skel."), lin(rep, "head_sens.")


def fog_baseline_loss(out, ys, yt, kind: str, synchronized: bool):
    """fusion_train.py:234-242: CE on the skeleton label (single-output models), else the mean of the two CE losses."""
    if synchronized and kind != "share_latent":
        return weighted_ce(out, ys)
    return 0.5 * (weighted_ce(out[0], ys) + weighted_ce(out[1], yt))


def eval_mask_sync(p: Params, xs, y, mask, bdim=8) -> Tuple[int, int]:
    """(#correct of the softmax-mean ensemble over enabled streams, B)."""
    with torch.no_grad():
        logits = weargait_forward(_HeadAlias(p), *apply_mask(xs, mask), bdim=bdim)
        probs = [torch.softmax(l, 1) for l, m in zip(logits, mask) if m]
        pr = sum(probs) / len(probs)
        return int((pr.argmax(1) == y).sum()), int(y.numel())


# --------------------------------------------------------------------------
# synthetic data of WearGait / FoG shape
# --------------------------------------------------------------------------

def synth_weargait_batch(B: int, T: int = 64, seed: int = 0, p_pd: float = 0.6):
    rng = np.random.default_rng(seed)
    y = (rng.random(B) < p_pd).astype(np.int64)
    if B > 1:
        y[0], y[1] = 0, 1
    xw = rng.random((B, T, 2), dtype=np.float32)
    xi = rng.standard_normal((B, T, 13), dtype=np.float32)
    sig = np.where(y == 1, 2.0, 1.0).astype(np.float32)[:, None, None]
    xm = rng.standard_normal((B, T, 24), dtype=np.float32) * sig
    return [xw, xi, xm], y


def synth_weargait_labels(B: int, seed: int = 0, p_pd: float = 0.6):
    """The label vector synth_weargait_batch(B, seed=seed) returns, without generating the sensor data."""
    rng = np.random.default_rng(seed)
    y = (rng.random(B) < p_pd).astype(np.int64)
    if B > 1:
        y[0], y[1] = 0, 1
    return y


def synth_fog_batch(B: int, seed: int = 0, pose_len=101, sens_len=426, joints=7, sens_ch=6):
    rng = np.random.default_rng(seed)
    y = rng.choice(3, size=B, p=[0.5, 0.3, 0.2]).astype(np.int64)
    sk = rng.random((B, pose_len, joints * 3), dtype=np.float32)
    se = rng.standard_normal((B, sens_len, sens_ch), dtype=np.float32)
    Ls = rng.integers(min(40, pose_len // 3), pose_len + 1, size=B)
    Lt = rng.integers(min(140, sens_len // 3), sens_len + 1, size=B)
    for b in range(B):
        sk[b, Ls[b]:] = 0.0; se[b, Lt[b]:] = 0.0
    return sk, se, y
