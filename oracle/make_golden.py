"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (CPU, this container).

Test infrastructure only.  Imports the reference read-only from /root/reference
(never copies its sources), applies the two documented compatibility shims
(S1 pandas>=3 copy-on-write in apply_stats, S2
torch.cuda.FloatTensor on a CPU-only box), drives the reference's own functions
(`WearGaitThreeModal`, `GCLLoss`, `CAGrad`, `step_cagrad_three`,
`process_batch`, `prepare_split`, ...) on seeded synthetic inputs and stores
inputs + outputs.  The fixtures travel to the GPU box; the reference does not.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np

REF = Path(os.environ.get("GAIT_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"
sys.path[:0] = [str(REF / "train"), str(REF / "data" / "WearGait"), str(REF)]

import pandas as pd  # noqa: E402
import torch  # noqa: E402

torch.cuda.FloatTensor = torch.FloatTensor  # shim S2 (CPU construction of GCL/LDAM)

import weargait_encoders as WE  # noqa: E402
import feature_encoder as FE  # noqa: E402
import weargait_train as WT  # noqa: E402
import fbg_fog_train as FT  # noqa: E402
from learning.optimizers.classification_losses import GCLLoss, LDAMLoss  # noqa: E402
from learning.optimizers.multitask_weighting import CAGrad  # noqa: E402
from data_processing import dataloader_weargait as DW  # noqa: E402
from data_processing import dataloader_fbg_fog as DF  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent))
from gait_oracle import synth_weargait_batch, synth_fog_batch  # noqa: E402  (input generators only)


def _np(t):
    return t.detach().cpu().numpy().copy()


def _state(model):
    return {k: _np(v) for k, v in model.state_dict().items()}


def _grads(model):
    return {k: (None if p.grad is None else _np(p.grad)) for k, p in model.named_parameters()}


def _save(name, **arrs):
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                if vv is not None:
                    flat[f"{k}/{kk}"] = np.asarray(vv)
        else:
            flat[k] = np.asarray(v)
    np.savez_compressed(OUT / f"{name}.npz", **flat)
    print(f"  wrote {name}.npz ({(OUT / (name + '.npz')).stat().st_size / 1024:.0f} KiB, {len(flat)} arrays)")


# ------------------------------------------------------------------ WearGait
def weargait_case(name, *, synchronized, wm, use_norm=False, use_cosine=False, B=8, steps=3,
                  alpha=0.5, seed=43, model_kw=None, drw=False, T=64, masks=None):
    WT.set_seed(seed)
    model = WE.WearGaitThreeModal(synchronized=synchronized, use_norm=use_norm, use_cosine=use_cosine,
                                  **(model_kw or {}))
    # nudge LN affine / biases away from their (1, 0) init so parity exercises them
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if ".ln" in k or ".norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    state0 = _state(model)
    batches = [synth_weargait_batch(B, T=T, seed=100 + i) for i in range(2)]
    if synchronized:
        labels = [[b[1]] * 3 for b in batches]
    else:
        labels = []
        for i, b in enumerate(batches):
            r = np.random.default_rng(500 + i)
            labels.append([b[1], r.permutation(b[1]), r.permutation(b[1])])
    counts = [np.bincount(np.concatenate([l[s] for l in labels]), minlength=2).tolist() for s in range(3)]
    # unequal counts are required by GCL (0/0 otherwise, classification_losses.py:104)
    counts = [[c[0] + 3, c[1] + 11] for c in counts]
    args = SimpleNamespace(wm=wm, gcl_m=0.2, gcl_s=25.0, noise_mul=0.0)
    cnt = {"walkway": counts[0], "insole": counts[1], "imu": counts[2]}
    crit = WT.make_criteria(args, cnt)
    if drw:
        for c, k in zip(crit, ("walkway", "insole", "imu")):
            c.weight = WT.inv_freq_weights(cnt[k])
    cags = {k: CAGrad(n_tasks=k, device=torch.device("cpu"), c=alpha) for k in (1, 2, 3)}   # one per live-task count
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    rec = {}
    for st in range(steps):
        xs, _ = batches[st % 2]
        ys = labels[st % 2]
        xt = [torch.from_numpy(x) for x in xs]
        yt = [torch.from_numpy(y) for y in ys]
        model.train()
        mask = (True, True, True) if masks is None else masks[st]
        xt = [WT._maybe_zero(x, u) for x, u in zip(xt, mask)]           # weargait_train.py:355-358
        lw, li, lm = model(*xt)
        Lall = [crit[0](lw, yt[0]), crit[1](li, yt[1]), crit[2](lm, yt[2])]
        L = [l if u else None for l, u in zip(Lall, mask)]
        live = [l for l in L if l is not None]
        cag = cags[len(live)]
        # what CAGrad sees: per-task shared-gradient matrix (recomputed, non-destructively)
        shared = list(model.get_shared_parameters())
        G = torch.stack([torch.cat([g_.reshape(-1) for g_ in torch.autograd.grad(l, shared, retain_graph=True)])
                         for l in live], 1)
        # step_cagrad_three discards CAGrad's return value; capture it with the same call
        # signature it uses (weargait_train.py:214), forwarding unchanged.
        got = {}
        orig_backward = cag.backward

        def spy(*a, **k):
            out = orig_backward(*a, **k)
            got["extra"] = out[1]
            return out
        cag.backward = spy
        WT.step_cagrad_three(model, L[0], L[1], L[2], opt, cag)
        cag.backward = orig_backward
        rec[f"s{st}"] = dict(logits=np.stack([_np(lw), _np(li), _np(lm)]),
                             losses=np.array([float(l.detach()) for l in Lall]), G=_np(G),
                             GTG=np.asarray(got["extra"]["GTG"]), w=np.asarray(got["extra"]["weights"]))
        for k, v in _grads(model).items():
            if v is not None:
                rec[f"s{st}"][f"grad:{k}"] = v
        for k, v in _state(model).items():
            rec[f"s{st}"][f"param:{k}"] = v
    flat = {}
    for s, d in rec.items():
        for k, v in d.items():
            flat[f"{s}/{k}"] = v
    for i, (xs, _) in enumerate(batches):
        for j, x in enumerate(xs):
            flat[f"x{i}_{j}"] = x
        for j in range(3):
            flat[f"y{i}_{j}"] = labels[i][j]
    meta = dict(synchronized=synchronized, wm=wm, use_norm=use_norm, use_cosine=use_cosine, B=B, steps=steps,
                alpha=alpha, counts=counts, drw=drw, T=T, model_kw=model_kw or {}, masks=masks)
    _save(name, meta=json.dumps(meta), state0=state0, **flat)


def weargait_mask_case(name, B=32, seed=7):
    WT.set_seed(seed)
    model = WE.WearGaitThreeModal(synchronized=True)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(3.0)  # spread logits so argmax is not degenerate
    xs, y = synth_weargait_batch(B, seed=seed)
    batch = {"xs": [torch.from_numpy(x) for x in xs], "y": torch.from_numpy(y)}
    accs = {k: WT.eval_with_mask(model, [batch], False, k) for k in WT.MASK_COMBOS}
    # async flavour: mean per-batch accuracy of the enabled streams
    abatch = {"walkway": batch["xs"][0], "insole": batch["xs"][1], "imu": batch["xs"][2],
              "y": {"walkway": batch["y"], "insole": batch["y"], "imu": batch["y"]}}
    am = WE.WearGaitThreeModal(synchronized=False)
    am.load_state_dict({k: v for k, v in model.state_dict().items() if not k.startswith("_shared")}, strict=True)
    aacc = {k: WT.eval_with_mask(am, [abatch], True, k)["macro_enabled"] for k in WT.MASK_COMBOS}
    _save(name, state0=_state(model), x0=xs[0], x1=xs[1], x2=xs[2], y=y,
          mask_names=np.array(list(WT.MASK_COMBOS)), mask_table=np.array(list(WT.MASK_COMBOS.values())),
          acc_sync=np.array([accs[k] for k in WT.MASK_COMBOS]),
          acc_async=np.array([aacc[k] for k in WT.MASK_COMBOS]))


# ------------------------------------------------------------------ FoG / FBG
def fog_case(name, *, dataset, synchronized, wm, use_nc=False, B=8, steps=3, alpha=0.1, seed=43, cons_lambda=1.0):
    FT.set_random_seed(seed)
    params = FT.FBG_FOG_PARAMS[dataset]
    args = SimpleNamespace(modality="multimodal", synchronized_loading=synchronized, wm=wm, use_norm_and_cos=use_nc,
                           consistency_lambda=cons_lambda, gcl_m=0.2, gcl_s=25.0, noise_mul=0, ldam_m=0.5, ldam_s=30.0,
                           alpha=alpha, max_norm=1.0, dataset=dataset, drw_warmup=0)
    dev = torch.device("cpu")
    model = FT.choose_model(args, params, dev)
    state0 = _state(model)
    J = params["skeleton_input_dim"] // 3
    batches = [synth_fog_batch(B, seed=200 + i, pose_len=params["pose_length"], sens_len=params["sensor_length"],
                               joints=J, sens_ch=params["sensor_in_channels"]) for i in range(2)]
    ys = [b[2] for b in batches]
    yt = ys if synchronized else [np.random.default_rng(900 + i).permutation(y) for i, y in enumerate(ys)]
    sk_counts = (np.bincount(np.concatenate(ys), minlength=3) + np.array([7, 3, 1])).tolist()
    se_counts = (np.bincount(np.concatenate(yt), minlength=3) + np.array([5, 2, 9])).tolist()
    ldam_skel, ldam_sens, gcl_skel, gcl_sens, drw = FT.build_branch_losses(args, sk_counts, se_counts, dev)
    opt = torch.optim.SGD(model.parameters(), lr=params["learning_rate"], momentum=0.9, weight_decay=1e-4)
    cag = CAGrad(n_tasks=2, device=dev, c=alpha, max_norm=1.0)
    flat = {}
    for st in range(steps):
        sk, se, _ = batches[st % 2]
        batch = {"skeleton": torch.from_numpy(sk).view(B, params["pose_length"], J, 3),
                 "sensor": torch.from_numpy(se),
                 "label_skeleton": torch.from_numpy(ys[st % 2]), "label_sensor": torch.from_numpy(yt[st % 2])}
        model.train()
        with torch.no_grad():
            ls, lt = model(FT.flatten_skel(batch["skeleton"]).float(), batch["sensor"].float())
        got = {}
        orig = cag.backward

        def spy(*a, **k):
            got["losses"] = [float(l) for l in k["losses"]]
            out = orig(*a, **k)
            got["extra"] = out[1]
            return out
        cag.backward = spy
        lossv, cs, ce, n = FT.process_batch(batch, model, opt, args, sk_counts, se_counts, ldam_skel, ldam_sens,
                                            gcl_skel, gcl_sens, cag, dev, True)
        cag.backward = orig
        flat[f"s{st}/logits"] = np.stack([_np(ls), _np(lt)])
        flat[f"s{st}/losses"] = np.array(got["losses"])
        flat[f"s{st}/loss_mean"] = np.array(lossv)
        flat[f"s{st}/correct"] = np.array([cs, ce, n])
        flat[f"s{st}/GTG"] = np.asarray(got["extra"]["GTG"]); flat[f"s{st}/w"] = np.asarray(got["extra"]["weights"])
        for k, v in _grads(model).items():
            if v is not None:
                flat[f"s{st}/grad:{k}"] = v
        for k, v in _state(model).items():
            flat[f"s{st}/param:{k}"] = v
    for i, (sk, se, _) in enumerate(batches):
        flat[f"sk{i}"] = sk; flat[f"se{i}"] = se; flat[f"ys{i}"] = ys[i]; flat[f"yt{i}"] = yt[i]
    meta = dict(dataset=dataset, synchronized=synchronized, wm=wm, use_nc=use_nc, B=B, steps=steps, alpha=alpha,
                sk_counts=sk_counts, se_counts=se_counts, cons_lambda=cons_lambda, params=params)
    _save(name, meta=json.dumps(meta), state0=state0, **flat)


# ------------------------------------------------------------------ fusion baselines (--baseline)
def baseline_case(name, baseline, synchronized, B=8, steps=2, seed=43):
    WT.set_seed(seed)
    args = SimpleNamespace(baseline=baseline, enc_out_ch=12, backbone_dim=8, shared_out_ch=16, num_classes=2, use_norm=False,
                           use_cosine=False, proj_ch=16, win_len=64)
    WT.DEVICE = torch.device("cpu")
    model = WT.build_model(args, synchronized)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if ".ln" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    flat = {f"state0/{k}": v for k, v in _state(model).items()}
    batches = [synth_weargait_batch(B, seed=400 + i) for i in range(2)]
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    ce = torch.nn.CrossEntropyLoss()
    for st in range(steps):
        xs, y = batches[st % 2]
        r = np.random.default_rng(600 + st)
        ys = [y, y, y] if synchronized else [y, r.permutation(y), r.permutation(y)]
        yt = [torch.from_numpy(v) for v in ys]
        model.train()
        lw, li, lm = model(*[torch.from_numpy(x) for x in xs])
        L = [ce(lw, yt[0]), ce(li, yt[1]), ce(lm, yt[2])]
        WT.step_cagrad_three(model, L[0], L[1], L[2], opt, None)          # baselines: plain averaged loss (:244-248)
        flat[f"s{st}/logits"] = np.stack([_np(lw), _np(li), _np(lm)])
        flat[f"s{st}/losses"] = np.array([float(l.detach()) for l in L])
        for k, v in _grads(model).items():
            if v is not None:
                flat[f"s{st}/grad:{k}"] = v
        for k, v in _state(model).items():
            flat[f"s{st}/param:{k}"] = v
        for j in range(3):
            flat[f"s{st}/y{j}"] = ys[j]
    for i, (xs, _) in enumerate(batches):
        for j, x in enumerate(xs):
            flat[f"x{i}_{j}"] = x
    flat["meta"] = json.dumps(dict(baseline=baseline, synchronized=synchronized, B=B, steps=steps))
    np.savez_compressed(OUT / f"{name}.npz", **flat)
    print(f"  wrote {name}.npz ({(OUT / (name + '.npz')).stat().st_size / 1024:.0f} KiB, {len(flat)} arrays)")


# ------------------------------------------------------------------ 2-stream fusion baselines (baselines/fusion_train.py)
def fog_baseline_case(name, kind, synchronized, B=6, seed=43):
    FT.set_random_seed(seed)
    common = dict(skeleton_input_dim=21, skeleton_output_dim=6, sensor_in_channels=6, sensor_out_channels=6, sensor_length=426,
                  shared_out_channels=16, backbone_dim=8, num_classes=3, synchronized_loading=synchronized)
    if kind == "early":
        model = FE.EarlyFusionModel(**common)
    elif kind == "late":
        model = FE.LateFusionModel(**common)
    elif kind == "share_latent":
        model = FE.ShareLatentModel(**common, taskhead_input_dim=8 * 16)
    else:
        model = FE.CheapXAttnModel(**common)
    sk, se, y = synth_fog_batch(B, seed=seed)
    r = np.random.default_rng(seed + 1)
    yt = y if synchronized else r.permutation(y)
    out = model(torch.from_numpy(sk), torch.from_numpy(se))
    ce = torch.nn.CrossEntropyLoss()
    if synchronized and kind != "share_latent":                     # fusion_train.py:234-242
        loss = ce(out, torch.from_numpy(y)); logits = [_np(out)]
    else:
        loss = 0.5 * (ce(out[0], torch.from_numpy(y)) + ce(out[1], torch.from_numpy(yt))); logits = [_np(out[0]), _np(out[1])]
    model.zero_grad(); loss.backward()
    flat = {f"state0/{k}": v for k, v in _state(model).items()}
    for k, v in _grads(model).items():
        if v is not None:
            flat[f"grad:{k}"] = v
    flat.update(x_skel=sk, x_sens=se, ys=y, yt=yt, loss=np.array(float(loss.detach())))
    for i, l in enumerate(logits):
        flat[f"logits{i}"] = l
    flat["meta"] = json.dumps(dict(kind=kind, synchronized=synchronized, B=B, sensor_length=426))
    np.savez_compressed(OUT / f"{name}.npz", **flat)
    print(f"  wrote {name}.npz ({(OUT / (name + '.npz')).stat().st_size / 1024:.0f} KiB, {len(flat)} arrays)")


# ------------------------------------------------------------------ single-modality paths
def single_modality_case(name, seed=5, B=6):
    FT.set_random_seed(seed)
    flat = {}
    prm = FT.FBG_FOG_PARAMS["fog"]
    sk, se, y = synth_fog_batch(B, seed=77)
    yt = torch.from_numpy(y)
    args = SimpleNamespace(modality="skeleton", use_norm_and_cos=False, synchronized_loading=False)
    for mod, x in (("skeleton", sk), ("sensor", se)):
        args.modality = mod
        model = FT.choose_model(args, prm, torch.device("cpu"))        # SkelModalityModel / SensorModalityModel (use_norm=True)
        for k, v in _state(model).items():
            flat[f"{mod}/state0/{k}"] = v
        logits = model(torch.from_numpy(x))
        loss = torch.nn.CrossEntropyLoss()(logits, yt)
        model.zero_grad(); loss.backward()
        flat[f"{mod}/logits"] = _np(logits); flat[f"{mod}/loss"] = np.array(float(loss.detach()))
        for k, v in _grads(model).items():
            flat[f"{mod}/grad/{k}"] = v
    flat["sk"] = sk; flat["se"] = se; flat["y"] = y
    # WearGait --single_mod branch: enc_? -> shared backbone -> head_? (weargait_train.py:252-271)
    WT.set_seed(seed)
    model = WE.WearGaitThreeModal(synchronized=True)
    xs, yw = synth_weargait_batch(B, seed=78)
    batch = {"xs": [torch.from_numpy(x) for x in xs], "y": torch.from_numpy(yw)}
    for k, v in _state(model).items():
        flat[f"wg/state0/{k}"] = v
    for mod in ("walkway", "insole", "imu"):
        logits, yy = WT._single_logits_and_labels(model, batch, False, mod)
        loss = torch.nn.CrossEntropyLoss()(logits, yy)
        model.zero_grad(); loss.backward()
        flat[f"wg/{mod}/logits"] = _np(logits); flat[f"wg/{mod}/loss"] = np.array(float(loss.detach()))
        for k, v in _grads(model).items():
            if v is not None:
                flat[f"wg/{mod}/grad/{k}"] = v
    for j in range(3):
        flat[f"wg/x{j}"] = xs[j]
    flat["wg/y"] = yw
    np.savez_compressed(OUT / f"{name}.npz", **flat)
    print(f"  wrote {name}.npz ({(OUT / (name + '.npz')).stat().st_size / 1024:.0f} KiB, {len(flat)} arrays)")


# ------------------------------------------------------------------ CAGrad solver corpus
def cagrad_corpus(name):
    rng = np.random.default_rng(11)
    Gs, gs, ws, alphas = [], [], [], []
    for n in (2, 3):
        for case in range(40):
            P = 24
            G = rng.standard_normal((P, n)).astype(np.float32)
            kind = case % 8
            if kind == 1:   G[:, 1] = G[:, 0] * 0.7                       # collinear
            elif kind == 2: G[:, -1] = -G[:, 0] + 0.05 * G[:, -1]          # conflicting
            elif kind == 3: G *= 1e-3                                      # tiny gradients
            elif kind == 4: G[:, 0] *= 30.0                                # one dominant task
            elif kind == 5: G[:, -1] = 0.0                                 # dead task
            elif kind == 6: G = np.abs(G)                                  # all aligned-ish
            elif kind == 7: G *= 1e2
            alpha = [0.5, 0.1, 0.4, 1.0][case % 4]
            cag = CAGrad(n_tasks=n, device=torch.device("cpu"), c=alpha)
            g, GG, w = cag.cagrad(torch.from_numpy(G), alpha=alpha, rescale=1)
            Gp = np.zeros((P, 3), np.float32); Gp[:, :n] = G
            wp = np.zeros(3); wp[:n] = w
            Gs.append(Gp); gs.append(_np(g)); ws.append(wp); alphas.append((n, alpha))
    _save(name, G=np.stack(Gs), g=np.stack(gs), w=np.stack(ws), n_alpha=np.array(alphas))


# ------------------------------------------------------------------ data path
def _patched_apply_stats(df, stats):  # shim S1: identical body + .copy() after to_numpy
    out = df.copy()
    for c, (m, s) in stats.items():
        if c not in out.columns:
            continue
        x = pd.to_numeric(out[c], errors="coerce").to_numpy(dtype=float).copy()
        x[~np.isfinite(x)] = m if np.isfinite(m) else 0.0
        s_eff = s if (np.isfinite(s) and s > DW.MIN_STD) else DW.MIN_STD
        z = (x - (m if np.isfinite(m) else 0.0)) / s_eff
        out[c] = np.nan_to_num(z, nan=0.0, posinf=0.0, neginf=0.0)
    return out


def data_case(name):
    DW.apply_stats = _patched_apply_stats
    rng = np.random.default_rng(0)
    flat = {}
    # A1
    cases = [(0, 64, 64), (63, 64, 64), (64, 64, 64), (65, 64, 64), (200, 64, 64), (200, 64, 32), (1000, 256, 256),
             (130, 64, 16), (5, 1, 1)]
    flat["win_cases"] = np.array(cases)
    for i, (n, w, h) in enumerate(cases):
        flat[f"win_{i}"] = np.array(DW.window_indices(n, w, h), dtype=np.int64).reshape(-1, 3)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        sids = [f"pd{i:02d}" for i in range(3)] + [f"hc{i:02d}" for i in range(3)]
        raw = {}
        for si, sid in enumerate(sids):
            N = int(rng.integers(150, 420))
            Ni = N - int(rng.integers(0, 70)); Nm = N - int(rng.integers(0, 70))
            if si == 4:
                Ni = 40  # fewer than one window -> subject dropped from the sync map
            walk = pd.DataFrame({"L Foot Pressure_BW": rng.random(N), "R Foot Pressure_BW": rng.random(N)})
            ins = {c: rng.standard_normal(Ni) * (1 + j) + j for j, c in enumerate(DW.INSOLE_NUMERIC[:7])}
            ins_df = pd.DataFrame(ins)
            ins_df["Linsole_Acc"] = [tuple(v) for v in rng.standard_normal((Ni, 3))]
            ins_df["Rinsole_Acc"] = [tuple(v) for v in rng.standard_normal((Ni, 3)) * 3 + 1]
            if si == 1:
                ins_df.loc[5:20, "LCoP_X"] = np.nan            # NaN run -> mean fill
            if si == 2:
                ins_df["RCoP_Y"] = np.nan                      # all-NaN column -> mean fill
            imu_df = pd.DataFrame({f"{s}_FreeAcc": [tuple(v) for v in rng.standard_normal((Nm, 3)) * (2 if si < 3 else 1)]
                                   for s in DW.IMU_SITES})
            if si == 3:
                imu_df = imu_df.drop(columns=["R_LatShank_FreeAcc"])   # missing site -> mean -> 0 after z-score
            walk.to_pickle(td / f"{sid}_walkway.pkl"); ins_df.to_pickle(td / f"{sid}_insole.pkl")
            imu_df.to_pickle(td / f"{sid}_imu.pkl")
            raw[sid] = (DW.ensure_cols(walk, DW.WALKWAY_FIXED).to_numpy(dtype=float),
                        DW.expand_insole(ins_df).to_numpy(dtype=float), DW.expand_imu(imu_df).to_numpy(dtype=float))
        train, test = sids[:2] + sids[3:5], [sids[2], sids[5]]
        prep = DW.prepare_split(train, test, data_dir=td, win=64, hop=64)
        stats = prep["stats"]
        flat["stat_cols"] = np.array(DW.INSOLE_FIXED + DW.IMU_FIXED)
        flat["stat_mean"] = np.array([stats[c][0] for c in DW.INSOLE_FIXED + DW.IMU_FIXED])
        flat["stat_std"] = np.array([stats[c][1] for c in DW.INSOLE_FIXED + DW.IMU_FIXED])
        flat["sids"] = np.array(sids); flat["train"] = np.array(train); flat["test"] = np.array(test)
        for sid in sids:
            for m, a in zip(("walkway", "insole", "imu"), raw[sid]):
                flat[f"raw/{sid}/{m}"] = a
        for split in ("train", "test"):
            flat[f"{split}_sync"] = np.array([[t[0].split("|")[0], t[0].split("|")[2]] for t in prep[f"{split}_sync"]])
            st = prep[f"{split}_stores"]
            for m in ("walkway", "insole", "imu"):
                keys = sorted(st[m].keys())
                flat[f"{split}_keys/{m}"] = np.array(keys)
                flat[f"{split}_win/{m}"] = np.stack([st[m][k] for k in keys]) if keys else np.zeros((0,))
        subj2label = DW.build_subj2label(sids[:3], sids[3:])
        ds = DW.WearGaitMultiAsyncDataset(prep["train_stores"], ("walkway", "insole", "imu"), subj2label, seed=43)
        flat["async_perm_seed43"] = np.array([ds._perms[m] for m in ds.modalities])
        ds.reseed(44)
        flat["async_perm_seed44"] = np.array([ds._perms[m] for m in ds.modalities])
        item = ds[3]
        flat["async_item3_keys"] = np.array([item["keys"][m] for m in ds.modalities])
        flat["async_item3_y"] = np.array([int(item["y"][m]) for m in ds.modalities])
        sd = DW.WearGaitSyncDataset(tuple(prep["train_stores"][m] for m in ds.modalities), prep["train_sync"], subj2label)
        b = DW._collate_sync([sd[i] for i in range(min(4, len(sd)))])
        for j in range(3):
            flat[f"sync_batch_x{j}"] = b["xs"][j].numpy()
        flat["sync_batch_y"] = b["y"].numpy()
        folds = DW.make_fixed_balanced_folds_no_overlap(sids[:3], sids[3:], n_folds=1, per_class=1, seed=43)
        flat["fold0_test"] = np.array(folds[0][1])
        folds = DW.make_fixed_balanced_folds_no_overlap(sids[:3], sids[3:], n_folds=3, per_class=1, seed=7)
        flat["folds3_seed7"] = np.array([[",".join(tr), ",".join(te)] for tr, te in folds])
        # loader order: what the trainer sees over two epochs (train pass, then test pass, sharing one generator,
        # weargait_train.py:551,573-589; make_sync_loaders/make_async_loaders :420-455)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr, te = DW.make_sync_loaders(prep, subj2label, batch_size=4, num_workers=0, seed=43)
            for ep in range(2):
                for nm, ld in (("train", tr), ("test", te)):
                    ks, ys = [], []
                    for b in ld:
                        ks += [k[0].split("|")[0] + "|" + k[0].split("|")[2] for k in b["keys"]]; ys += b["y"].tolist()
                    flat[f"loader_sync/{nm}_ep{ep}_keys"] = np.array(ks); flat[f"loader_sync/{nm}_ep{ep}_y"] = np.array(ys)
                    if ep == 0 and nm == "train":
                        for j in range(3):
                            flat[f"loader_sync/last_batch_x{j}"] = b["xs"][j].numpy()
            tr, te = DW.make_async_loaders(prep, subj2label, batch_size=4, num_workers=0, seed=43)
            for ep in range(1, 3):
                tr.dataset.reseed(43 + ep)                                  # weargait_train.py:574-575
                for nm, ld in (("train", tr), ("test", te)):
                    ks = {m: [] for m in ds.modalities}; ys = {m: [] for m in ds.modalities}
                    for b in ld:
                        for m in ds.modalities:
                            ks[m] += list(b["keys"][m]); ys[m] += b["y"][m].tolist()
                    for m in ds.modalities:
                        flat[f"loader_async/{nm}_ep{ep}_keys/{m}"] = np.array(ks[m])
                        flat[f"loader_async/{nm}_ep{ep}_y/{m}"] = np.array(ys[m])
                    if ep == 1 and nm == "train":
                        for m in ds.modalities:
                            flat[f"loader_async/last_batch/{m}"] = b[m].numpy()
    # A5 FoG clip prep
    for i, (T_, Ts) in enumerate([(57, 300), (101, 426), (140, 500)]):
        pose = rng.random((T_, 7, 3)) * 5 - 1; sens = rng.standard_normal((Ts, 6))
        pd_ = DF.normalize_poses(DF.center_poses({"k": pose}), "minmax")["k"]
        flat[f"fog_pose_in{i}"] = pose; flat[f"fog_sens_in{i}"] = sens
        flat[f"fog_pose_out{i}"] = DF.pad_or_trim(pd_, 101).astype(np.float32).reshape(101, 21)
        flat[f"fog_sens_out{i}"] = DF.pad_or_trim(sens, 426).astype(np.float32)
    # adaptive pooling tables (ATen), recovered by pooling one-hot time series
    for L, O in [(64, 8), (101, 8), (426, 101), (65, 101), (256, 8), (128, 8)]:
        eye = torch.eye(L).unsqueeze(0)                      # (1, L(ch), L(time))
        flat[f"pool_{L}_{O}"] = torch.nn.AdaptiveAvgPool1d(O)(eye)[0].numpy().T   # (O, L) weights
    _save(name, **flat)



# ------------------------------------------------------------------ FoG / FBG loaders (A5: pairing, oversampling, order)
def fog_loader_case(name):
    """create_fusion_loaders (dataloader_fbg_fog.py:269-494) on synthetic readers (oracle/ref_harness.synthetic_fog_reader):
    key lists / pairs of both datasets after all the oversampling, and what two epochs of both loaders deliver."""
    import warnings
    import ref_harness as H
    flat = {}
    for cname, dataset, sync, modality, pad_skel, pad_sens in H.FOG_LOADER_CASES:
        reader, subs = H.synthetic_fog_reader(dataset, seed=5)
        tr_s, ev_s = H.fog_loader_split(subs, dataset)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr, ev = DF.create_fusion_loaders(dataset, reader, tr_s, ev_s, batch_size=7, synchronized=sync, seed=43, num_workers=0,
                                              pad_skel=pad_skel, pad_sens=pad_sens, modality=modality)
            for nm, ld in (("train", tr), ("eval", ev)):
                ds = ld.dataset
                flat[f"{cname}/{nm}_pose_keys"] = np.array(ds.pose_ds.keys); flat[f"{cname}/{nm}_sens_keys"] = np.array(ds.sens_ds.keys)
                flat[f"{cname}/{nm}_len"] = np.array(len(ds))
                if sync:
                    flat[f"{cname}/{nm}_pairs"] = np.array(ds.pairs)
            for k, v in H.loader_trace(tr, ev).items():
                flat[f"{cname}/{k}"] = v
    pm = DF.group_by_subject(["A_x_1", "A_y_2", "B_x_1", "A_x_3"]); sm = DF.group_by_subject(["A_x_1", "A_q_x_1", "B_x_1", "B_z_1", "C_x_1"])
    flat["pairs_small"] = np.array(DF.build_synced_pairs(pm, sm))
    _save(name, **flat)

def versions():
    import scipy, sklearn
    v = dict(torch=torch.__version__, numpy=np.__version__, scipy=scipy.__version__, pandas=pd.__version__,
             sklearn=sklearn.__version__, python=sys.version.split()[0],
             reference_pins=dict(torch="2.3.1", numpy="1.26.4", scipy="1.12.0", pandas="2.2.1"),
             note="goldens produced by running /root/reference on CPU with shims S1 (apply_stats .copy()) and S2 "
                  "(torch.cuda.FloatTensor=torch.FloatTensor); see SURVEY.md 8(c)")
    (OUT / "VERSIONS.json").write_text(json.dumps(v, indent=1))


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--only", default=None); a = ap.parse_args()
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)
    jobs = {
        "wg_sync_gcl": lambda: weargait_case("wg_sync_gcl", synchronized=True, wm="gcl"),
        "wg_sync_gcl_drw": lambda: weargait_case("wg_sync_gcl_drw", synchronized=True, wm="gcl", drw=True, steps=2),
        "wg_async_ce": lambda: weargait_case("wg_async_ce", synchronized=False, wm="ce"),
        "wg_async_gcl": lambda: weargait_case("wg_async_gcl", synchronized=False, wm="gcl", steps=2),
        "wg_sync_classwt_nc": lambda: weargait_case("wg_sync_classwt_nc", synchronized=True, wm="class_wt",
                                                    use_norm=True, use_cosine=True, steps=2),
        "wg_sync_norm": lambda: weargait_case("wg_sync_norm", synchronized=True, wm="gcl", use_norm=True, steps=2),
        "wg_scaled": lambda: weargait_case("wg_scaled", synchronized=True, wm="gcl", B=3, steps=2, T=256,
                                           model_kw=dict(enc_out_ch=24, shared_out_ch=32)),
        "wg_masks": lambda: weargait_mask_case("wg_masks"),
        "wg_sync_gcl_dropped": lambda: weargait_case("wg_sync_gcl_dropped", synchronized=True, wm="gcl", steps=4,
                                                     masks=[(True, False, True), (False, True, False), (True, True, True), (False, True, True)]),
        "wg_async_gcl_dropped": lambda: weargait_case("wg_async_gcl_dropped", synchronized=False, wm="gcl", steps=3,
                                                      masks=[(False, True, True), (True, False, False), (True, True, False)]),
        "fog_async_gcl": lambda: fog_case("fog_async_gcl", dataset="fog", synchronized=False, wm="gcl"),
        "fog_sync_gcl": lambda: fog_case("fog_sync_gcl", dataset="fog", synchronized=True, wm="gcl"),
        "fog_async_ldam": lambda: fog_case("fog_async_ldam", dataset="fog", synchronized=False, wm="ldam", steps=2),
        "fog_sync_ce_nc": lambda: fog_case("fog_sync_ce_nc", dataset="fog", synchronized=True, wm="ce", use_nc=True, steps=2),
        "fbg_async_classwt": lambda: fog_case("fbg_async_classwt", dataset="fbg", synchronized=False, wm="class_wt", steps=2),
        "single_modality": lambda: single_modality_case("single_modality"),
        "bl_late_sync": lambda: baseline_case("bl_late_sync", "late_fusion", True),
        "bl_late_async": lambda: baseline_case("bl_late_async", "late_fusion", False),
        "bl_shared_latent_sync": lambda: baseline_case("bl_shared_latent_sync", "shared_latent", True),
        "bl_shared_latent_async": lambda: baseline_case("bl_shared_latent_async", "shared_latent", False),
        **{f"fogbl_{k}_{'sync' if sy else 'async'}": (lambda k=k, sy=sy: fog_baseline_case(f"fogbl_{k}_{'sync' if sy else 'async'}", k, sy))
           for k in ("early", "late", "share_latent", "cheap_xattn") for sy in (True, False)},
        "bl_early_sync": lambda: baseline_case("bl_early_sync", "early_fusion", True),
        "bl_early_async": lambda: baseline_case("bl_early_async", "early_fusion", False),
        "bl_xattn_sync": lambda: baseline_case("bl_xattn_sync", "cheap_xattn", True),
        "bl_xattn_async": lambda: baseline_case("bl_xattn_async", "cheap_xattn", False),
        "cagrad_corpus": lambda: cagrad_corpus("cagrad_corpus"),
        "data_path": lambda: data_case("data_path"),
        "fog_loaders": lambda: fog_loader_case("fog_loaders"),
    }
    for k, fn in jobs.items():
        if a.only and a.only != k:
            continue
        print(k); fn()
    versions()


if __name__ == "__main__":
    main()
