import json
"""Pins the CPU oracle (oracle/gait_oracle.py) to golden vectors produced by
running the reference itself (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import gait_oracle as O
from conftest import load_golden, sub

WG_CASES = ["wg_sync_gcl", "wg_sync_gcl_drw", "wg_async_ce", "wg_async_gcl", "wg_sync_classwt_nc",
            "wg_sync_norm", "wg_scaled", "wg_sync_gcl_dropped", "wg_async_gcl_dropped"]
FOG_CASES = ["fog_async_gcl", "fog_sync_gcl", "fog_async_ldam", "fog_sync_ce_nc", "fbg_async_classwt"]


def _close(a, b, rtol=2e-5, atol=2e-6):
    """|a-b| <= atol + rtol*|b| elementwise, OR max|a-b| <= rtol * max|b| (fp32 cancellation in
    large-magnitude tensors)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if not a.size:
        return
    err = np.abs(a - b).max(); scale = np.abs(b).max()
    ok = np.allclose(a, b, rtol=rtol, atol=atol) or err <= rtol * scale
    assert ok, f"max abs err {err:.3e}, ref max {scale:.3e}"


@pytest.mark.parametrize("name", WG_CASES)
def test_weargait_step_matches_reference(name):
    g = load_golden(name); meta = g["meta"]
    sync = meta["synchronized"]
    p = O.canonical_params(sub(g, "state0"), sync)
    bufs = {}
    weights = None
    if meta["wm"] == "class_wt" or meta["drw"]:
        weights = [O.inv_freq_weights(c) for c in meta["counts"]]
    for st in range(meta["steps"]):
        i = st % 2
        xs = [torch.from_numpy(g[f"x{i}_{j}"]) for j in range(3)]
        ys = [torch.from_numpy(g[f"y{i}_{j}"]) for j in range(3)]
        mask = tuple(meta["masks"][st]) if meta.get("masks") else None
        ex = O.weargait_train_step(p, bufs, xs, ys, synchronized=sync, wm=meta["wm"], counts=meta["counts"],
                                   alpha=meta["alpha"], weights=weights, tasks=mask)
        ref = sub(g, f"s{st}")
        _close(torch.stack(ex["logits"]).numpy(), ref["logits"])
        _close(ex["losses"], ref["losses"])
        _close(ex["G"].numpy(), ref["G"], rtol=1e-4, atol=1e-6)
        _close(ex["GTG"], ref["GTG"], rtol=1e-4, atol=1e-7)
        _close(ex["weights"], ref["w"], rtol=1e-3, atol=1e-4)      # same SciPy, tiny input differences
        for k, v in ref.items():
            if k.startswith("grad:"):
                name_ = k[5:]
                if sync and (name_.startswith("head_i.") or name_.startswith("head_m.") or name_.startswith("_shared")):
                    continue
                _close(ex["grads"][name_].numpy(), v, rtol=2e-4, atol=2e-6)
        # parameters no loss reaches stay gradient-free (enc_i.ln1; encoders of dropped streams)
        assert ex["grads"].get("enc_i.ln1.weight") is None and "grad:enc_i.ln1.weight" not in ref
        if mask is not None:
            for i, pre in enumerate(("enc_w.", "enc_i.", "enc_m.")):
                if not mask[i]:
                    assert not any(k.startswith("grad:" + pre) for k in ref)
        for k, v in ref.items():
            if k.startswith("param:"):
                name_ = k[6:]
                if name_ in p:
                    _close(p[name_].detach().numpy(), v, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", FOG_CASES)
def test_fog_step_matches_reference(name):
    g = load_golden(name); meta = g["meta"]
    sync = meta["synchronized"]; prm = meta["params"]
    p = {k: torch.tensor(v).requires_grad_(True) for k, v in sub(g, "state0").items()}
    bufs = {}
    for st in range(meta["steps"]):
        i = st % 2
        ex = O.fog_train_step(p, bufs, torch.from_numpy(g[f"sk{i}"]), torch.from_numpy(g[f"se{i}"]),
                              torch.from_numpy(g[f"ys{i}"]), torch.from_numpy(g[f"yt{i}"]),
                              sensor_length=prm["sensor_length"], synchronized=sync, wm=meta["wm"],
                              counts=[meta["sk_counts"], meta["se_counts"]], alpha=meta["alpha"],
                              consistency_lambda=meta["cons_lambda"], bdim=prm["backbone_dim"])
        ref = sub(g, f"s{st}")
        _close(torch.stack(ex["logits"]).numpy(), ref["logits"])
        _close(ex["losses"], ref["losses"])
        _close(ex["GTG"], ref["GTG"], rtol=1e-4, atol=1e-7)
        _close(ex["weights"], ref["w"], rtol=1e-3, atol=1e-4)
        for k, v in ref.items():
            if k.startswith("grad:"):
                _close(ex["grads"][k[5:]].numpy(), v, rtol=2e-4, atol=2e-6)
            if k.startswith("param:"):
                _close(p[k[6:]].detach().numpy(), v, rtol=1e-5, atol=1e-7)


def test_masks_match_reference():
    g = load_golden("wg_masks")
    assert list(g["mask_names"]) == list(O.MASK_COMBOS)
    assert (g["mask_table"] == np.array(list(O.MASK_COMBOS.values()))).all()
    p = O.canonical_params(sub(g, "state0"), True)
    xs = [torch.from_numpy(g[f"x{j}"]) for j in range(3)]; y = torch.from_numpy(g["y"])
    for i, k in enumerate(O.MASK_COMBOS):
        c, n = O.eval_mask_sync(p, xs, y, O.MASK_COMBOS[k])
        assert abs(100.0 * c / n - float(g["acc_sync"][i])) < 1e-9


def test_cagrad_corpus_matches_reference():
    g = load_golden("cagrad_corpus")
    for G, gref, wref, (n, alpha) in zip(g["G"], g["g"], g["w"], g["n_alpha"]):
        n = int(n)
        gg, A, w = O.cagrad_combine(torch.from_numpy(G[:, :n].copy()), float(alpha))
        _close(w, wref[:n], rtol=1e-6, atol=1e-9)
        _close(gg.numpy(), gref, rtol=1e-5, atol=1e-8)


def test_data_path_matches_reference():
    g = load_golden("data_path")
    # A1 window indices: bit exact
    for i, (n, w, h) in enumerate(g["win_cases"]):
        got = np.array(O.window_indices(int(n), int(w), int(h)), dtype=np.int64).reshape(-1, 3)
        assert (got == g[f"win_{i}"]).all()
    sids = [str(s) for s in g["sids"]]; train = [str(s) for s in g["train"]]; test = [str(s) for s in g["test"]]
    raw = {s: [g[f"raw/{s}/{m}"] for m in ("walkway", "insole", "imu")] for s in sids}
    # A2 statistics on train only (insole 13 + imu 24 channels)
    mi, si, seen_i = O.fit_channel_stats([raw[s][1] for s in train])
    mm, sm, seen_m = O.fit_channel_stats([raw[s][2] for s in train])
    np.testing.assert_allclose(np.concatenate([mi, mm]), g["stat_mean"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(np.concatenate([si, sm]), g["stat_std"], rtol=1e-13, atol=0)
    # windows after normalisation: float64 bit exact; sync map: bit exact
    for split, subs in (("train", train), ("test", test)):
        nwin = {}
        for s in subs:
            xw = raw[s][0]
            xi = O.apply_channel_stats(raw[s][1], mi, si, seen_i)
            xm = O.apply_channel_stats(raw[s][2], mm, sm, seen_m)
            nwin[s] = []
            for m, x in zip(("walkway", "insole", "imu"), (xw, xi, xm)):
                idx = O.window_indices(len(x), 64, 64)
                nwin[s].append(len(idx))
                keys = [str(k) for k in g[f"{split}_keys/{m}"]]
                for wid, a, b in idx:
                    ref = g[f"{split}_win/{m}"][keys.index(f"{s}|{m}|{wid}")]
                    assert np.array_equal(x[a:b], ref), (s, m, wid, np.abs(x[a:b] - ref).max())
        got = O.sync_index_map(nwin)
        assert [[s, str(w)] for s, w in got] == g[f"{split}_sync"].tolist()
        if split == "train":
            lens = [len(g[f"train_keys/{m}"]) for m in ("walkway", "insole", "imu")]
            assert O.async_permutations(lens, 43) == g["async_perm_seed43"].tolist()
            assert O.async_permutations(lens, 44) == g["async_perm_seed44"].tolist()
            for j, m in enumerate(("walkway", "insole", "imu")):
                keys = O.async_key_order([str(k) for k in g[f"train_keys/{m}"]])
                assert keys[O.async_permutations(lens, 44)[j][3]] == str(g["async_item3_keys"][j])
    # A5 FoG clip preparation: float32 bit exact
    for i in range(3):
        assert np.array_equal(O.prepare_pose_clip(g[f"fog_pose_in{i}"], 101), g[f"fog_pose_out{i}"])
        assert np.array_equal(O.prepare_sensor_clip(g[f"fog_sens_in{i}"], 426), g[f"fog_sens_out{i}"])
    # adaptive pooling bin tables
    for L, Oo in [(64, 8), (101, 8), (426, 101), (65, 101), (256, 8), (128, 8)]:
        Wt = np.zeros((Oo, L), np.float32)
        for i, (a, b) in enumerate(O.adaptive_bins(L, Oo)):
            Wt[i, a:b] = np.float32(1.0) / np.float32(b - a)
        np.testing.assert_allclose(Wt, g[f"pool_{L}_{Oo}"], rtol=1e-6, atol=0)
        assert ((Wt > 0) == (g[f"pool_{L}_{Oo}"] > 0)).all()


def test_single_modality_paths_match_reference():
    g = load_golden("single_modality")
    y = torch.from_numpy(g["y"])
    for mod, xk in (("skeleton", "sk"), ("sensor", "se")):
        p = {k: torch.tensor(v).requires_grad_(True) for k, v in sub(sub(g, mod), "state0").items()}
        lg = O.fog_single_forward(p, torch.from_numpy(g[xk]), mod)
        _close(lg.detach().numpy(), g[f"{mod}/logits"])
        loss = O.weighted_ce(lg, y); loss.backward()
        _close(float(loss), g[f"{mod}/loss"])
        for k, v in sub(sub(g, mod), "grad").items():
            _close(p[k].grad.numpy(), v, rtol=2e-4, atol=2e-6)
    yw = torch.from_numpy(g["wg/y"])
    for j, mod in enumerate(("walkway", "insole", "imu")):
        p = O.canonical_params(sub(sub(g, "wg"), "state0"), True)
        lg = O.weargait_single_forward(O._HeadAlias(p), torch.from_numpy(g[f"wg/x{j}"]), mod)
        _close(lg.detach().numpy(), g[f"wg/{mod}/logits"])
        loss = O.weighted_ce(lg, yw); loss.backward()
        for k, v in sub(sub(sub(g, "wg"), mod), "grad").items():
            if k in p:
                _close(p[k].grad.numpy(), v, rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("name", ["bl_late_sync", "bl_late_async", "bl_shared_latent_sync", "bl_shared_latent_async",
                                  "bl_early_sync", "bl_early_async", "bl_xattn_sync", "bl_xattn_async"])
def test_fusion_baselines_match_reference(name):
    """The four 3-stream fusion baselines trained the way ``--baseline`` trains them (plain mean of the CE losses).
    EarlyFusion3 / CheapXAttn3 have no CUDA path yet: the oracle and its goldens are the parity infrastructure for it."""
    z = load_golden(name)
    meta = z["meta"]
    sync = meta["synchronized"]
    p = O.canonical_params(sub(z, "state0"), sync)
    bufs = {}
    for st in range(meta["steps"]):
        xs = [torch.from_numpy(z[f"x{st % 2}_{j}"]) for j in range(3)]
        ys = [torch.from_numpy(z[f"s{st}/y{j}"]) for j in range(3)]
        out = O.baseline_train_step(p, bufs, xs, ys, kind=meta["baseline"], synchronized=sync)
        for j in range(3):
            _close(out["logits"][j].numpy(), z[f"s{st}/logits"][j])
        _close(np.array(out["losses"]), z[f"s{st}/losses"])
        for k, g in out["grads"].items():
            gk = f"s{st}/grad:{k}"
            if gk not in z and k.startswith("head_w."):      # named_parameters() de-duplicates to the first name
                gk = f"s{st}/grad:_shared_head.{k[7:]}"
            if g is None:
                assert gk not in z, k
            else:
                _close(g.numpy(), z[gk])
        for k, v in p.items():
            _close(v.detach().numpy(), z[f"s{st}/param:{k}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", ["early", "late", "share_latent", "cheap_xattn"])
@pytest.mark.parametrize("sync", [True, False])
def test_two_stream_fusion_baselines_match_reference(kind, sync):
    """The 2-stream twins behind baselines/fusion_train.py (feature_encoder.py:346-596): logits, loss and every
    gradient of the oracle restatement against the reference classes.  No CUDA path yet: this
    pins the oracle it will be checked against."""
    z = load_golden(f"fogbl_{kind}_{'sync' if sync else 'async'}")
    p = {k: torch.tensor(v, dtype=torch.float32).requires_grad_(True) for k, v in sub(z, "state0").items()}
    out = O.fog_baseline_forward(p, torch.from_numpy(z["x_skel"]), torch.from_numpy(z["x_sens"]), kind,
                                 sensor_length=z["meta"]["sensor_length"], synchronized=sync)
    outs = [out] if torch.is_tensor(out) else list(out)
    for i, l in enumerate(outs):
        _close(l.detach().numpy(), z[f"logits{i}"])
    loss = O.fog_baseline_loss(out, torch.from_numpy(z["ys"]), torch.from_numpy(z["yt"]), kind, sync)
    _close(np.array(float(loss.detach())), z["loss"])
    keys = list(p)
    gs = torch.autograd.grad(loss, [p[k] for k in keys], allow_unused=True)
    n = 0
    for k, g in zip(keys, gs):
        if g is None:
            assert f"grad:{k}" not in z, k
        else:
            _close(g.numpy(), z[f"grad:{k}"]); n += 1
    assert n >= 8
