"""GPU parity: the CUDA path (through the C-ABI) against the golden vectors produced by the reference and
against the CPU oracle on seeded inputs.  Tolerances: fp32 path 1e-5 relative (scale-relative for
large-magnitude tensors); quantities downstream of the CAGrad simplex solve are compared at the
tolerance SLSQP's own ftol=1e-6 stopping rule allows."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sub

pytestmark = pytest.mark.gpu

WG_CASES = ["wg_sync_gcl", "wg_sync_gcl_drw", "wg_async_ce", "wg_async_gcl", "wg_sync_classwt_nc", "wg_sync_norm",
            "wg_scaled"]        # wg_scaled = BASELINE configs[4]: T = 256, enc_out_ch 24 (H = 48), shared_out_ch 32 -> 2-CTA clusters
FOG_ASYNC = ["fog_async_gcl", "fog_async_ldam", "fbg_async_classwt"]
FOG_ALL = ["fog_async_gcl", "fog_sync_gcl", "fog_async_ldam", "fog_sync_ce_nc", "fbg_async_classwt"]


@pytest.fixture(scope="module")
def gk():
    import gaitk
    assert torch.cuda.is_available()
    return gaitk


def close(a, b, rel=1e-5, what=""):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not a.size:
        return
    err = np.abs(a - b).max(); scale = max(np.abs(b).max(), 1e-30)
    ok = np.allclose(a, b, rtol=rel, atol=rel * 1e-1) or err <= rel * scale
    assert ok, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.2e} > {rel:g})"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def wg_model(gk, g):
    meta = g["meta"]
    m = gk.WearGaitThreeModal(synchronized=meta["synchronized"], use_norm=meta["use_norm"], use_cosine=meta["use_cosine"],
                              **meta["model_kw"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "state0").items()}, strict=True)
    return m.cuda()


def wg_criteria(gk, meta):
    crit = []
    for c in meta["counts"]:
        w = None
        if meta["wm"] == "class_wt" or meta["drw"]:
            wt = 1.0 / (torch.tensor(c, dtype=torch.float32) + 1e-8); w = (wt / wt.sum() * len(c)).cuda()
        if meta["wm"] == "gcl":
            crit.append(gk.GCLLoss(cls_num_list=c, m=0.2, s=25.0, noise_mul=0.0, weight=w))
        else:
            crit.append(gk.CrossEntropyLoss(weight=w))
    return crit


def grads_by_name(plan, flat):
    flat = flat.cpu().numpy()
    return {p.name: flat[p.offset:p.offset + p.numel].reshape(p.shape) for p in plan.params}


@pytest.mark.parametrize("name", WG_CASES)
def test_weargait_forward_matches_reference(gk, name):
    g = load_golden(name)
    m = wg_model(gk, g)
    with torch.no_grad():
        out = m(*[dev(g[f"x0_{j}"]) for j in range(3)])
    close(torch.stack(out).cpu().numpy(), g["s0/logits"], 2e-5, "logits")


@pytest.mark.parametrize("name", WG_CASES)
def test_weargait_fused_step_matches_reference(gk, name):
    g = load_golden(name); meta = g["meta"]
    m = wg_model(gk, g)
    step = gk.FusedTrainStep(m, wg_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=2.0)
    for st in range(meta["steps"]):
        i = st % 2
        xs = [dev(g[f"x{i}_{j}"]) for j in range(3)]; ys = [dev(g[f"y{i}_{j}"]) for j in range(3)]
        plan = m.set_window(xs[0].shape[1]).plan()
        gout = torch.zeros(plan.NP, device="cuda")
        loss, correct = step.step(xs, ys, grads_out=gout)
        ref = sub(g, f"s{st}")
        close(loss.cpu().numpy(), ref["losses"], 2e-5, "losses")
        G = step._gbuf[:3 * plan.P].view(3, plan.P).t().cpu().numpy()
        close(G, ref["G"], 5e-5, "G")
        d = step.diag().cpu().numpy()
        close(d[3:12].reshape(3, 3), ref["GTG"], 1e-4, "GTG")
        # the device runs a restatement of SLSQP's own iteration -> the simplex weights match SciPy's
        assert np.abs(d[:3] - ref["w"]).max() < 2e-4, (d[:3], ref["w"])
        got = grads_by_name(plan, gout)
        for k, v in ref.items():
            if not k.startswith("grad:"):
                continue
            nm = k[5:]
            if nm not in got:
                continue                      # aliases of the shared head
            shared = next(p.group for p in plan.params if p.name == nm) == 0
            close(got[nm], v, 2e-4 if shared else 5e-5, f"step {st} grad {nm}")
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


def reference_style_step(model, losses, opt, cagrad, private_twice):
    """What step_cagrad_three (weargait_train.py:187-248) / process_batch (fbg_fog_train.py:146-152) do,
    expressed with the public torch API only -- exercises the autograd bridge and the CAGrad drop-in."""
    opt.zero_grad(set_to_none=True)
    cagrad.backward(losses=losses, shared_parameters=list(model.get_shared_parameters()))
    if private_twice:
        groups = [model.walkway_parameters(), model.insole_parameters(), model.imu_parameters()]
        for i, (L, priv) in enumerate(zip(losses, groups)):
            gs = torch.autograd.grad(L, priv, retain_graph=i < 2, allow_unused=True)
            for p, gg in zip(priv, gs):
                if gg is not None:
                    p.grad = gg if p.grad is None else p.grad.add_(gg)
    opt.step()


@pytest.mark.parametrize("name", ["wg_sync_gcl", "wg_async_ce", "wg_sync_classwt_nc"])
def test_weargait_autograd_path_matches_reference(gk, name):
    g = load_golden(name); meta = g["meta"]
    m = wg_model(gk, g)
    crit = wg_criteria(gk, meta)
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    cag = gk.CAGrad(n_tasks=3, device=torch.device("cuda"), c=meta["alpha"])
    for st in range(meta["steps"]):
        i = st % 2
        xs = [dev(g[f"x{i}_{j}"]) for j in range(3)]; ys = [dev(g[f"y{i}_{j}"]) for j in range(3)]
        m.train()
        lg = m(*xs)
        L = [c(l, y) for c, l, y in zip(crit, lg, ys)]
        ref = sub(g, f"s{st}")
        close(torch.stack([l.detach() for l in L]).cpu().numpy(), ref["losses"], 2e-5, "losses")
        reference_style_step(m, L, opt, cag, True)
        named = dict(m.named_parameters())
        for k, v in ref.items():
            if k.startswith("grad:") and k[5:] in named:
                nm = k[5:]
                shared = any(named[nm] is p for p in m.get_shared_parameters())
                close(named[nm].grad.cpu().numpy(), v, 2e-4 if shared else 5e-5, f"step {st} grad {nm}")
        assert named["enc_i.ln1.weight"].grad is None and named["enc_i.ln1.bias"].grad is None
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


def fog_model(gk, g):
    meta = g["meta"]; prm = meta["params"]
    m = gk.MultiModalMultiTaskModel(
        skeleton_input_dim=prm["skeleton_input_dim"], skeleton_output_dim=prm["skeleton_output_dim"],
        sensor_in_channels=prm["sensor_in_channels"], sensor_out_channels=prm["sensor_out_channels"],
        sensor_length=prm["sensor_length"], shared_out_channels=prm["shared_out_channels"], backbone_dim=prm["backbone_dim"],
        taskhead_input_dim=prm["taskhead_input_dim"], num_classes=prm["num_classes"], use_norm=meta["use_nc"],
        use_cosine=meta["use_nc"], synchronized_loading=meta["synchronized"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "state0").items()}, strict=True)
    return m.cuda()


def fog_criteria(gk, meta):
    out = []
    for c in (meta["sk_counts"], meta["se_counts"]):
        wt = 1.0 / (torch.tensor(c, dtype=torch.float32) + 1e-8); w = (wt / wt.sum() * len(c)).cuda()
        if meta["wm"] == "gcl":
            out.append(gk.GCLLoss(cls_num_list=c, m=0.2, s=25.0, noise_mul=0.0, weight=None))
        elif meta["wm"] == "ldam":
            out.append(gk.LDAMLoss(cls_num_list=c, max_m=0.5, s=30.0, weight=w))
        elif meta["wm"] == "class_wt":
            out.append(gk.CrossEntropyLoss(weight=w))
        else:
            out.append(gk.CrossEntropyLoss())
    return out


@pytest.mark.parametrize("name", FOG_ALL)
def test_fog_forward_matches_reference(gk, name):
    g = load_golden(name)
    m = fog_model(gk, g)
    with torch.no_grad():
        ls, lt = m(dev(g["sk0"]), dev(g["se0"]))
    close(torch.stack([ls, lt]).cpu().numpy(), g["s0/logits"], 2e-5, "logits")


@pytest.mark.parametrize("name", FOG_ASYNC)
def test_fog_fused_step_matches_reference(gk, name):
    g = load_golden(name); meta = g["meta"]
    m = fog_model(gk, g)
    step = gk.FusedTrainStep(m, fog_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=1.0)
    for st in range(meta["steps"]):
        i = st % 2
        xs = [dev(g[f"sk{i}"]), dev(g[f"se{i}"])]; ys = [dev(g[f"ys{i}"]), dev(g[f"yt{i}"])]
        plan = m.set_window(xs[0].shape[1]).plan()
        gout = torch.zeros(plan.NP, device="cuda")
        loss, correct = step.step(xs, ys, grads_out=gout)
        ref = sub(g, f"s{st}")
        close(loss.cpu().numpy(), ref["losses"], 2e-5, "losses")
        assert correct.cpu().numpy().round().astype(int).tolist() == ref["correct"][:2].tolist()
        got = grads_by_name(plan, gout)
        for k, v in ref.items():
            if k.startswith("grad:") and k[5:] in got:
                nm = k[5:]
                shared = next(p.group for p in plan.params if p.name == nm) == 0
                close(got[nm], v, 2e-4 if shared else 5e-5, f"step {st} grad {nm}")
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


def test_fog_fused_step_matches_oracle_many_clips(gk):
    """Seeded random FoG batch with MORE clips than the kernels' grids hold (B = 1001 > 3 CTAs x 148 SMs): every CTA runs
    several clips through the bulk-prefetch ring, odd clips start on a 4-byte boundary only (101 x 21 floats), the sensor
    stream runs in pooled-taps form (csrc/stream_common.cuh, ENC_POOL_LINEAR) -- losses, accuracy and the parameters after
    two steps against the CPU oracle (which convolves all 426 rows and pools afterwards, as the reference does)."""
    import gait_oracle as O
    g = load_golden("fog_async_gcl"); meta = g["meta"]; prm = meta["params"]
    m = fog_model(gk, g)
    B = 1001
    sk, se, y = O.synth_fog_batch(B, seed=4)
    yt = np.random.default_rng(2).permutation(y)
    step = gk.FusedTrainStep(m, fog_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=1.0)
    p = {k: torch.tensor(v).requires_grad_(True) for k, v in sub(g, "state0").items()}; bufs = {}
    for it in range(2):
        loss, correct = step.step([dev(sk), dev(se)], [dev(y), dev(yt)])
        ex = O.fog_train_step(p, bufs, torch.from_numpy(sk), torch.from_numpy(se), torch.from_numpy(y), torch.from_numpy(yt),
                              sensor_length=prm["sensor_length"], synchronized=False, wm=meta["wm"],
                              counts=[meta["sk_counts"], meta["se_counts"]], alpha=meta["alpha"], bdim=prm["backbone_dim"])
        close(loss.cpu().numpy(), ex["losses"], 5e-5, "loss")
        ref_correct = [int((l.argmax(1) == torch.from_numpy(v)).sum()) for l, v in zip(ex["logits"], (y, yt))]
        assert correct.cpu().numpy().round().astype(int).tolist()[:2] == ref_correct
    for k, v in m.state_dict().items():
        if k in p:
            close(v.cpu().numpy(), p[k].detach().numpy(), 2e-5, k)


@pytest.mark.parametrize("name", ["fog_sync_gcl", "fog_sync_ce_nc"])
def test_fog_sync_fused_step_matches_reference(gk, name):
    """Synchronised FoG in the FUSED step: shared head, and -- for wm = gcl -- the symmetric-KL consistency term
    (fbg_fog_train.py:81-89,121-124) inside gaitk_step_grads (logits pass, coupling kernel, one recompute + backward pass per
    (task, stream)): losses, every gradient and the parameters after each step against the reference goldens."""
    g = load_golden(name); meta = g["meta"]
    m = fog_model(gk, g)
    lam = meta["cons_lambda"] if meta["wm"] == "gcl" else 0.0
    step = gk.FusedTrainStep(m, fog_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=1.0, consistency_lambda=lam)
    for st in range(meta["steps"]):
        i = st % 2
        xs = [dev(g[f"sk{i}"]), dev(g[f"se{i}"])]; ys = [dev(g[f"ys{i}"]), dev(g[f"yt{i}"])]
        plan = m.set_window(xs[0].shape[1]).plan()
        gout = torch.zeros(plan.NP, device="cuda")
        loss, correct = step.step(xs, ys, grads_out=gout)
        ref = sub(g, f"s{st}")
        close(loss.cpu().numpy(), ref["losses"], 2e-5, "losses")
        assert correct.cpu().numpy().round().astype(int).tolist() == ref["correct"][:2].tolist()
        got = grads_by_name(plan, gout)
        for k, v in ref.items():
            if k.startswith("grad:") and k[5:] in got:
                nm = k[5:]
                shared = next(p.group for p in plan.params if p.name == nm) == 0
                close(got[nm], v, 2e-4 if shared else 5e-5, f"step {st} grad {nm}")
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


def test_fog_sync_consistency_graph_replay_and_shards(gk):
    """The coupled step at a batch with ragged tiles: CUDA-graph replay equals the eager launch sequence bit for bit, and
    two half-batch shards with GLOBAL denominators / batch size add up to the whole batch (what the data-parallel
    all-reduce relies on)."""
    torch.manual_seed(5)
    B = 301
    def make():
        torch.manual_seed(11)
        return gk.MultiModalMultiTaskModel(21, 6, 6, 6, 426, 16, 8, 128, 3, synchronized_loading=True).cuda()
    crit = lambda: [gk.GCLLoss(cls_num_list=c, m=0.2, s=25.0, noise_mul=0.0) for c in ([50, 30, 20], [45, 33, 22])]
    sk = torch.rand(B, 101, 21, device="cuda"); se = torch.randn(B, 426, 6, device="cuda")
    y = torch.randint(0, 3, (B,), device="cuda")
    m0 = make(); s0 = gk.FusedTrainStep(m0, crit(), cagrad_c=0.1, private_mult=1.0, consistency_lambda=1.0)
    plan = m0.set_window(101).plan()
    g_all = torch.zeros(plan.NP, device="cuda")
    loss_all, _ = s0.step([sk, se], [y, y], grads_out=g_all, update=False)
    loss_all = loss_all.clone()
    # graph replay
    m1 = make(); s1 = gk.FusedTrainStep(m1, crit(), cagrad_c=0.1, private_mult=1.0, consistency_lambda=1.0, use_graph=True)
    m2 = make(); s2 = gk.FusedTrainStep(m2, crit(), cagrad_c=0.1, private_mult=1.0, consistency_lambda=1.0, use_graph=False)
    for _ in range(3):
        s1.step([sk, se], [y, y]); s2.step([sk, se], [y, y])
    assert torch.equal(m1.flat_params(), m2.flat_params())
    # shards: raw gbuf of the two halves adds up to the whole batch's gbuf
    h = 150
    def gbuf_of(lo, hi):
        m = make(); s = gk.FusedTrainStep(m, crit(), cagrad_c=0.1, private_mult=1.0, consistency_lambda=1.0)
        s._step_impl([sk[lo:hi].contiguous(), se[lo:hi].contiguous()], [y[lo:hi].contiguous()] * 2, ys_global=[y, y], part="grads")
        return s._gbuf.clone()
    whole = gbuf_of(0, B); parts = gbuf_of(0, h) + gbuf_of(h, B)
    err = float((whole - parts).abs().max() / whole.abs().max())
    assert err < 2e-6, err


@pytest.mark.parametrize("name", ["fog_sync_gcl", "fog_sync_ce_nc", "fog_async_gcl"])
def test_fog_autograd_path_matches_reference(gk, name):
    """Sync FoG adds the symmetric-KL consistency term (fbg_fog_train.py:81-89,121-124), which couples the two
    streams: it runs through the autograd bridge (forward kernels + per-task backward kernels)."""
    import torch.nn.functional as F
    g = load_golden(name); meta = g["meta"]
    m = fog_model(gk, g)
    crit = fog_criteria(gk, meta)
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    cag = gk.CAGrad(n_tasks=2, device=torch.device("cuda"), c=meta["alpha"], max_norm=1.0)
    for st in range(meta["steps"]):
        i = st % 2
        ls, lt = m(dev(g[f"sk{i}"]), dev(g[f"se{i}"]))
        l1 = crit[0](ls, dev(g[f"ys{i}"])); l2 = crit[1](lt, dev(g[f"yt{i}"]))
        if meta["synchronized"] and meta["wm"] == "gcl":
            kl1 = F.kl_div(F.log_softmax(ls, 1), F.softmax(lt, 1), reduction="batchmean")
            kl2 = F.kl_div(F.log_softmax(lt, 1), F.softmax(ls, 1), reduction="batchmean")
            cons = kl1 + kl2
            l1 = l1 + 0.5 * meta["cons_lambda"] * cons; l2 = l2 + 0.5 * meta["cons_lambda"] * cons
        ref = sub(g, f"s{st}")
        close(torch.stack([l1.detach(), l2.detach()]).cpu().numpy(), ref["losses"], 2e-5, "losses")
        reference_style_step(m, [l1, l2], opt, cag, False)
        named = dict(m.named_parameters())
        shared_ids = {id(p) for p in m.get_shared_parameters()}
        for k, v in ref.items():
            if k.startswith("grad:"):
                nm = k[5:]
                close(named[nm].grad.cpu().numpy(), v, 2e-4 if id(named[nm]) in shared_ids else 5e-5, f"step {st} grad {nm}")
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


def test_cagrad_device_vs_scipy_corpus(gk):
    """gaitk_cagrad (Gram + SLSQP restatement + combine on the device) against what the reference's
    CAGrad.cagrad returned (golden).  Weights always; the combined gradient wherever the reference's own
    formula is well conditioned (lambda = c / ||G w|| explodes when G w ~ 0)."""
    import gait_oracle as O
    g = load_golden("cagrad_corpus")
    cag = gk.CAGrad(n_tasks=3, device=torch.device("cuda"))
    worst_w = worst_g = 0.0; n_g = 0
    for G, gref, wref, (n, alpha) in zip(g["G"], g["g"], g["w"], g["n_alpha"]):
        n = int(n); alpha = float(alpha)
        Gn = G[:, :n].astype(np.float32)
        if (Gn.T @ Gn).max() > 1e6:
            continue                                  # SciPy's LSQ sub-solver breaks down there (cagrad_solver.cuh)
        out, diag = cag.cagrad_device(dev(Gn.T.copy()), alpha=alpha, max_norm=0.0)
        out = out.cpu().numpy() / n; d = diag.cpu().numpy()
        worst_w = max(worst_w, np.abs(d[:n] - wref[:n]).max())
        gw = Gn.astype(np.float64) @ wref[:n]
        if np.linalg.norm(gw) > 1e-2 * np.linalg.norm(Gn):
            worst_g = max(worst_g, np.abs(out - gref).max() / max(np.abs(gref).max(), 1e-30)); n_g += 1
        # exact mode: never a worse objective than SLSQP
        cag.solver = gk._lib.SOLVER_EXACT
        _, d2 = cag.cagrad_device(dev(Gn.T.copy()), alpha=alpha, max_norm=0.0)
        cag.solver = gk._lib.SOLVER_SLSQP
        A = (torch.from_numpy(Gn).t().mm(torch.from_numpy(Gn))).numpy()
        f2 = O.cagrad_objective(A, d2.cpu().numpy()[:n].astype(np.float64), alpha); fr = O.cagrad_objective(A, wref[:n], alpha)
        assert f2 <= fr + 1e-6 * max(1.0, abs(fr))
    assert worst_w < 2e-4 and n_g > 40 and worst_g < 5e-4, (worst_w, worst_g, n_g)


def test_masks_match_reference(gk):
    g = load_golden("wg_masks")
    m = gk.WearGaitThreeModal(synchronized=True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "state0").items()}, strict=True)
    m = m.cuda().eval()
    xs = [dev(g[f"x{j}"]) for j in range(3)]; y = dev(g["y"])
    for i, (nm, mask) in enumerate(zip(g["mask_names"], g["mask_table"])):
        with torch.no_grad():
            # (a) the reference's way: zero-filled tensors through the ordinary forward (weargait_train.py:355-382)
            lg_a = m(*[x if u else torch.zeros_like(x) for x, u in zip(xs, mask)])
            # (b) the mask flag consumed by the kernel (no zero tensors materialised)
            lg_b = m(*xs, enabled=tuple(bool(u) for u in mask))
        for a, b in zip(lg_a, lg_b):
            assert torch.equal(a, b)
        probs = [torch.softmax(l, 1) for l, u in zip(lg_b, mask) if u]
        acc = 100.0 * float(((sum(probs) / len(probs)).argmax(1) == y).sum()) / y.numel()
        assert abs(acc - float(g["acc_sync"][i])) < 1e-9, (nm, acc, g["acc_sync"][i])


def test_fused_step_matches_oracle_random(gk):
    """Seeded random batch at a size the oracle finishes in seconds (B=200, ragged last tile)."""
    import gait_oracle as O
    torch.manual_seed(3)
    m = gk.WearGaitThreeModal(synchronized=False).cuda()
    state = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    B = 201
    xs, y = O.synth_weargait_batch(B, seed=9)
    r = np.random.default_rng(1); ys = [y, r.permutation(y), r.permutation(y)]
    counts = [[40, 90], [55, 60], [20, 30]]
    crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in counts]
    step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0)
    p = O.canonical_params(state, False); bufs = {}
    for it in range(2):
        loss, correct = step.step([dev(x) for x in xs], [dev(v) for v in ys])
        ex = O.weargait_train_step(p, bufs, [torch.from_numpy(x) for x in xs], [torch.from_numpy(v) for v in ys],
                                   synchronized=False, wm="gcl", counts=counts, alpha=0.5)
        close(loss.cpu().numpy(), ex["losses"], 5e-5, "loss")
        ref_correct = [int((l.argmax(1) == torch.from_numpy(v)).sum()) for l, v in zip(ex["logits"], ys)]
        assert correct.cpu().numpy().round().astype(int).tolist() == ref_correct
    bad = []                                               # every parameter, so that a failure names all of them
    for k, v in m.state_dict().items():
        if k in p:
            try:
                close(v.cpu().numpy(), p[k].detach().numpy(), 2e-5, k)
            except AssertionError as e:
                bad.append(str(e))
    assert not bad, "; ".join(bad)


def test_data_path_matches_reference(gk):
    import ctypes as C
    g = load_golden("data_path")
    L = gk.lib()
    st = gk._lib.stream_handle
    sids = [str(s) for s in g["sids"]]; train = [str(s) for s in g["train"]]
    # A2: statistics over the train subjects, then normalisation, on the device
    for mod, D, lo in (("insole", 13, 0), ("imu", 24, 13)):
        acc = torch.zeros(3 * D, dtype=torch.float64, device="cuda")
        for s in train:
            x = dev(g[f"raw/{s}/{mod}"])
            gk._lib.check(L.gaitk_stats_accumulate(x.data_ptr(), x.shape[0], D, acc.data_ptr(), st()))
        mean = torch.empty(D, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
        gk._lib.check(L.gaitk_stats_finalize(acc.data_ptr(), D, mean.data_ptr(), std.data_ptr(), st()))
        np.testing.assert_allclose(mean.cpu().numpy(), g["stat_mean"][lo:lo + D], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(std.cpu().numpy(), g["stat_std"][lo:lo + D], rtol=1e-12)
        # normalise with the REFERENCE statistics -> float32 windows must be bit exact
        mean = dev(g["stat_mean"][lo:lo + D].copy()); std = dev(g["stat_std"][lo:lo + D].copy())
        for split in ("train", "test"):
            keys = [str(k) for k in g[f"{split}_keys/{mod}"]]
            for s in sids:
                mine = [k for k in keys if k.startswith(s + "|")]
                if not mine:
                    continue
                x = dev(g[f"raw/{s}/{mod}"])
                out = torch.empty(x.shape, dtype=torch.float32, device="cuda")
                gk._lib.check(L.gaitk_normalize_frames(x.data_ptr(), x.shape[0], D, mean.data_ptr(), std.data_ptr(),
                                                       out.data_ptr(), st()))
                nwin = int(L.gaitk_window_indices(x.shape[0], 64, 64, None, 0))
                assert nwin == len(mine)
                starts = torch.arange(nwin, dtype=torch.int64, device="cuda") * 64
                win = torch.empty(nwin, 64, D, dtype=torch.float32, device="cuda")
                gk._lib.check(L.gaitk_window_gather(out.data_ptr(), D, starts.data_ptr(), nwin, 64, 1, win.data_ptr(), st()))
                for wid in range(nwin):
                    ref = g[f"{split}_win/{mod}"][keys.index(f"{s}|{mod}|{wid}")].astype(np.float32)
                    assert np.array_equal(win[wid].cpu().numpy(), ref), (s, mod, wid)
    # masked gather writes zeros
    z = torch.ones(2, 64, 13, device="cuda"); fr = torch.rand(200, 13, device="cuda")
    ws = torch.tensor([0, 64], dtype=torch.int64, device="cuda")
    gk._lib.check(L.gaitk_window_gather(fr.data_ptr(), 13, ws.data_ptr(), 2, 64, 0, z.data_ptr(), st()))
    assert float(z.abs().sum()) == 0.0
    # A5: FoG clip preparation, float32 bit exact
    for i in range(3):
        pose = g[f"fog_pose_in{i}"]; sens = g[f"fog_sens_in{i}"]
        o = torch.empty(1, 101, 21, device="cuda"); cs = torch.zeros(1, dtype=torch.int64, device="cuda")
        cl = torch.tensor([pose.shape[0]], dtype=torch.int64, device="cuda")
        gk._lib.check(L.gaitk_fog_prepare_pose(dev(pose).data_ptr(), cs.data_ptr(), cl.data_ptr(), 1, 7, 101, o.data_ptr(), st()))
        assert np.array_equal(o[0].cpu().numpy(), g[f"fog_pose_out{i}"])
        o2 = torch.empty(1, 426, 6, device="cuda"); cl2 = torch.tensor([sens.shape[0]], dtype=torch.int64, device="cuda")
        gk._lib.check(L.gaitk_fog_prepare_sensor(dev(sens).data_ptr(), cs.data_ptr(), cl2.data_ptr(), 1, 6, 426, o2.data_ptr(), st()))
        assert np.array_equal(o2[0].cpu().numpy(), g[f"fog_sens_out{i}"])


def test_fused_gather_equals_dense(gk):
    """win_start path (windows read straight from the frame store) == materialised batch."""
    torch.manual_seed(0)
    m = gk.WearGaitThreeModal().cuda()
    N = 5000
    stores = [torch.rand(N, 2, device="cuda"), torch.randn(N, 13, device="cuda"), torch.randn(N, 24, device="cuda")]
    B = 37
    ws = (torch.randperm(N // 64, device="cuda")[:B] * 64).to(torch.int64)
    dense = [torch.stack([s[int(w):int(w) + 64] for w in ws.cpu()]) for s in stores]
    m.set_window(64)
    a = m.plan().forward(m.flat_params(), dense)
    b = m.plan().forward(m.flat_params(), stores, win_start=[ws, ws, ws])
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_data_parallel_shards_add_up(gk):
    """Two half-batches with global denominators reduce (sum of gbuf) to the full-batch step: the
    property the NCCL all-reduce relies on."""
    import gait_oracle as O
    torch.manual_seed(1)
    m = gk.WearGaitThreeModal().cuda()
    B = 256
    xs, y = O.synth_weargait_batch(B, seed=4)
    xs = [dev(x) for x in xs]; y = dev(y)
    w = torch.tensor([1.7, 0.3], device="cuda")
    crit = [gk.GCLLoss(cls_num_list=[30, 70], m=0.2, s=25, noise_mul=0.0, weight=w) for _ in range(3)]
    step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, process_group=False)
    plan = m.set_window(64).plan()
    g_full = torch.zeros(plan.NP, device="cuda")
    step.step(xs, [y] * 3, update=False, grads_out=g_full)
    full = step._gbuf.clone()
    acc = torch.zeros_like(full)
    for h in range(2):
        sl = slice(h * B // 2, (h + 1) * B // 2)
        step.step([x[sl].contiguous() for x in xs], [y[sl].contiguous()] * 3, ys_global=[y] * 3, update=False,
                  grads_out=torch.zeros(plan.NP, device="cuda"))
        acc += step._gbuf
    close(acc.cpu().numpy(), full.cpu().numpy(), 2e-5, "gbuf")


# ------------------------------------------------------------------------------------------------ tensor-core path
def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("name", ["wg_sync_gcl", "wg_async_ce", "wg_sync_classwt_nc", "wg_sync_norm"])
def test_weargait_tf32_tensor_core_path(gk, name):
    """tcgen05 (forward / dgrad) + mma.sync (wgrad) tf32 kernels against the reference goldens: north_star's
    reduced-precision bar is 1e-3 relative (norm-wise) on outputs and gradients."""
    g = load_golden(name); meta = g["meta"]
    m = wg_model(gk, g); m.compute_dtype = gk.DTYPE_TF32
    xs = [dev(g[f"x0_{j}"]) for j in range(3)]; ys = [dev(g[f"y0_{j}"]) for j in range(3)]
    with torch.no_grad():
        out = m(*xs)
    ref = sub(g, "s0")
    assert relerr(torch.stack(out).cpu().numpy(), ref["logits"]) < 1e-3
    step = gk.FusedTrainStep(m, wg_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=2.0, dtype=gk.DTYPE_TF32)
    plan = m.plan()
    gout = torch.zeros(plan.NP, device="cuda")
    loss, correct = step.step(xs, ys, grads_out=gout, update=False)
    assert relerr(loss.cpu().numpy(), ref["losses"]) < 1e-3
    G = step._gbuf[:3 * plan.P].view(3, plan.P).t().cpu().numpy()
    # B = 8 here: tf32 rounding noise does not average out (4e-3 at B=8 -> 2.4e-4 at B=32768)
    assert relerr(G, ref["G"]) < 5e-2, relerr(G, ref["G"])
    got = grads_by_name(plan, gout)
    worst = 0.0
    for k, v in ref.items():
        if k.startswith("grad:") and k[5:] in got:
            grp = next(p.group for p in plan.params if p.name == k[5:])
            if grp > 0:
                e = relerr(got[k[5:]], v); worst = max(worst, e)
                assert e < 0.15, (k, e)       # small-norm bias gradients at B=8: cancellation amplifies the noise
    # fp32 and tf32 paths agree with each other the same way
    step32 = gk.FusedTrainStep(m, wg_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=2.0, dtype=gk.DTYPE_F32)
    g32 = torch.zeros(plan.NP, device="cuda")
    step32.step(xs, ys, grads_out=g32, update=False)
    G32 = step32._gbuf[:3 * plan.P].view(3, plan.P).t().cpu().numpy()
    assert relerr(G, G32) < 5e-2


def test_tf32_path_large_batch_vs_fp32_path(gk):
    """Full-size property check: both arithmetic paths on B=4097 windows (ragged tile, many CTAs)."""
    import gait_oracle as O
    torch.manual_seed(5)
    m = gk.WearGaitThreeModal().cuda()
    B = 4097
    xs, y = O.synth_weargait_batch(B, seed=11)
    xs = [dev(x) for x in xs]; y = dev(y)
    crit = [gk.GCLLoss(cls_num_list=[400, 600], m=0.2, s=25, noise_mul=0.0) for _ in range(3)]
    plan = m.set_window(64).plan()
    res = {}
    for dt in (gk.DTYPE_F32, gk.DTYPE_TF32):
        st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=dt, process_group=False)
        gout = torch.zeros(plan.NP, device="cuda")
        loss, correct = st.step(xs, [y, y, y], grads_out=gout, update=False)
        res[dt] = (loss.cpu().numpy(), correct.cpu().numpy(), st._gbuf[:3 * plan.P].cpu().numpy(), gout.cpu().numpy())
    a, b = res[gk.DTYPE_F32], res[gk.DTYPE_TF32]
    assert relerr(b[0], a[0]) < 1e-3
    assert np.abs(b[1] - a[1]).max() <= 0.002 * B           # argmax flips only on near-ties
    assert relerr(b[2], a[2]) < 1e-3, relerr(b[2], a[2])          # shared-gradient matrix G
    assert relerr(b[3], a[3]) < 4e-3, relerr(b[3], a[3])          # all final gradients (private ones dominate)


@pytest.mark.parametrize("T", [32, 128])
def test_tf32_path_other_window_lengths(gk, T):
    """Window lengths other than the default use the runtime-geometry instantiation of the tensor-core kernel
    (T = 32 -> 4 windows per 128-row tile, T = 128 -> 1): same properties as the default geometry."""
    import gait_oracle as O
    torch.manual_seed(6)
    m = gk.WearGaitThreeModal().cuda()
    B = 1031
    xs, y = O.synth_weargait_batch(B, T=T, seed=12)
    xs = [dev(x) for x in xs]; y = dev(y)
    crit = [gk.GCLLoss(cls_num_list=[400, 600], m=0.2, s=25, noise_mul=0.0) for _ in range(3)]
    plan = m.set_window(T).plan()
    res = {}
    for dt in (gk.DTYPE_F32, gk.DTYPE_TF32):
        st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=dt, process_group=False)
        gout = torch.zeros(plan.NP, device="cuda")
        loss, correct = st.step(xs, [y, y, y], grads_out=gout, update=False)
        res[dt] = (loss.cpu().numpy(), correct.cpu().numpy(), st._gbuf[:3 * plan.P].cpu().numpy(), gout.cpu().numpy())
    a, b = res[gk.DTYPE_F32], res[gk.DTYPE_TF32]
    assert relerr(b[0], a[0]) < 1e-3
    assert np.abs(b[1] - a[1]).max() <= 0.004 * B + 1
    assert relerr(b[2], a[2]) < 2e-3, relerr(b[2], a[2])
    assert relerr(b[3], a[3]) < 8e-3, relerr(b[3], a[3])
    # and the fp32 path itself against the CPU oracle at this window length
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items() if not k.startswith("_shared") and not k.startswith("head_i") and not k.startswith("head_m")}
    lg = O.weargait_forward(O._HeadAlias(p), *[x.cpu() for x in xs])
    with torch.no_grad():
        out = m(*xs)
    for u, v in zip(out, lg):
        close(u.cpu().numpy(), v.detach().numpy(), 2e-5, f"logits T={T}")


def test_end_to_end_training_matches_oracle(gk):
    """30 training steps + evaluation under the 7 modality masks: fused CUDA path vs the CPU oracle on the same
    data / seeds.  north_star: end-to-end metrics within 0.5 points."""
    import gait_oracle as O
    torch.manual_seed(11)
    m = gk.WearGaitThreeModal().cuda()
    state = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    p = O.canonical_params(state, True); bufs = {}
    B, steps = 64, 30
    counts = [[130, 190]] * 3
    crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in counts]
    st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, lr=1e-2, process_group=False)
    tl_gpu, tl_cpu = [], []
    for it in range(steps):
        xs, y = O.synth_weargait_batch(B, seed=1000 + it % 5)
        loss, _ = st.step([dev(x) for x in xs], [dev(y)] * 3)
        ex = O.weargait_train_step(p, bufs, [torch.from_numpy(x) for x in xs], [torch.from_numpy(y)] * 3, synchronized=True,
                                   wm="gcl", counts=counts, alpha=0.5, lr=1e-2)
        tl_gpu.append(loss.cpu().numpy()); tl_cpu.append(np.array(ex["losses"]))
    close(np.array(tl_gpu), np.array(tl_cpu), 2e-3, "loss trajectory")
    # held-out evaluation under every mask (weargait_train.eval_with_mask, sync ensemble)
    xs, y = O.synth_weargait_batch(512, seed=77)
    xd = [dev(x) for x in xs]; yd = dev(y); xt = [torch.from_numpy(x) for x in xs]; yt = torch.from_numpy(y)
    m.eval()
    for name, mask in O.MASK_COMBOS.items():
        with torch.no_grad():
            lg = m(*xd, enabled=mask)
        probs = [torch.softmax(l, 1) for l, u in zip(lg, mask) if u]
        acc_gpu = 100.0 * float(((sum(probs) / len(probs)).argmax(1) == yd).sum()) / yd.numel()
        c, n = O.eval_mask_sync(p, xt, yt, mask)
        assert abs(acc_gpu - 100.0 * c / n) <= 0.5, (name, acc_gpu, 100.0 * c / n)


def test_end_to_end_training_tf32_close_to_fp32(gk):
    """Same protocol, tensor-core path vs fp32 path: accuracies within 0.5 points after 30 steps."""
    import gait_oracle as O
    accs = {}
    for dt in (gk.DTYPE_F32, gk.DTYPE_TF32):
        torch.manual_seed(11)
        m = gk.WearGaitThreeModal().cuda(); m.compute_dtype = dt
        crit = [gk.GCLLoss(cls_num_list=[130, 190], m=0.2, s=25, noise_mul=0.0) for _ in range(3)]
        st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, lr=1e-2, process_group=False, dtype=dt)
        for it in range(30):
            xs, y = O.synth_weargait_batch(256, seed=1000 + it % 5)
            st.step([dev(x) for x in xs], [dev(y)] * 3)
        xs, y = O.synth_weargait_batch(2048, seed=77)
        with torch.no_grad():
            lg = m(*[dev(x) for x in xs])
        p = sum(torch.softmax(l, 1) for l in lg) / 3
        accs[dt] = 100.0 * float((p.argmax(1) == dev(y)).sum()) / len(y)
    assert abs(accs[gk.DTYPE_F32] - accs[gk.DTYPE_TF32]) <= 0.5, accs


@pytest.mark.parametrize("sync", [True, False])
def test_relaxed_input_training_with_dropped_streams(gk, sync):
    """SURVEY 8(d) cfg 3(ii): a per-batch modality mask during training -- masked streams zero-filled
    (_maybe_zero, weargait_train.py:355-358) and their losses dropped from the CAGrad task list
    (step_cagrad_three :200-203), n_tasks in {1,2,3}.  Fused path vs oracle over a sequence of masks."""
    import gait_oracle as O
    torch.manual_seed(21)
    m = gk.WearGaitThreeModal(synchronized=sync).cuda()
    state = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    p = O.canonical_params(state, sync); bufs = {}
    counts = [[30, 70], [45, 55], [20, 80]]
    crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in counts]
    st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, process_group=False)
    masks = [(True, True, True), (True, False, True), (False, True, False), (False, True, True), (True, True, False), (True, False, False)]
    for it, mask in enumerate(masks):
        xs, y = O.synth_weargait_batch(48, seed=300 + it)
        r = np.random.default_rng(it)
        ys = [y, y, y] if sync else [y, r.permutation(y), r.permutation(y)]
        loss, correct = st.step([dev(x) for x in xs], [dev(v) for v in ys], enabled=mask, tasks=mask)
        ex = O.weargait_train_step(p, bufs, [torch.from_numpy(x) for x in xs], [torch.from_numpy(v) for v in ys],
                                   synchronized=sync, wm="gcl", counts=counts, alpha=0.5, tasks=mask)
        got = loss.cpu().numpy()
        for i in range(3):
            if mask[i]:
                assert abs(got[i] - ex["losses"][i]) <= 5e-5 * max(1.0, abs(ex["losses"][i])), (it, i, got, ex["losses"])
        for k, v in m.state_dict().items():
            if k in p:
                close(v.cpu().numpy(), p[k].detach().numpy(), 2e-5, f"step {it} mask {mask} {k}")


@pytest.mark.parametrize("name", ["wg_sync_gcl_dropped", "wg_async_gcl_dropped"])
def test_dropped_stream_training_matches_reference_goldens(gk, name):
    """The same relaxed-input case against what the REFERENCE produced (step_cagrad_three with None losses,
    CAGrad(n_tasks=k)): parameters after every step, live losses, simplex weights."""
    g = load_golden(name); meta = g["meta"]
    m = wg_model(gk, g)
    step = gk.FusedTrainStep(m, wg_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=2.0, process_group=False)
    for st in range(meta["steps"]):
        i = st % 2; mask = tuple(bool(u) for u in meta["masks"][st])
        xs = [dev(g[f"x{i}_{j}"]) for j in range(3)]; ys = [dev(g[f"y{i}_{j}"]) for j in range(3)]
        loss, _ = step.step(xs, ys, enabled=mask, tasks=mask)
        ref = sub(g, f"s{st}")
        live = [j for j in range(3) if mask[j]]
        close(loss.cpu().numpy()[live], ref["losses"][live], 2e-5, "live losses")
        d = step.diag().cpu().numpy()
        assert np.abs(d[:3][live] - ref["w"]).max() < 2e-4, (d[:3], ref["w"])
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} mask {mask} param {k[6:]}")


def test_single_modality_paths_match_reference(gk):
    """SkelModalityModel / SensorModalityModel (feature_encoder.py:268-344) and the --single_mod branch of the
    WearGait model driven exactly as weargait_train._single_logits_and_labels does (sub-module calls)."""
    g = load_golden("single_modality")
    prm = dict(skeleton_input_dim=21, skeleton_output_dim=6, sensor_in_channels=6, sensor_out_channels=6, sensor_length=426,
               shared_out_channels=16, backbone_dim=8, taskhead_input_dim=128, num_classes=3)
    y = dev(g["y"])
    for mod, xk in (("skeleton", "sk"), ("sensor", "se")):
        if mod == "skeleton":
            m = gk.SkelModalityModel(21, 6, 6, 16, 8, 128, 3)
        else:
            m = gk.SensorModalityModel(6, 6, 426, 16, 8, 128, 3)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(sub(g, mod), "state0").items()}, strict=True)
        m = m.cuda()
        lg = m(dev(g[xk]))
        close(lg.detach().cpu().numpy(), g[f"{mod}/logits"], 2e-5, f"{mod} logits")
        loss = gk.CrossEntropyLoss()(lg, y); loss.backward()
        close(float(loss.detach()), g[f"{mod}/loss"], 2e-5, "loss")
        named = dict(m.named_parameters())
        for k, v in sub(sub(g, mod), "grad").items():
            close(named[k].grad.cpu().numpy(), v, 5e-5, f"{mod} grad {k}")
    yw = dev(g["wg/y"])
    for j, mod in enumerate(("walkway", "insole", "imu")):
        m = gk.WearGaitThreeModal(synchronized=True)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(sub(g, "wg"), "state0").items()}, strict=True)
        m = m.cuda()
        x = dev(g[f"wg/x{j}"])
        enc = {"walkway": m.enc_w, "insole": m.enc_i, "imu": m.enc_m}[mod]
        head = {"walkway": m.head_w, "insole": m.head_i, "imu": m.head_m}[mod]
        lg = head(m.backbone(enc(x)).flatten(1))                       # weargait_train.py:262-270
        close(lg.detach().cpu().numpy(), g[f"wg/{mod}/logits"], 2e-5, f"wg {mod} logits")
        loss = gk.CrossEntropyLoss()(lg, yw); loss.backward()
        named = dict(m.named_parameters())
        for k, v in sub(sub(sub(g, "wg"), mod), "grad").items():
            if k in named:
                close(named[k].grad.cpu().numpy(), v, 5e-5, f"wg {mod} grad {k}")
        others = [p for n_, p in named.items() if n_.startswith("enc_") and not n_.startswith({"walkway": "enc_w", "insole": "enc_i", "imu": "enc_m"}[mod])]
        assert all(p.grad is None for p in others)


# ---------------------------------------------------------------------------------------------------------------
# fusion baselines (weargait_train.py --baseline late_fusion | shared_latent): plain mean of the three CE losses
BL_CASES = ["bl_late_sync", "bl_late_async", "bl_shared_latent_sync", "bl_shared_latent_async",
            "bl_early_sync", "bl_early_async", "bl_xattn_sync", "bl_xattn_async"]       # the last four: staged execution (staged.py)


def bl_model(gk, g):
    meta = g["meta"]
    kw = dict(enc_out_ch=12, backbone_dim=8, shared_out_ch=16, num_classes=2, synchronized=meta["synchronized"])
    if meta["baseline"] == "shared_latent":
        m = gk.SharedLatent3(proj_ch=16, **kw)
    elif meta["baseline"] == "early_fusion":
        m = gk.EarlyFusion3(**kw)
    elif meta["baseline"] == "cheap_xattn":
        m = gk.CheapXAttn3(**kw)
    else:
        m = gk.LateFusion3(**kw)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "state0").items()}, strict=True)
    return m.cuda()


@pytest.mark.parametrize("name", BL_CASES)
def test_fusion_baseline_autograd_training_matches_reference(gk, name):
    g = load_golden(name); meta = g["meta"]
    m = bl_model(gk, g)
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    ce = torch.nn.CrossEntropyLoss()
    for st in range(meta["steps"]):
        xs = [dev(g[f"x{st % 2}_{j}"]) for j in range(3)]; ys = [dev(g[f"s{st}/y{j}"]) for j in range(3)]
        m.train()
        lg = m(*xs)
        ref = sub(g, f"s{st}")
        close(torch.stack([l.detach() for l in lg]).cpu().numpy(), ref["logits"], 2e-5, "logits")
        L = [ce(l, y) for l, y in zip(lg, ys)]
        close(torch.stack([l.detach() for l in L]).cpu().numpy(), ref["losses"], 2e-5, "losses")
        opt.zero_grad(set_to_none=True)
        torch.stack(L).mean().backward()                      # step_cagrad_three with cagrad=None (:244-248)
        named = dict(m.named_parameters())
        n_checked = 0
        for k, v in ref.items():
            if k.startswith("grad:"):
                close(named[k[5:]].grad.cpu().numpy(), v, 5e-5, f"step {st} grad {k[5:]}"); n_checked += 1
        assert n_checked >= 14
        g1 = named["enc_i.ln1.weight"].grad                  # constructed, never applied (weargait_encoders.py:81,93-101)
        assert g1 is None or float(g1.abs().max()) == 0.0
        opt.step()
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


@pytest.mark.parametrize("name", ["bl_shared_latent_sync", "bl_shared_latent_async", "bl_late_async"])
def test_fusion_baseline_fused_step_matches_reference(gk, name):
    """The plain averaged step is the fused step in GAITK_SOLVER_MEAN mode (shared grad = mean of the task rows), no
    clipping and private gradients scaled by 1/3."""
    g = load_golden(name); meta = g["meta"]
    m = bl_model(gk, g)
    crit = [gk.CrossEntropyLoss() for _ in range(3)]
    step = gk.FusedTrainStep(m, crit, cagrad_c=0.0, max_norm=0.0, private_mult=1.0 / 3.0, solver=gk.SOLVER_MEAN)
    for st in range(meta["steps"]):
        xs = [dev(g[f"x{st % 2}_{j}"]) for j in range(3)]; ys = [dev(g[f"s{st}/y{j}"]) for j in range(3)]
        loss, _ = step.step(xs, ys)
        ref = sub(g, f"s{st}")
        close(loss.cpu().numpy()[:3], ref["losses"], 2e-5, "losses")
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                close(sd[k[6:]].cpu().numpy(), v, 1e-5, f"step {st} param {k[6:]}")


# ---------------------------------------------------------------------------------------------------------------
# device-resident fold preparation + loaders (package dataloader_weargait.py vs the reference's prepare_split/loaders)
def _golden_frames(g):
    sids = [str(s) for s in g["sids"]]
    return sids, {s: {m: g[f"raw/{s}/{m}"] for m in ("walkway", "insole", "imu")} for s in sids}


def test_device_prepare_split_matches_reference(gk):
    dl = gk.dataloader_weargait
    g = load_golden("data_path")
    sids, frames = _golden_frames(g)
    train = [str(s) for s in g["train"]]; test = [str(s) for s in g["test"]]
    prep = dl.prepare_split(train, test, frames=frames, win=64, hop=64)
    mean = torch.cat([prep["stats"]["insole"][0], prep["stats"]["imu"][0]]).cpu().numpy()
    std = torch.cat([prep["stats"]["insole"][1], prep["stats"]["imu"][1]]).cpu().numpy()
    np.testing.assert_allclose(mean, g["stat_mean"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(std, g["stat_std"], rtol=1e-12)
    n_checked = 0
    for split in ("train", "test"):
        assert [[p[0].split("|")[0], p[0].split("|")[2]] for p in prep[f"{split}_sync"]] == [[str(a), str(b)] for a, b in g[f"{split}_sync"]]
        for m in ("walkway", "insole", "imu"):
            st = prep[f"{split}_stores"][m]
            keys = [str(k) for k in g[f"{split}_keys/{m}"]]
            assert sorted(st.keys()) == keys
            if not keys:
                continue
            pos = torch.tensor([st.starts[st.position(k)] for k in keys], dtype=torch.int64, device="cuda")
            win = st.gather(pos).cpu().numpy()
            ref = g[f"{split}_win/{m}"].astype(np.float32)
            # statistics agree to ~1e-13 relative, so a float32 z-score may differ from the reference's by one ulp on a
            # rounding tie; everything else is bit exact
            same = win == ref
            assert same.mean() > 0.9999 and np.allclose(win, ref, rtol=2e-7, atol=1e-12), (split, m, same.mean())
            n_checked += len(keys)
    assert n_checked > 50


def test_device_loaders_yield_reference_batches(gk):
    dl = gk.dataloader_weargait
    g = load_golden("data_path")
    sids, frames = _golden_frames(g)
    prep = dl.prepare_split([str(s) for s in g["train"]], [str(s) for s in g["test"]], frames=frames)
    s2l = dl.build_subj2label(sids[:3], sids[3:])
    tr, te = dl.make_sync_loaders(prep, s2l, batch_size=4, num_workers=0, seed=43, with_keys=True)
    last = None; ks = []
    for b in tr:
        last = b; ks += [k[0].split("|")[0] + "|" + k[0].split("|")[2] for k in b["keys"]]
    assert ks == [str(k) for k in g["loader_sync/train_ep0_keys"]]
    for j in range(3):
        np.testing.assert_allclose(last["xs"][j].cpu().numpy(), g[f"loader_sync/last_batch_x{j}"], rtol=2e-7, atol=1e-12)
    assert last["y"].dtype == torch.int64 and last["xs"][1].shape == (4, 64, 13)
    tr, te = dl.make_async_loaders(prep, s2l, batch_size=4, num_workers=0, seed=43, with_keys=True)
    tr.dataset.reseed(44)
    for b in tr:
        last = b
    for m in ("walkway", "insole", "imu"):
        np.testing.assert_allclose(last[m].cpu().numpy(), g[f"loader_async/last_batch/{m}"], rtol=2e-7, atol=1e-12)
        assert last["y"][m].shape == (1,)


def test_resident_fold_training_equals_dense_batch_training(gk):
    """An epoch trained from index batches (windows read from the resident stores by the stream kernels) ends with the
    same parameters as the same epoch trained from the dense batches the loaders hand to an unmodified trainer loop."""
    dl = gk.dataloader_weargait
    rng = np.random.default_rng(3)
    sids = [f"pd{i}" for i in range(4)] + [f"hc{i}" for i in range(4)]
    frames = {}
    for i, s in enumerate(sids):
        n = int(rng.integers(700, 1500))
        frames[s] = {"walkway": rng.random((n, 2)), "insole": rng.standard_normal((n - int(rng.integers(0, 70)), 13)) * 3 + 1,
                     "imu": rng.standard_normal((n - int(rng.integers(0, 70)), 24)) * (2 if i < 4 else 1)}
    frames[sids[1]]["insole"][5:40, 3] = np.nan
    s2l = dl.build_subj2label(sids[:4], sids[4:])
    prep = dl.prepare_split(sids[:3] + sids[4:7], [sids[3], sids[7]], frames=frames)
    params = []
    for mode in ("dense", "index"):
        torch.manual_seed(0)
        m = gk.WearGaitThreeModal().cuda().set_window(64)      # index batches carry no window length
        crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
        step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=gk.DTYPE_F32)
        tr, _ = dl.make_sync_loaders(prep, s2l, batch_size=16, seed=43)
        n = 0
        for ep in range(2):
            if mode == "dense":
                for b in tr:
                    step.step(b["xs"], [b["y"]] * 3); n += 1
            else:
                for ib in tr.index_batches():
                    step.step(ib.frames, ib.ys, win_start=ib.win_start); n += 1
        assert n == 2 * len(tr) and n >= 10
        params.append(m.flat_params().detach().clone())
    assert torch.equal(params[0], params[1])


# ---------------------------------------------------------------------------------------------------------------
# relaxed-input evaluation: seven masks from one forward pass (package evaluation.py vs eval_with_mask goldens)
def test_all_masks_from_one_pass_match_reference(gk):
    g = load_golden("wg_masks")
    state = {k: torch.from_numpy(v) for k, v in sub(g, "state0").items()}
    xs = [torch.from_numpy(g[f"x{j}"]) for j in range(3)]; y = torch.from_numpy(g["y"])
    names = [str(n) for n in g["mask_names"]]
    assert names == list(gk.MASK_COMBOS) and [tuple(bool(u) for u in r) for r in g["mask_table"]] == list(gk.MASK_COMBOS.values())
    m = gk.WearGaitThreeModal(synchronized=True); m.load_state_dict(state, strict=True); m = m.cuda()
    batch = {"xs": xs, "y": y}
    table = gk.eval_all_masks(m, [batch], False)
    for i, nm in enumerate(names):
        assert abs(table[nm] - float(g["acc_sync"][i])) < 1e-9, (nm, table[nm], g["acc_sync"][i])
        assert abs(gk.eval_with_mask(m, [batch], False, nm) - float(g["acc_sync"][i])) < 1e-9
    am = gk.WearGaitThreeModal(synchronized=False)
    am.load_state_dict({k: v for k, v in state.items() if not k.startswith("_shared")}, strict=True); am = am.cuda()
    abatch = {"walkway": xs[0], "insole": xs[1], "imu": xs[2], "y": {"walkway": y, "insole": y, "imu": y}}
    atab = gk.eval_all_masks(am, [abatch, abatch], True)
    for i, nm in enumerate(names):
        assert abs(atab[nm]["macro_enabled"] - float(g["acc_async"][i])) < 1e-9, (nm, atab[nm])
        assert set(atab[nm]) == {s for s, u in zip(("walkway", "insole", "imu"), gk.MASK_COMBOS[nm]) if u} | {"macro_enabled"}
    # the kernel's ensembles against torch ops on the same logits, larger batch, three classes
    torch.manual_seed(0)
    lg = [torch.randn(5000, 3, device="cuda") * 2 for _ in range(3)]; yy = torch.randint(0, 3, (5000,), device="cuda")
    cnt = torch.zeros(10, dtype=torch.int32, device="cuda")
    import ctypes as C
    arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    gk._lib.check(gk.lib().gaitk_mask_eval(arr(lg), arr([yy] * 3), 5000, 3, cnt.data_ptr(), gk._lib.stream_handle()))
    cnt = cnt.cpu().numpy()
    for i, mask in enumerate(gk.MASK_COMBOS.values()):
        probs = [torch.softmax(l, 1) for l, u in zip(lg, mask) if u]
        assert int(((sum(probs) / len(probs)).argmax(1) == yy).sum()) == cnt[i]
    for s in range(3):
        assert int((lg[s].argmax(1) == yy).sum()) == cnt[7 + s]


def test_pipelined_resident_steps_equal_synchronous_ones(gk):
    """FusedTrainStep.step_indices_async (two device slots, copy stream, results read one step late) against the blocking
    step_indices on the same stores, indices and labels: identical per-step (loss, correct) and identical parameters."""
    import gait_oracle as O
    B, T, n_steps = 96, 64, 5
    xs, _ = O.synth_weargait_batch(4 * B, seed=5)
    stores = [dev(x.reshape(-1, x.shape[-1])) for x in xs]
    rng = np.random.default_rng(7)
    idx = [torch.from_numpy((rng.permutation(4 * B)[:B] * T).astype(np.int64)).pin_memory() for _ in range(n_steps)]
    ys = [torch.from_numpy(rng.integers(0, 2, B).astype(np.int64)).pin_memory() for _ in range(n_steps)]
    outs = {}
    params = {}
    for mode in ("sync", "async"):
        torch.manual_seed(11)
        m = gk.WearGaitThreeModal().cuda(); m.set_window(T)
        crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
        step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=gk.DTYPE_F32)
        got = []
        if mode == "sync":
            for i in range(n_steps):
                got.append(step.step_indices(stores, [idx[i]] * 3, [ys[i]] * 3).clone())
        else:
            pending = None
            for i in range(n_steps):
                h = step.step_indices_async(stores, [idx[i]] * 3, [ys[i]] * 3, slot=i % 2)
                if pending is not None:
                    got.append(pending.result())
                pending = h
            got.append(pending.result())
        outs[mode] = torch.stack(got); params[mode] = m.flat_params().clone()
    assert torch.equal(outs["sync"], outs["async"])
    assert torch.equal(params["sync"], params["async"])


def test_eval_one_epoch_on_resident_loader_matches_dense_torch(gk):
    dl = gk.dataloader_weargait
    rng = np.random.default_rng(5)
    sids = [f"pd{i}" for i in range(3)] + [f"hc{i}" for i in range(3)]
    frames = {s: {"walkway": rng.random((400, 2)), "insole": rng.standard_normal((380, 13)), "imu": rng.standard_normal((390, 24)) * (2 if i < 3 else 1)}
              for i, s in enumerate(sids)}
    s2l = dl.build_subj2label(sids[:3], sids[3:])
    prep = dl.prepare_split(sids[:2] + sids[3:5], [sids[2], sids[5]], frames=frames)
    torch.manual_seed(1)
    m = gk.WearGaitThreeModal().cuda()
    crit = [gk.CrossEntropyLoss() for _ in range(3)]
    _, te = dl.make_sync_loaders(prep, s2l, batch_size=4, seed=1)
    loss, acc, ens = gk.eval_one_epoch(m, te, False, crit)
    # the reference's loop (weargait_train.py:322-350) with torch ops on the dense batches of the same loader
    _, te2 = dl.make_sync_loaders(prep, s2l, batch_size=4, seed=1)
    n = 0; ls = np.zeros(3); ac = np.zeros(3); corr = tot = 0
    with torch.no_grad():
        for b in te2:
            lg = m(*b["xs"]); y = b["y"]
            ls += np.array([float(torch.nn.functional.cross_entropy(l, y)) for l in lg])
            ac += np.array([(l.argmax(1) == y).float().mean().item() * 100 for l in lg]); n += 1
            p = sum(torch.softmax(l, 1) for l in lg) / 3.0
            corr += int((p.argmax(1) == y).sum()); tot += y.numel()
    np.testing.assert_allclose(loss, ls / n, rtol=2e-6)
    np.testing.assert_allclose(acc, ac / n, rtol=1e-12)
    assert abs(ens - 100.0 * corr / tot) < 1e-9


def test_peer_memory_exchange_matches_nccl_on_two_gpus(gk):
    """gaitk_p2p_allreduce (the data-parallel exchange as a kernel over NVLink peer memory) against the NCCL path:
    launched through torchrun when two devices are visible (tests/dist_p2p_check.py), skipped on a single-GPU box."""
    import subprocess, sys
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", str(ROOT / "tests" / "dist_p2p_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "P2P_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_peer_exchange_timeout_is_sticky_and_skips_the_update(gk, monkeypatch):
    """a peer that never publishes its step must (a) not hang the device, (b) leave a STICKY failure status
    that the update kernel does not overwrite, (c) keep parameters and momentum untouched.  One GPU is enough: the
    'peer' of this world-2 exchange is a flag word nobody writes (no second kernel waits on anything)."""
    import ctypes as C
    L = gk.lib(); lib_ = gk._lib
    monkeypatch.setenv("GAITK_P2P_TIMEOUT_CYCLES", str(1 << 22))          # ~2 ms of SM clocks
    m = gk.WearGaitThreeModal().cuda(); m.set_window(64)
    plan = m.plan(); flat = m.flat_params(); before = flat.clone()
    n = plan.gbuf_floats
    gb = [torch.randn(n, device="cuda") * 1e-3 for _ in range(2)]
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    peer_g = torch.tensor([g.data_ptr() for g in gb], dtype=torch.int64, device="cuda")
    peer_f = torch.tensor([flags.data_ptr(), flags.data_ptr() + 4], dtype=torch.int64, device="cuda")
    counter = torch.zeros(1, dtype=torch.int32, device="cuda"); gsum = torch.zeros(n, device="cuda")
    diag = torch.zeros(lib_.DIAG_FLOATS, device="cuda"); mom = torch.zeros(plan.NP, device="cuda")
    st = lib_.stream_handle()
    lib_.check(L.gaitk_p2p_allreduce(plan.handle, peer_g.data_ptr(), peer_f.data_ptr(), counter.data_ptr(), 0, 2, gsum.data_ptr(),
                                     diag.data_ptr(), st), "p2p")
    lib_.check(L.gaitk_step_update(plan.handle, flat.data_ptr(), mom.data_ptr(), gsum.data_ptr(), 0b111, 0.5, 1.0, 1e-3, 0.9, 1e-4,
                                   None, diag.data_ptr(), lib_.SOLVER_SLSQP | lib_.SOLVER_FLAG_CHECK_EXCHANGE, st), "update")
    torch.cuda.synchronize()
    assert float(diag[lib_.DIAG_EXCHANGE]) == -1.0
    assert torch.equal(flat, before) and float(mom.abs().max()) == 0.0
    # the same update with a healthy status does move the parameters
    diag.zero_()
    lib_.check(L.gaitk_step_update(plan.handle, flat.data_ptr(), mom.data_ptr(), gsum.data_ptr(), 0b111, 0.5, 1.0, 1e-3, 0.9, 1e-4,
                                   None, diag.data_ptr(), lib_.SOLVER_SLSQP | lib_.SOLVER_FLAG_CHECK_EXCHANGE, st), "update")
    torch.cuda.synchronize()
    assert not torch.equal(flat, before)


# ------------------------------------------------------------------------------------------------ split-bf16 (bf16x3) path
# (normalising / cosine heads -- wg_sync_classwt_nc, wg_sync_norm -- are served by the fp32 and tf32 kernels: the ws kernel's
# head is the plain linear head the trainers use by default and refuses anything else with GAITK_E_DTYPE)
BF16X3_CASES = ["wg_sync_gcl", "wg_sync_gcl_drw", "wg_async_ce", "wg_async_gcl"]


def test_bf16x3_refuses_heads_it_does_not_implement(gk):
    m = gk.WearGaitThreeModal(use_norm=True).cuda(); m.compute_dtype = gk.DTYPE_BF16X3
    with pytest.raises(gk.GaitkError, match="GAITK_E_DTYPE"):
        m(torch.zeros(4, 64, 2, device="cuda"), torch.zeros(4, 64, 13, device="cuda"), torch.zeros(4, 64, 24, device="cuda"))


@pytest.mark.parametrize("name", BF16X3_CASES)
def test_weargait_bf16x3_path_matches_reference_goldens(gk, name):
    """The benched arithmetic (warp-specialised all-tcgen05 kernel, split-bf16 operands) DIRECTLY against the reference
    goldens at B = 8: north_star's reduced-precision bar, 1e-3 relative (norm-wise), on logits, losses, the per-task shared
    gradient matrix G, EVERY private gradient and the parameters after each of the three steps."""
    g = load_golden(name); meta = g["meta"]
    m = wg_model(gk, g); m.compute_dtype = gk.DTYPE_BF16X3
    step = gk.FusedTrainStep(m, wg_criteria(gk, meta), cagrad_c=meta["alpha"], private_mult=2.0, dtype=gk.DTYPE_BF16X3)
    for st in range(meta["steps"]):
        i = st % 2
        xs = [dev(g[f"x{i}_{j}"]) for j in range(3)]; ys = [dev(g[f"y{i}_{j}"]) for j in range(3)]
        plan = m.set_window(xs[0].shape[1]).plan()
        ref = sub(g, f"s{st}")
        if st == 0:
            with torch.no_grad():
                out = m(*xs)                                  # forward-only mode of the kernel
            assert relerr(torch.stack(out).cpu().numpy(), ref["logits"]) < 1e-4
        gout = torch.zeros(plan.NP, device="cuda")
        lg = [torch.zeros(xs[0].shape[0], plan.K, device="cuda") for _ in range(3)]
        loss, correct = step.step(xs, ys, grads_out=gout, logits_out=lg)
        assert relerr(torch.stack(lg).cpu().numpy(), ref["logits"]) < 1e-4, relerr(torch.stack(lg).cpu().numpy(), ref["logits"])
        assert relerr(loss.cpu().numpy(), ref["losses"]) < 1e-4
        G = step._gbuf[:3 * plan.P].view(3, plan.P).t().cpu().numpy()
        assert relerr(G, ref["G"]) < 1e-3, ("G", relerr(G, ref["G"]))
        got = grads_by_name(plan, gout)
        for k, v in ref.items():
            if k.startswith("grad:") and k[5:] in got:
                grp = next(p.group for p in plan.params if p.name == k[5:])
                if grp > 0:
                    assert relerr(got[k[5:]], v) < 1e-3, (st, k, relerr(got[k[5:]], v))
        sd = m.state_dict()
        for k, v in ref.items():
            if k.startswith("param:"):
                # parameters after the step: the update is lr * grad (1e-3 relative on grad => ~1e-6 of the parameter scale per step)
                close(sd[k[6:]].cpu().numpy(), v, 1e-4, f"step {st} param {k[6:]}")


@pytest.mark.parametrize("B,sync", [(64, True), (4097, True), (333, False)])
def test_bf16x3_path_vs_oracle_at_reference_and_large_batch(gk, B, sync):
    """B = 64 is the reference's default batch (weargait_train.py:661); 4097 = ragged last tile over many CTAs and all
    groups; async = three heads / three label vectors.  Gradients against the CPU oracle, 1e-3 norm-wise per tensor."""
    import gait_oracle as O
    torch.manual_seed(3)
    m = gk.WearGaitThreeModal(synchronized=sync).cuda()
    state = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    xs, y = O.synth_weargait_batch(B, seed=9)
    r = np.random.default_rng(1)
    ys = [y, y, y] if sync else [y, r.permutation(y), r.permutation(y)]
    counts = [[40, 90], [55, 60], [20, 30]]
    crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in counts]
    step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=gk.DTYPE_BF16X3, process_group=False)
    plan = m.set_window(64).plan()
    gout = torch.zeros(plan.NP, device="cuda")
    loss, correct = step.step([dev(x) for x in xs], [dev(v) for v in ys], grads_out=gout, update=False)
    p = O.canonical_params(state, sync)
    ex = O.weargait_train_step(p, {}, [torch.from_numpy(x) for x in xs], [torch.from_numpy(v) for v in ys],
                               synchronized=sync, wm="gcl", counts=counts, alpha=0.5)
    assert relerr(loss.cpu().numpy(), ex["losses"]) < 1e-4
    ref_correct = [int((l.argmax(1) == torch.from_numpy(v)).sum()) for l, v in zip(ex["logits"], ys)]
    assert np.abs(correct.cpu().numpy().round().astype(int) - np.array(ref_correct)).max() <= 1 + B // 2000
    G = step._gbuf[:3 * plan.P].view(3, plan.P).t().cpu().numpy()
    Gref = ex["G"].detach().numpy()
    assert relerr(G, Gref) < 1e-3, relerr(G, Gref)
    got = grads_by_name(plan, gout)
    for k, v in ex["grads"].items():
        if k in got and v is not None:
            grp = next(q.group for q in plan.params if q.name == k)
            if grp > 0:
                assert relerr(got[k], v.detach().numpy()) < 1e-3, (k, relerr(got[k], v.detach().numpy()))


def test_bf16x3_resident_gather_and_masks(gk):
    """win_start gather (aligned and unaligned windows), masked streams (zero input) and dropped tasks through the
    warp-specialised kernel equal the fp32 path on the same inputs."""
    import gait_oracle as O
    torch.manual_seed(4)
    m = gk.WearGaitThreeModal().cuda()
    B = 77
    xs, y = O.synth_weargait_batch(B + 3, seed=21)
    stores = [dev(np.ascontiguousarray(x.reshape(-1, x.shape[2]))) for x in xs]
    starts = dev((np.arange(B) * 64 + np.arange(B) % 3).astype(np.int64))       # odd starts -> unaligned windows
    yv = dev(y[:B])
    crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25, noise_mul=0.0) for _ in range(3)]
    plan = m.set_window(64).plan()
    res = {}
    for dt in (gk.DTYPE_F32, gk.DTYPE_BF16X3):
        st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=dt, process_group=False)
        gout = torch.zeros(plan.NP, device="cuda")
        loss, _ = st.step(stores, [yv] * 3, win_start=[starts] * 3, enabled=(True, False, True), tasks=(True, False, True),
                          grads_out=gout, update=False)
        res[dt] = (loss.cpu().numpy(), gout.cpu().numpy())
    assert relerr(res[gk.DTYPE_BF16X3][0], res[gk.DTYPE_F32][0]) < 1e-4
    assert relerr(res[gk.DTYPE_BF16X3][1], res[gk.DTYPE_F32][1]) < 1e-3


def test_fog_device_loaders_yield_reference_batches(gk):
    """A5 / f2: create_fusion_loaders on the device (ClipStore prepared by gaitk_fog_prepare_pose / _sensor, batches gathered by
    gaitk_window_gather) delivers, over two epochs of both loaders, exactly the samples the reference's DataLoaders delivered:
    labels and an order-sensitive fp64 checksum of every sample (data bit exact => checksums equal to rounding of the dot)."""
    import ref_harness as H
    DF = gk.dataloader_fbg_fog
    g = load_golden("fog_loaders")
    for cname, dataset, sync, modality, pad_skel, pad_sens in H.FOG_LOADER_CASES:
        reader, subs = H.synthetic_fog_reader(dataset, seed=5)
        tr_s, ev_s = H.fog_loader_split(subs, dataset)
        tr, ev = DF.create_fusion_loaders(dataset, reader, tr_s, ev_s, batch_size=7, synchronized=sync, seed=43, num_workers=0,
                                          pad_skel=pad_skel, pad_sens=pad_sens, modality=modality)
        got = H.loader_trace(tr, ev)
        for k, v in got.items():
            ref = g[f"{cname}/{k}"]
            assert v.shape == ref.shape, (cname, k)
            if k.endswith(("/cs", "/ct")):
                assert np.allclose(v, ref, rtol=0, atol=1e-9 * max(1.0, float(np.abs(ref).max()))), (cname, k, np.abs(v - ref).max())
            else:
                assert np.array_equal(v, ref), (cname, k)
        b = next(iter(ev))
        assert b["skeleton"].is_cuda and b["skeleton"].shape[1:] == (pad_skel, 7 if dataset == "fog" else 17, 3)
        assert b["sensor"].shape[1:] == (pad_sens, 6 if dataset == "fog" else 3) and b["label_sensor"].dtype == torch.int64


def test_fog_resident_clip_training_equals_dense_batch_training(gk):
    """index_batches(): the fused step reads the clips in place from the resident stores (win_start path) and lands on the
    same parameters as training on the dense batches the loader materialises."""
    import ref_harness as H
    DF = gk.dataloader_fbg_fog
    reader, subs = H.synthetic_fog_reader("fog", seed=5)
    tr_s, ev_s = H.fog_loader_split(subs, "fog")
    def run(resident):
        torch.manual_seed(3)
        m = gk.MultiModalMultiTaskModel(21, 6, 6, 6, 426, 16, 8, 128, 3, synchronized_loading=False).cuda()
        crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25.0, noise_mul=0.0) for c in ([30, 20, 10], [25, 22, 12])]
        step = gk.FusedTrainStep(m, crit, cagrad_c=0.1, private_mult=1.0)
        tr, _ = DF.create_fusion_loaders("fog", reader, tr_s, ev_s, batch_size=16, synchronized=False, seed=43, num_workers=0,
                                         pad_skel=101, pad_sens=426)
        if resident:
            m.set_window(101)                                      # pose length of the resident clips (no dense batch shows it)
            for ib in tr.index_batches():
                step.step(ib.frames, ib.ys, win_start=ib.win_start)
        else:
            for b in tr:
                step.step([b["skeleton"].flatten(2), b["sensor"]], [b["label_skeleton"], b["label_sensor"]])
        return m.flat_params().clone()
    a = run(False); b = run(True)
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------------------------
# 2-stream fusion baselines (feature_encoder.py:346-596, baselines/fusion_train.py) on the staged CUDA path
FOGBL_CASES = [f"fogbl_{k}_{s}" for k in ("early", "late", "share_latent", "cheap_xattn") for s in ("sync", "async")]


def fogbl_model(gk, g):
    meta = g["meta"]
    common = dict(skeleton_input_dim=21, skeleton_output_dim=6, sensor_in_channels=6, sensor_out_channels=6, sensor_length=426,
                  shared_out_channels=16, backbone_dim=8, num_classes=3, synchronized_loading=meta["synchronized"])
    cls = {"early": gk.EarlyFusionModel, "late": gk.LateFusionModel, "cheap_xattn": gk.CheapXAttnModel}.get(meta["kind"])
    m = gk.ShareLatentModel(**common, taskhead_input_dim=128) if cls is None else cls(**common)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "state0").items()}, strict=True)
    return m.cuda()


@pytest.mark.parametrize("name", FOGBL_CASES)
def test_two_stream_fusion_baselines_match_reference(gk, name):
    """EarlyFusionModel / LateFusionModel / ShareLatentModel / CheapXAttnModel: logits, loss and EVERY gradient of the staged CUDA
    path (encoder stages, cross attention, trunk stage, linear heads) against what the reference produced (fusion_train.py:234-242)."""
    g = load_golden(name); meta = g["meta"]
    m = fogbl_model(gk, g)
    out = m(dev(g["x_skel"]), dev(g["x_sens"]))
    ce = gk.CrossEntropyLoss()
    ys, yt = dev(g["ys"]), dev(g["yt"])
    if meta["synchronized"] and meta["kind"] != "share_latent":
        close(out.detach().cpu().numpy(), g["logits0"], 2e-5, "logits")
        loss = ce(out, ys)
    else:
        close(out[0].detach().cpu().numpy(), g["logits0"], 2e-5, "logits0"); close(out[1].detach().cpu().numpy(), g["logits1"], 2e-5, "logits1")
        loss = 0.5 * (ce(out[0], ys) + ce(out[1], yt))
    assert abs(float(loss.detach()) - float(g["loss"])) < 2e-5 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    n = 0
    for k, p in m.named_parameters():
        if f"grad:{k}" in g:
            close(p.grad.cpu().numpy(), g[f"grad:{k}"], 5e-5, f"grad {k}"); n += 1
    assert n == sum(1 for k in g if k.startswith("grad:")) and n >= 8


def test_fused_adam_matches_torch_adam(gk):
    """gaitk_adam (one launch over all parameter tensors) against torch.optim.Adam with the defaults of baselines/fusion_train.py:202,
    five steps on a 2-stream fusion baseline."""
    g = load_golden("fogbl_cheap_xattn_async")
    ma, mb = fogbl_model(gk, g), fogbl_model(gk, g)
    oa = gk.FusedAdam(ma.parameters(), lr=1e-3); ob = torch.optim.Adam(mb.parameters(), lr=1e-3)
    ce = gk.CrossEntropyLoss()
    xs, xt, ys, yt = dev(g["x_skel"]), dev(g["x_sens"]), dev(g["ys"]), dev(g["yt"])
    for _ in range(5):
        for m, o in ((ma, oa), (mb, ob)):
            o.zero_grad()
            a, b = m(xs, xt)
            (0.5 * (ce(a, ys) + ce(b, yt))).backward()
            o.step()
    for (k, p), q in zip(ma.named_parameters(), mb.parameters()):
        err = float((p - q).abs().max()); ref = float(q.abs().max())
        assert err <= 2e-6 * max(ref, 1.0), (k, err)


def test_xattn_and_linear_kernels_match_torch(gk):
    """The fusion ops alone against torch fp32 references of the same ops (forward and every gradient), at the shapes of both
    model families, ragged batch sizes included."""
    from importlib import import_module
    staged = gk.staged
    torch.manual_seed(0)
    for (n, T, d) in [(5, 64, 12), (3, 101, 6), (130, 64, 12)]:
        A = torch.randn(n, T, d, device="cuda", requires_grad=True); B = torch.randn(n, T, d, device="cuda", requires_grad=True)
        G = torch.randn(n, T, d, device="cuda")
        out = staged.cheap_xattn(A, B); out.backward(G)
        A2 = A.detach().clone().requires_grad_(); B2 = B.detach().clone().requires_grad_()
        ref = torch.softmax((A2 @ B2.transpose(1, 2)) * d ** -0.5, dim=-1) @ B2; ref.backward(G)
        close(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), 2e-5, "xattn out")
        close(A.grad.cpu().numpy(), A2.grad.cpu().numpy(), 5e-5, "xattn dA"); close(B.grad.cpu().numpy(), B2.grad.cpu().numpy(), 5e-5, "xattn dB")
    for (R, I, O, bias) in [(7, 128, 3, True), (600, 256, 2, True), (3 * 101, 6, 16, True), (1000, 128, 4, False)]:
        x = torch.randn(R, I, device="cuda", requires_grad=True); W = torch.randn(O, I, device="cuda", requires_grad=True)
        b = torch.randn(O, device="cuda", requires_grad=True) if bias else None
        G = torch.randn(R, O, device="cuda")
        y = staged.linear(x, W, b); y.backward(G)
        x2 = x.detach().clone().requires_grad_(); W2 = W.detach().clone().requires_grad_(); b2 = b.detach().clone().requires_grad_() if bias else None
        y2 = torch.nn.functional.linear(x2, W2, b2); y2.backward(G)
        close(y.detach().cpu().numpy(), y2.detach().cpu().numpy(), 2e-5, "linear y")
        close(x.grad.cpu().numpy(), x2.grad.cpu().numpy(), 2e-5, "linear dx"); close(W.grad.cpu().numpy(), W2.grad.cpu().numpy(), 5e-5, "linear dW")
        if bias:
            close(b.grad.cpu().numpy(), b2.grad.cpu().numpy(), 5e-5, "linear db")


# ---------------------------------------------------------------------------------------------------------------
# long windows: a window split by time over a thread-block cluster (distributed-shared-memory halos)
@pytest.mark.parametrize("T,kw,B", [(256, dict(enc_out_ch=24, shared_out_ch=32), 37), (256, {}, 21), (512, {}, 9), (128, dict(enc_out_ch=24, shared_out_ch=32), 19),
                                    (64, dict(enc_out_ch=24, shared_out_ch=32), 33)])
def test_long_windows_and_wide_models_match_oracle(gk, T, kw, B):
    """T = 256 / 512 run as clusters of 2 / 4 CTAs (conv halos through DSMEM, pooled features all-gathered, head on CTA 0);
    enc_out_ch / shared_out_ch other than the defaults use their own instantiations.  Two fused steps (async heads, GCL, CAGrad,
    SGD) against the CPU oracle: losses, accuracy counts, the per-task shared-gradient matrix G, parameters."""
    import gait_oracle as O
    torch.manual_seed(7)
    m = gk.WearGaitThreeModal(synchronized=False, **kw).cuda()
    state = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    xs, y = O.synth_weargait_batch(B, T, seed=31)
    r = np.random.default_rng(5); ys = [y, r.permutation(y), r.permutation(y)]
    counts = [[40, 60], [45, 55], [30, 70]]
    crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25.0, noise_mul=0.0) for c in counts]
    step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0)
    p = O.canonical_params(state, False); bufs = {}
    for it in range(2):
        loss, correct = step.step([dev(x) for x in xs], [dev(v) for v in ys])
        plan = m.plan()
        G = step._gbuf[:3 * plan.P].view(3, plan.P).cpu().numpy().copy()
        ex = O.weargait_train_step(p, bufs, [torch.from_numpy(x) for x in xs], [torch.from_numpy(v) for v in ys],
                                   synchronized=False, wm="gcl", counts=counts, alpha=0.5)
        close(loss.cpu().numpy(), ex["losses"], 5e-5, "loss")
        close(G, ex["G"].numpy().T, 2e-4, "G")
        ref_correct = [int((l.argmax(1) == torch.from_numpy(v)).sum()) for l, v in zip(ex["logits"], ys)]
        assert correct.cpu().numpy().round().astype(int).tolist() == ref_correct
    for k, v in m.state_dict().items():
        if k in p:
            close(v.cpu().numpy(), p[k].detach().numpy(), 3e-5, k)


def test_fused_step_with_gcl_noise_matches_the_criterion_path(gk):
    """GCL with noise_mul != 0 (classification_losses.py:99-105) inside the fused step: with the same seed the fused step
    draws the same clamped-normal noise as the criterion's own call, so one fused step equals one autograd-path step."""
    torch.manual_seed(0)
    B = 50
    xs = [torch.rand(B, 64, 2, device="cuda"), torch.randn(B, 64, 13, device="cuda"), torch.randn(B, 64, 24, device="cuda")]
    y = torch.randint(0, 2, (B,), device="cuda"); y[0], y[1] = 0, 1
    def build():
        torch.manual_seed(4)
        m = gk.WearGaitThreeModal(synchronized=True).cuda()
        crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.5) for _ in range(3)]
        return m, crit
    ma, ca = build()
    step = gk.FusedTrainStep(ma, ca, cagrad_c=0.5, private_mult=2.0)
    torch.manual_seed(123)
    loss_a, _ = step.step(xs, [y, y, y]); loss_a = loss_a.clone()
    mb, cb = build()
    opt = torch.optim.SGD(mb.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    cag = gk.CAGrad(n_tasks=3, device=torch.device("cuda"), c=0.5, max_norm=1.0)
    torch.manual_seed(123)
    lg = mb(*xs)
    L = [c(l, y) for c, l in zip(cb, lg)]
    close(loss_a.cpu().numpy(), torch.stack([l.detach() for l in L]).cpu().numpy(), 2e-5, "losses with noise")
    reference_style_step(mb, L, opt, cag, True)
    close(ma.flat_params().detach().cpu().numpy(), mb.flat_params().detach().cpu().numpy(), 2e-5, "parameters after one noisy step")


def test_graph_replay_with_fresh_tensors_every_step(gk):
    """The CUDA-graph path must not re-capture when the trainer hands over NEW tensors every step (the normal case): small
    inputs are staged into static buffers, so the graph cache holds one entry per shape and the result equals the eager path."""
    torch.manual_seed(2)
    def build(graph):
        torch.manual_seed(9)
        m = gk.WearGaitThreeModal(synchronized=True).cuda()
        crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
        return m, gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, use_graph=graph)
    ma, sa = build(True); mb, sb = build(False)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(12):
        B = 40 if it % 3 else 24                                      # two shapes
        xs = [torch.rand(B, 64, 2, device="cuda", generator=g), torch.randn(B, 64, 13, device="cuda", generator=g),
              torch.randn(B, 64, 24, device="cuda", generator=g)]     # fresh tensors, fresh addresses
        y = torch.randint(0, 2, (B,), device="cuda", generator=g); y[0], y[1] = 0, 1
        la, _ = sa.step([x.clone() for x in xs], [y.clone()] * 3)
        lb, _ = sb.step(xs, [y, y, y])
        assert torch.equal(la, lb), (it, la, lb)
    assert torch.equal(ma.flat_params(), mb.flat_params())
    # one staged graph per batch shape (+ at most a few address-keyed ones when the allocator hands the same blocks back), not one per step
    assert 2 <= len(sa._graphs) <= 6, len(sa._graphs)
