"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/gaitk.h
declares, host-only entries (window_indices) match the reference goldens, the drop-in modules keep the
reference's state_dict contract, and compute entries refuse to run without a GPU (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import load_golden, ROOT


def test_library_exports_every_declared_symbol():
    import gaitk
    hdr = (ROOT / "include" / "gaitk.h").read_text()
    declared = sorted(set(re.findall(r"\b(gaitk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = gaitk.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in gaitk.h but not exported"
    assert sorted(gaitk._lib.EXPORTS) == declared
    assert L.gaitk_version() == 100


def test_window_indices_c_abi_bit_exact():
    import gaitk
    L = gaitk.lib()
    g = load_golden("data_path")
    for i, (n, w, h) in enumerate(g["win_cases"]):
        cap = 4096
        buf = (C.c_int64 * (3 * cap))()
        cnt = L.gaitk_window_indices(int(n), int(w), int(h), buf, cap)
        got = np.frombuffer(buf, dtype=np.int64)[:3 * cnt].reshape(-1, 3)
        assert (got == g[f"win_{i}"]).all()


def test_dropin_state_dict_contract():
    import gaitk
    for sync in (True, False):
        m = gaitk.WearGaitThreeModal(synchronized=sync, use_norm=True)
        keys = list(m.state_dict().keys())
        assert "enc_i.ln1.weight" in keys and "enc_i.skip.weight" in keys and "backbone.conv.weight" in keys
        assert ("_shared_head.fc.weight" in keys) == sync
        assert m.head_w is m.head_i if sync else m.head_w is not m.head_i
        assert len(m.get_shared_parameters()) == (6 if sync else 2)
    g = load_golden("wg_sync_gcl")
    m = gaitk.WearGaitThreeModal()
    ref_keys = [k[len("state0/"):] for k in g if k.startswith("state0/")]
    assert sorted(m.state_dict().keys()) == sorted(ref_keys)
    f = gaitk.MultiModalMultiTaskModel(21, 6, 6, 6, 426, 16, 8, 128, 3, synchronized_loading=True)
    g2 = load_golden("fog_sync_gcl")
    assert sorted(f.state_dict().keys()) == sorted(k[len("state0/"):] for k in g2 if k.startswith("state0/"))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import gaitk
    m = gaitk.WearGaitThreeModal()
    with pytest.raises(gaitk.GaitkError):
        m(torch.zeros(2, 64, 2), torch.zeros(2, 64, 13), torch.zeros(2, 64, 24))
    with pytest.raises(gaitk.GaitkError):
        gaitk.CrossEntropyLoss()(torch.zeros(2, 2), torch.zeros(2, dtype=torch.long))


def _solve_host(A32, n, alpha, solver=0):
    import gaitk
    L = gaitk.lib()
    a = np.zeros((3, 3), np.float32); a[:n, :n] = A32
    w = (C.c_double * 3)(); it = C.c_int()
    mode = L.gaitk_cagrad_solve_host(a.ctypes.data_as(C.POINTER(C.c_float)), n, float(alpha), solver, w, C.byref(it))
    return np.array(w[:n]), mode, it.value


def test_slsqp_restatement_matches_scipy_golden_corpus():
    """The C++ SLSQP restatement (identical code runs on the device) against the weights the reference's
    SciPy call produced (golden).  Parity range: Gram entries <= 1e6 (see cagrad_solver.cuh)."""
    g = load_golden("cagrad_corpus")
    worst = 0.0; checked = 0
    for G, wref, (n, alpha) in zip(g["G"], g["w"], g["n_alpha"]):
        n = int(n); Gn = G[:, :n].astype(np.float32)
        A = (torch.from_numpy(Gn).t().mm(torch.from_numpy(Gn))).numpy()
        if A.max() > 1e6:
            continue
        w, mode, it = _solve_host(A, n, alpha)
        assert mode == 0
        worst = max(worst, np.abs(w - wref[:n]).max()); checked += 1
    assert checked >= 60 and worst < 2e-4, (checked, worst)


def test_slsqp_restatement_matches_scipy_random():
    import gait_oracle as O
    rng = np.random.default_rng(7)
    errs = []
    for t in range(300):
        n = int(rng.choice([2, 3])); P = int(rng.choice([8, 24, 300]))
        G = rng.standard_normal((P, n)).astype(np.float32)
        kind = t % 8
        if kind == 1: G[:, 1] = G[:, 0] * 0.7
        elif kind == 2: G[:, -1] = -G[:, 0] + 0.05 * G[:, -1]
        elif kind == 3: G *= 1e-3
        elif kind == 4: G[:, 0] *= 6
        elif kind == 5: G[:, -1] = 0
        elif kind == 6: G = np.abs(G)
        elif kind == 7: G *= 10 ** rng.uniform(-3, 1.2)
        alpha = float(rng.choice([0.5, 0.1, 0.4, 1.0]))
        _, A, wref = O.cagrad_combine(torch.from_numpy(G), alpha)
        if A.max() > 1e6:
            continue
        w, mode, it = _solve_host(A, n, alpha)
        errs.append(np.abs(w - wref).max())
    errs = np.array(errs)
    assert errs.max() < 5e-4 and np.quantile(errs, 0.95) < 1e-5, (errs.max(), np.quantile(errs, 0.95))


def test_exact_solver_never_worse_than_slsqp():
    import gait_oracle as O
    g = load_golden("cagrad_corpus")
    for G, wref, (n, alpha) in zip(g["G"], g["w"], g["n_alpha"]):
        n = int(n); Gn = G[:, :n].astype(np.float32)
        A = (torch.from_numpy(Gn).t().mm(torch.from_numpy(Gn))).numpy()
        w, _, _ = _solve_host(A, n, alpha, solver=1)
        assert abs(w.sum() - 1) < 1e-12 and (w >= 0).all()
        f_ours = O.cagrad_objective(A, w, float(alpha)); f_ref = O.cagrad_objective(A, wref[:n], float(alpha))
        assert f_ours <= f_ref + 1e-9 * max(1.0, abs(f_ref))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a tiny sample: one JSON line
    with the contract's keys; under a multi-rank launch only rank 0 prints."""
    import json, os, subprocess, sys
    env = dict(os.environ); env.pop("RANK", None)
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-batch", "128"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "windows/s"
    # "reference" when the unmodified reference is staged under oracle/_ref (python oracle/build_ref.py), else the oracle port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    env["RANK"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("t_in,t_out,cin,cout", [(426, 101, 6, 6), (426, 101, 3, 3), (300, 101, 3, 3), (101, 101, 6, 6), (7, 5, 2, 4)])
def test_pooled_taps_identity_of_the_sensor_encoder(t_in, t_out, cin, cout):
    """The pooled-taps kernel (ENC_POOL_LINEAR, csrc/stream_common.cuh) rests on: AdaptiveAvgPool1d(conv1d_k3(x)) ==
    Linear over the three tap-shifted bin means of x with conv1d.weight (C, Cin, 3) read as (C, 3 Cin), because
    SensorEncoder (feature_encoder.py:27-58) has nothing non-linear between the two.  Checked here against torch
    (fp32 CPU) with the bins the kernel uses: [floor(i T_in / T), ceil((i + 1) T_in / T)), zero padding outside the clip."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, t_in, cin, generator=g)
    w = torch.randn(cout, cin, 3, generator=g) * 0.3
    b = torch.randn(cout, generator=g)
    ref = torch.nn.functional.adaptive_avg_pool1d(torch.nn.functional.conv1d(x.transpose(1, 2), w, b, padding=1), t_out).transpose(1, 2)
    xp = torch.nn.functional.pad(x, (0, 0, 1, 1))                       # frame t of x is row t + 1
    P = torch.zeros(3, t_out, cin * 3)
    for i in range(t_out):
        s0, s1 = (i * t_in) // t_out, -((-(i + 1) * t_in) // t_out)
        for tap in range(3):
            P[:, i, tap::3] = xp[:, s0 + tap:s1 + tap].mean(1)          # channel ci * 3 + tap
    got = P @ w.reshape(cout, cin * 3).t() + b
    assert torch.allclose(got, ref, rtol=1e-5, atol=2e-6), float((got - ref).abs().max())
