"""oracle/ref_harness.py -- drive the UNMODIFIED reference (staged by oracle/build_ref.py, or /root/reference when it is
mounted) for measurement and trainer-level parity.  Test / bench infrastructure: never imported by the product package.

What it gives:
  * `load_reference()`        sys.path set-up + the two documented compat shims (SURVEY.md 8(c) S1 pandas>=3 copy-on-write,
                              S2 torch.cuda.FloatTensor on a CPU run) and the reference modules as a namespace
  * `RefWearGaitStep`         the reference's own training step, verbatim call sequence of train/weargait_train.py:300-311:
                              forward_batch (:163-185) -> three criteria (make_criteria :111-130) -> step_cagrad_three
                              (:187-248, CAGrad.backward multitask_weighting.py:745-776 incl. SciPy SLSQP) -> SGD;
                              on CPU (reference arm) or on the B200 under stock torch-CUDA (north_star's ">= 20x" denominator)
  * `write_synthetic_weargait` synthetic WearGait PKLs + subject CSV names in the directory layout run_cv expects
                              (weargait_train.py:45-47,60-62; dataloader_weargait.py:388-418)
"""
from __future__ import annotations

import contextlib
import os
import sys
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
COUNTS = {"walkway": [400, 600], "insole": [400, 600], "imu": [400, 600]}      # unequal class counts: GCL needs them


def reference_root() -> Path | None:
    env = os.environ.get("GAIT_REFERENCE")
    for c in ([Path(env)] if env else []) + [HERE / "_ref", Path("/root/reference")]:
        if (c / "train" / "weargait_train.py").exists():
            return c
    return None


_REF = None


def load_reference():
    """-> namespace(root, WT, WE, FE, FT, DW, DF, GCLLoss, CAGrad) or None when no reference copy is available."""
    global _REF
    if _REF is not None:
        return _REF
    root = reference_root()
    if root is None:
        return None
    for p in (root, root / "data" / "WearGait", root / "train"):
        if str(p) not in sys.path:
            sys.path.insert(0, str(p))
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = torch.FloatTensor                      # shim S2
    import pandas as pd                                                 # noqa: F401
    import weargait_encoders as WE
    import feature_encoder as FE
    import weargait_train as WT
    import fbg_fog_train as FT
    from learning.optimizers.classification_losses import GCLLoss, LDAMLoss
    from learning.optimizers.multitask_weighting import CAGrad
    from data_processing import dataloader_weargait as DW
    from data_processing import dataloader_fbg_fog as DF

    # shim S1: apply_stats assigns into Series.to_numpy() (dataloader_weargait.py:217-220), read-only under pandas >= 3;
    # identical body plus .copy()
    def apply_stats(df, stats):
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
    if int(pd.__version__.split(".")[0]) >= 3:
        DW.apply_stats = apply_stats
    _REF = SimpleNamespace(root=root, WT=WT, WE=WE, FE=FE, FT=FT, DW=DW, DF=DF, GCLLoss=GCLLoss, LDAMLoss=LDAMLoss, CAGrad=CAGrad)
    return _REF


@contextlib.contextmanager
def _float_tensor_on(device: torch.device):
    """GCLLoss builds its margin list with torch.cuda.FloatTensor (classification_losses.py:83); for a CPU run of the
    reference on a box that has a GPU the constructor must yield CPU tensors (shim S2, scoped)."""
    if device.type == "cpu":
        old = torch.cuda.FloatTensor
        torch.cuda.FloatTensor = torch.FloatTensor
        try:
            yield
        finally:
            torch.cuda.FloatTensor = old
    else:
        yield


class RefWearGaitStep:
    """The reference's WearGait training step (sync loader semantics, GCL m=0.2 s=25 noise_mul=0, CAGrad c=0.5, SGD
    lr 1e-3 momentum 0.9 wd 1e-4 -- weargait_train.py:655-673 defaults + --wm gcl), model / criteria / CAGrad / optimizer
    all constructed by the reference's own code."""

    def __init__(self, device="cpu", *, seed: int = 0, threads: int | None = None, alpha: float = 0.5, wm: str = "gcl",
                 model_kw: dict | None = None):
        R = load_reference()
        if R is None:
            raise RuntimeError("no reference copy (oracle/_ref or /root/reference)")
        self.R = R
        self.device = torch.device(device)
        if threads:
            torch.set_num_threads(threads)
        R.WT.DEVICE = self.device                                        # module-level global read by forward_batch & co
        R.WT.set_seed(seed)
        self.model = R.WE.WearGaitThreeModal(synchronized=True, **(model_kw or {})).to(self.device)
        args = SimpleNamespace(wm=wm, gcl_m=0.2, gcl_s=25.0, noise_mul=0.0)
        with _float_tensor_on(self.device):
            self.crit = R.WT.make_criteria(args, COUNTS)
        self.cagrad = R.CAGrad(n_tasks=3, device=self.device, c=alpha)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)

    def step(self, xs, y):
        """xs: three (B,T,D) tensors and y (B,) as the sync collate yields them (host or device); returns the 3 losses."""
        WT = self.R.WT
        self.model.train()
        (lw, li, lm), (yw, yi, ym) = WT.forward_batch(self.model, {"xs": xs, "y": y}, False)
        Lw, Li, Lm = self.crit[0](lw, yw), self.crit[1](li, yi), self.crit[2](lm, ym)
        WT.step_cagrad_three(self.model, Lw, Li, Lm, self.opt, self.cagrad)
        return Lw, Li, Lm

    def time_steps(self, batches, steps: int, warmup: int):
        """seconds per step (list), host wall clock with a device synchronise on both sides when on CUDA; the loop also
        reads the three losses back (`.item()`), as train_one_epoch does every step (:317)."""
        out = []
        sync = (lambda: torch.cuda.synchronize(self.device)) if self.device.type == "cuda" else (lambda: None)
        for it in range(warmup + steps):
            xs, y = batches[it % len(batches)]
            sync(); t0 = time.perf_counter()
            L = self.step(xs, y)
            _ = [float(l.item()) for l in L]
            sync(); dt = time.perf_counter() - t0
            if it >= warmup:
                out.append(dt)
        return out


# ------------------------------------------------------------------------------------------------ synthetic datasets
def write_synthetic_weargait(root: Path, n_per_class: int = 6, seed: int = 0, frames=(700, 1500)):
    """SURVEY.md 8(d) Cfg 1 recipe: per class `n_per_class` subjects; per subject N ~ U[700,1500] frames @30 Hz; walkway
    2 ch ~U[0,1); insole 7 scalar ch ~N(0,1) + two 3-tuples, a little shorter (sync intersection); IMU 8 sites x (E,N,U)
    tuples ~N(0, sigma^2), sigma = 2 for PD and 1 for HC.  Column names come from the reference's own dataloader module;
    layout = what weargait_train.py:45-47,60-62 and dataloader_weargait.py:141-152 read under the CWD `root`.
    Returns (subject ids, labels)."""
    import pandas as pd
    DW = load_reference().DW
    rng = np.random.default_rng(seed)
    out = root / "data" / "WearGait" / "WearGait_preproc_SPmT_30Hz"
    out.mkdir(parents=True, exist_ok=True)
    sids, labels = [], []
    for cls, tag in enumerate(("hc", "pd")):
        d = root / "data" / "WearGait" / tag.upper()
        d.mkdir(parents=True, exist_ok=True)
        for i in range(n_per_class):
            sid = f"{tag}{i + 1:03d}"
            (d / f"{sid}_SelfPace_matTURN.csv").write_text("x\n")
            N = int(rng.integers(frames[0], frames[1]) * (1.4 if cls == 1 else 1.0))   # unequal class counts (GCL is NaN for equal ones)
            wk = pd.DataFrame({c: rng.random(N) for c in DW.WALKWAY_FIXED})
            Ni = N - int(rng.integers(0, 70))
            ins = pd.DataFrame({c: rng.standard_normal(Ni) for c in DW.INSOLE_NUMERIC[:7]})
            ins["Linsole_Acc"] = [tuple(v) for v in rng.standard_normal((Ni, 3))]
            ins["Rinsole_Acc"] = [tuple(v) for v in rng.standard_normal((Ni, 3))]
            sig = 2.0 if cls == 1 else 1.0
            imu = pd.DataFrame({f"{s_}_FreeAcc": [tuple(v) for v in rng.standard_normal((N, 3)) * sig] for s_ in DW.IMU_SITES})
            wk.to_pickle(out / f"{sid}_walkway.pkl"); ins.to_pickle(out / f"{sid}_insole.pkl"); imu.to_pickle(out / f"{sid}_imu.pkl")
            sids.append(sid); labels.append(cls)
    return sids, labels


# ---------------------------------------------------------------------------------------------- synthetic FoG / FBG readers
def synthetic_fog_reader(dataset: str = "fog", seed: int = 0, n_subjects: int = 9, scalar_labels: bool = True, informative: bool = False):
    """An object with the attributes `create_fusion_loaders` reads (dataloader_fbg_fog.py:291-327), in the key formats the
    reference's readers produce (preprocess_fog.py:107,146: `<SUB>_<video>_<segment>`; FBG: pose `<SUB>_<on|off>_walk_<i>`,
    GRF `<SUB>_<on|off>_walk` with a trial axis).  Clip lengths straddle the pad lengths; some sensor segments are missing
    and some have no pose partner, so the synchronised pairing is a proper intersection.  -> (reader, subjects)"""
    rng = np.random.default_rng(seed)
    pose, sens = {}, {}
    if dataset == "fog":
        subs = [f"SUB{i + 1:02d}" for i in range(n_subjects)]
        labels = {}
        for si, sub in enumerate(subs):
            # list and scalar forms (:314-318); the trainer's fold generator (utilities.py:97-101) needs the list form
            labels[sub] = [int(si % 3)] if (si % 2 == 0 or not scalar_labels) else int(si % 3)
            amp = 1.0 + (si % 3) if informative else 1.0                      # class-dependent scale: a learnable problem
            for v in range(2):
                for seg in range(1, 3 + int(rng.integers(0, 2))):
                    L = int(rng.integers(40, 150))
                    pose[f"{sub}_v{v}_{seg}"] = rng.random((L, 7, 3)) * 4.0 - 1.0
                    if informative:                                            # class-dependent jitter of the non-root joints (min-max keeps it)
                        pose[f"{sub}_v{v}_{seg}"][:, 1:, :] += rng.standard_normal((L, 6, 3)) * (si % 3)
                    if rng.random() < 0.85:
                        Ls = int(rng.integers(140, 520))
                        sens[f"{sub}_v{v}_{seg}"] = rng.standard_normal((Ls, 6)) * amp
                if rng.random() < 0.5:                                         # a sensor segment without a pose partner
                    sens[f"{sub}_v{v}_9"] = rng.standard_normal((int(rng.integers(140, 520)), 6))
        reader = SimpleNamespace(pose_dict=pose, sensor_dict=sens, labels_dict=labels)
        return reader, subs
    subs = []
    pl, sl = {}, {}
    for si in range(n_subjects):
        for state in ("on", "off"):
            if state == "off" and si % 3 == 2:
                continue
            sub = f"SUB{si + 1:02d}_{state}"; subs.append(sub)
            pl[sub] = int((si + (state == "off")) % 3)
            n_tr = 2 + int(rng.integers(0, 3))
            for i in range(n_tr):
                pose[f"{sub}_walk_{i}"] = rng.random((int(rng.integers(60, 130)), 17, 3)) * 2.0
            grf = rng.standard_normal((int(rng.integers(50, 90)), n_tr + int(rng.integers(0, 2)), 3))
            sens[f"{sub}_walk"] = grf; sl[f"{sub}_walk"] = pl[sub]
    reader = SimpleNamespace(pose_dict=pose, sensor_dict=sens, pose_label_dict=pl, sensor_label_dict=sl)
    return reader, subs


FOG_LOADER_CASES = [   # (name, dataset, synchronized, modality, pad_skel, pad_sens)
    ("fog_sync", "fog", True, "multimodal", 101, 426),
    ("fog_async", "fog", False, "multimodal", 101, 426),
    ("fog_skeleton", "fog", False, "skeleton", 101, 426),
    ("fog_sensor", "fog", False, "sensor", 101, 426),
    ("fbg_sync", "fbg", True, "multimodal", 101, 65),
    ("fbg_async", "fbg", False, "multimodal", 101, 65),
]


def fog_loader_split(subs, dataset):
    """deterministic train / eval split that keeps every class on both sides"""
    ev = subs[::3]
    return [s for s in subs if s not in ev], ev


def loader_trace(train_loader, eval_loader, epochs: int = 2):
    """what a trainer sees: per pass and batch, the labels and an order-sensitive fp64 checksum of every sample"""
    out = {}
    for ep in range(epochs):
        for nm, ld in (("train", train_loader), ("eval", eval_loader)):
            ys, yt, cs, ct, bs = [], [], [], [], []
            for b in ld:
                sk = b["skeleton"].detach().cpu().numpy().astype(np.float64); se = b["sensor"].detach().cpu().numpy().astype(np.float64)
                n = sk.shape[0]; bs.append(n)
                w1 = np.cos(np.arange(sk[0].size, dtype=np.float64)); w2 = np.cos(np.arange(se[0].size, dtype=np.float64))
                cs += [float(np.dot(sk[i].ravel(), w1)) for i in range(n)]; ct += [float(np.dot(se[i].ravel(), w2)) for i in range(n)]
                ys += b["label_skeleton"].tolist(); yt += b["label_sensor"].tolist()
            out[f"{nm}_ep{ep}/ys"] = np.array(ys); out[f"{nm}_ep{ep}/yt"] = np.array(yt)
            out[f"{nm}_ep{ep}/cs"] = np.array(cs); out[f"{nm}_ep{ep}/ct"] = np.array(ct); out[f"{nm}_ep{ep}/bs"] = np.array(bs)
    return out


# ---------------------------------------------------------------------------------------------- synthetic raw WearGait CSVs (ETL)
def write_synthetic_weargait_csvs(root: Path, n_per_class: int = 2, seed: int = 0, rows=(400, 700)):
    """Raw recordings in the layout preprocess_weargait.run_end_to_end expects (preprocess_weargait.py:14-19, 22-52): per-subject
    `<SID>_SelfPace_matTURN.csv` at ~100 Hz with a `Time` column written as "12,345 sec", a `GeneralEvent` column with some
    "Standing" rows, foot pressures, insole forces / CoP / accelerometers, eight IMU sites, NaN holes and one subject without an
    IMU site; demographic sheets whose real header is the second row.  -> (hc_root, pd_root, hc_demo, pd_demo, subject ids)"""
    import pandas as pd
    rng = np.random.default_rng(seed)
    sites = ["L_Ankle", "R_Ankle", "L_DorsalFoot", "R_DorsalFoot", "L_MidLatThigh", "R_MidLatThigh", "L_LatShank", "R_LatShank"]
    sids = []
    out = {}
    for cls, tag in ((0, "HC"), (1, "PD")):
        d = root / "data" / "WearGait" / tag
        d.mkdir(parents=True, exist_ok=True)
        demo = [["", "", ""], ["Subject ID", "Age", "Weight (kg)"]]
        for i in range(n_per_class):
            sid = f"{tag}{i + 1:03d}"; sids.append(sid)
            n = int(rng.integers(rows[0], rows[1]))
            t = np.cumsum(rng.uniform(0.006, 0.014, n)) - 0.2 * (i == 1)          # irregular ~100 Hz clock; one subject starts at t < 0
            cols = {"Time": [f"{v:.4f}".replace(".", ",") + " sec" for v in t],
                    "GeneralEvent": np.where(rng.random(n) < 0.1, "Standing", "Walking"),
                    "L Foot Pressure": rng.random(n) * 800, "R Foot Pressure": rng.random(n) * 800,
                    "LTotalForce": rng.random(n) * 700, "RTotalForce": rng.random(n) * 700,
                    "LCoP_X": rng.standard_normal(n), "LCoP_Y": rng.standard_normal(n), "RCoP_X": rng.standard_normal(n), "RCoP_Y": rng.standard_normal(n)}
            for side in ("Linsole", "Rinsole"):
                for ax in "XYZ":
                    cols[f"{side}:Acc_{ax}"] = rng.standard_normal(n) * 3 + 1
            for s_ in sites:
                if cls == 1 and i == 0 and s_ == "R_LatShank":
                    continue
                for ax in "ENU":
                    cols[f"{s_}_FreeAcc_{ax}"] = rng.standard_normal(n) * (2.0 if cls else 1.0)
            df = pd.DataFrame(cols)
            for c in ("L Foot Pressure", "LCoP_X", "Linsole:Acc_Y", "L_Ankle_FreeAcc_E"):
                df.loc[rng.random(n) < 0.05, c] = np.nan
            df.loc[5:9, "Time"] = "n/a"
            df.to_csv(d / f"{sid}_SelfPace_matTURN.csv", index=False)
            demo.append([sid, str(50 + i), f"{60 + 7.5 * i + 5 * cls:.1f} kg"])
        pd.DataFrame(demo).to_csv(d / f"{tag.lower()}_demographic.csv", header=False, index=False)
        out[tag] = (str(d), str(d / f"{tag.lower()}_demographic.csv"))
    return out["HC"][0], out["PD"][0], out["HC"][1], out["PD"][1], sids
