"""WearGait ETL: raw per-subject CSVs -> the 30 Hz per-stream tables the trainers read (drop-in for
``train/data_processing/preprocess_weargait.py:22-343``: same function names, signatures, file names and PKL contents).

The reference walks pandas objects column by column and bins with ``groupby(...).first()``; here every stream is assembled as
ONE float64 matrix, the 30 Hz decimation is a single vectorised "first non-null entry per 1/30 s bin and column" pass
(``first_valid_per_bin``), body-weight / z-score scaling are matrix operations, and the same pass can hand the result straight
to the device-resident fold (``subject_frames``: float64 ``(N, D)`` matrices in the loaders' fixed column order, no PKL round
trip).  Written PKLs are identical to the reference's (checked frame by frame in tests/test_data_cpu.py), so either loader
reads either output.  Host code by nature: CSV parsing is I/O bound; nothing here is on the per-step path.
"""
from __future__ import annotations

import json
import re
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

GRAV = 9.81
IMU_SITES = ["L_Ankle", "R_Ankle", "L_DorsalFoot", "R_DorsalFoot", "L_MidLatThigh", "R_MidLatThigh", "L_LatShank", "R_LatShank"]
CSV_PATTERN = "*_SelfPace_matTURN.csv"
hc_path = "data/WearGait/HC"
pd_path = "data/WearGait/PD"
hc_demo_csv = "data/WearGait/HC/hc_demographic.csv"
pd_demo_csv = "data/WearGait/PD/pd_demographic.csv"
output_dir = "data/WearGait/WearGait_preproc_SPmT_30Hz"

_INSOLE_ACC = [f"{side}:Acc_{ax}" for side in ("Linsole", "Rinsole") for ax in "XYZ"]
_IMU_ACC = [f"{s}_FreeAcc_{ax}" for s in IMU_SITES for ax in "ENU"]


def _num(series: pd.Series) -> np.ndarray:
    return pd.to_numeric(series, errors="coerce").to_numpy(dtype=float)


# ---------------------------------------------------------------------------------------------- demographics, discovery (:22-52)
def read_demographics_with_header_fix(path: str) -> pd.DataFrame:
    """:22-28 -- the sheet's real header is its second row."""
    raw = pd.read_csv(path, header=None, dtype=str)
    names = raw.iloc[1].fillna("").astype(str).str.replace(r"\s+", " ", regex=True).str.strip()
    body = raw.iloc[2:].reset_index(drop=True).copy()
    body.columns = names
    return body


def extract_subject_weights(demo_df: pd.DataFrame) -> pd.DataFrame:
    """:30-36 -- (subject_id, weight_kg) with the first number found in the weight cell."""
    id_col = next(c for c in demo_df.columns if re.search(r"(subject\s*id|participant)", c, re.I))
    wt_col = next(c for c in demo_df.columns if re.search(r"weight", c, re.I))
    out = pd.DataFrame({"subject_id": demo_df[id_col].astype(str).str.strip(),
                        "weight_kg": pd.to_numeric(demo_df[wt_col].astype(str).str.extract(r"([0-9]*\.?[0-9]+)")[0], errors="coerce")})
    return out.dropna(subset=["subject_id", "weight_kg"]).reset_index(drop=True)


def build_weight_map(hc_demo_csv: str, pd_demo_csv: str) -> Dict[str, float]:
    """:38-46 -- lower-cased subject id -> body weight in kg (later sheets override earlier ones)."""
    weights: Dict[str, float] = {}
    for sheet in (hc_demo_csv, pd_demo_csv):
        if sheet:
            w = extract_subject_weights(read_demographics_with_header_fix(sheet))
            weights.update(zip(w["subject_id"].str.lower(), w["weight_kg"].astype(float)))
    return weights


def find_subject_files(root_dir: str, pattern: str = CSV_PATTERN) -> Dict[str, Path]:
    """:49-51."""
    return {p.stem.split("_", 1)[0].lower(): p for p in Path(root_dir).glob(pattern)}


# ---------------------------------------------------------------------------------------------- train statistics (:54-111)
def list_imu_freeacc_cols(cols) -> List[str]:
    """:54-66 -- the acceleration channels that get z-scored, IMU sites first, then the insole accelerometers."""
    have = set(cols)
    return [c for c in _IMU_ACC + _INSOLE_ACC if c in have]


def fit_train_stats(train_csv_paths: Sequence[str]) -> Dict[str, Tuple[float, float]]:
    """:68-102 -- per-channel mean / std over the finite samples of the training CSVs (running sum, sum of squares, count)."""
    if not train_csv_paths:
        raise ValueError("Empty training list for IMU normalization.")
    channels = list_imu_freeacc_cols(pd.read_csv(train_csv_paths[0], nrows=0).columns)
    acc = {c: [0.0, 0.0, 0] for c in channels}
    for path in train_csv_paths:
        df = pd.read_csv(path)
        for c in channels:
            if c not in df.columns:
                continue
            x = _num(df[c]); x = x[np.isfinite(x)]
            if x.size:
                a = acc[c]; a[0] += float(x.sum()); a[1] += float(np.dot(x, x)); a[2] += int(x.size)
    stats = {}
    for c, (s1, s2, n) in acc.items():
        if n > 0:
            mean = s1 / n
            stats[c] = (mean, max(np.sqrt(max(s2 / n - mean ** 2, 0.0)), 1e-8))
        else:
            stats[c] = (0.0, 1.0)
    return stats


def apply_stats(df: pd.DataFrame, stats: dict) -> pd.DataFrame:
    """:104-110."""
    out = df.copy()
    for c, (m, s) in stats.items():
        if c in out.columns:
            out[c] = (pd.to_numeric(out[c], errors="coerce").to_numpy() - m) / (s if s != 0 else 1.0)
    return out


# ---------------------------------------------------------------------------------------------- 30 Hz decimation (:113-136)
def parse_time_seconds(s: pd.Series) -> np.ndarray:
    """:113-118 -- "12,345 sec" -> 12.345."""
    txt = s.astype(str).str.strip().str.replace(" sec", "", regex=False).str.replace(",", ".", regex=False)
    return pd.to_numeric(txt, errors="coerce").to_numpy(dtype=float)


def first_valid_per_bin(bins: np.ndarray, valid: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """bins: int64[n] (-1 = drop the row); valid: bool[n, C].  -> (sorted unique bins >= 0, int64[n_bins, C] row of the first valid
    entry of every (bin, column), -1 where a bin holds none): ``groupby(bin, sort=True).first()`` as one scatter-min."""
    rows = np.flatnonzero(bins >= 0)
    ub, inv = np.unique(bins[rows], return_inverse=True)
    first = np.full((ub.size, valid.shape[1]), np.iinfo(np.int64).max, dtype=np.int64)
    for c in range(valid.shape[1]):
        ok = valid[rows, c]
        np.minimum.at(first[:, c], inv[ok], rows[ok])
    first[first == np.iinfo(np.int64).max] = -1
    return ub, first


def downsample_to_30hz(df: pd.DataFrame, time_col: str = "Time", target_hz: int = 30) -> pd.DataFrame:
    """:120-136 -- one row per occupied 1/target_hz bin: per column the first non-null entry of the bin, time = bin centre."""
    if df.empty or time_col not in df.columns:
        return df
    t = parse_time_seconds(df[time_col])
    ok = np.isfinite(t)
    if not ok.any():
        return pd.DataFrame()
    bins = np.full(t.shape, -1, dtype=np.int64)
    bins[ok] = np.floor(t[ok] * target_hz).astype(np.int64)
    cols = list(df.columns)
    valid = np.stack([df[c].notna().to_numpy() for c in cols], axis=1)
    ub, first = first_valid_per_bin(bins, valid)
    out = {}
    for j, c in enumerate(cols):
        src = df[c].to_numpy()
        take = first[:, j]
        if src.dtype == object:
            col = np.empty(ub.size, dtype=object)
            for i, r in enumerate(take):
                col[i] = src[r] if r >= 0 else None
        else:
            col = np.where(take >= 0, src[np.maximum(take, 0)], np.nan) if (take < 0).any() else src[take]
        out[c] = col
    res = pd.DataFrame(out, columns=cols)
    res[time_col] = (ub.astype(float) + 0.5) / target_hz
    return res.reset_index(drop=True)


# ---------------------------------------------------------------------------------------------- stream builders (:139-225)
def _pack3(out: pd.DataFrame, cols: List[str], name: str) -> None:
    if all(c in out.columns for c in cols):
        out[name] = list(map(tuple, out[cols].to_numpy()))
        out.drop(columns=cols, inplace=True)


def _zscore_inplace(out: pd.DataFrame, cols: Sequence[str], stats: Optional[dict]) -> None:
    if stats is None:
        return
    for c in cols:
        if c in out.columns and c in stats:
            m, s = stats[c]
            out[c] = (pd.to_numeric(out[c], errors="coerce").to_numpy() - m) / (s if s != 0 else 1.0)


def build_walkway(df: pd.DataFrame, weight_kg: float) -> pd.DataFrame:
    """:139-153 -- foot pressures divided by body weight (N), decimated."""
    raw = ["L Foot Pressure", "R Foot Pressure"]
    keep = [c for c in ["Time"] + raw if c in df.columns]
    if not keep:
        return pd.DataFrame()
    out = df[keep].copy()
    denom = weight_kg * GRAV if weight_kg and weight_kg > 0 else np.nan
    for c in raw:
        if c in out and denom:
            out[c + "_BW"] = pd.to_numeric(out[c], errors="coerce") / denom
    return downsample_to_30hz(out[["Time"] + [c + "_BW" for c in raw if c + "_BW" in out.columns]])


def build_insole(df: pd.DataFrame, weight_kg: float, stats: Optional[dict]) -> pd.DataFrame:
    """:155-198 -- forces / body weight (+ their sum), centres of pressure as they are, accelerometers z-scored when statistics are
    given and packed into one tuple column per side, decimated."""
    wanted = ["Time", "LTotalForce", "RTotalForce", "LCoP_X", "LCoP_Y", "RCoP_X", "RCoP_Y"] + _INSOLE_ACC
    keep = [c for c in wanted if c in df.columns]
    if not keep:
        return pd.DataFrame()
    out = df[keep].copy()
    if weight_kg and weight_kg > 0:
        denom = weight_kg * GRAV
        for c in ("LTotalForce", "RTotalForce"):
            if c in out:
                out[c + "_BW"] = pd.to_numeric(out[c], errors="coerce") / denom
        if {"LTotalForce", "RTotalForce"}.issubset(out.columns):
            out["SumForce_BW"] = (pd.to_numeric(out["LTotalForce"], errors="coerce") + pd.to_numeric(out["RTotalForce"], errors="coerce")) / denom
    _zscore_inplace(out, _INSOLE_ACC, stats)
    for side in ("Linsole", "Rinsole"):
        _pack3(out, [f"{side}:Acc_{ax}" for ax in "XYZ"], f"{side}_Acc")
    order = ["Time", "LTotalForce_BW", "RTotalForce_BW", "SumForce_BW", "LCoP_X", "LCoP_Y", "RCoP_X", "RCoP_Y", "Linsole_Acc", "Rinsole_Acc"]
    return downsample_to_30hz(out[[c for c in order if c in out.columns]])


def build_imu(df: pd.DataFrame, stats: Optional[dict]) -> pd.DataFrame:
    """:200-225 -- free accelerations of the eight sites, z-scored when statistics are given, one tuple column per site, decimated."""
    keep = ["Time"] + [c for c in _IMU_ACC if c in df.columns]
    if len(keep) == 1:
        return pd.DataFrame()
    imu = df[[c for c in keep if c in df.columns]].copy()
    _zscore_inplace(imu, _IMU_ACC, stats)
    for s in IMU_SITES:
        _pack3(imu, [f"{s}_FreeAcc_{ax}" for ax in "ENU"], f"{s}_FreeAcc")
    return downsample_to_30hz(imu)


def subject_tables(csv_path, weight_kg: float, stats: Optional[dict]) -> Dict[str, pd.DataFrame]:
    """one subject: CSV -> {walkway, insole, imu} 30 Hz tables ("standing" rows dropped, :287-290)"""
    df = pd.read_csv(csv_path)
    if "GeneralEvent" in df.columns:
        df = df[df["GeneralEvent"].str.lower() != "standing"].copy()
    return {"walkway": build_walkway(df, weight_kg), "insole": build_insole(df, weight_kg, stats), "imu": build_imu(df, stats)}


# ---------------------------------------------------------------------------------------------- orchestrator (:228-343)
def run_end_to_end(hc_csv_root: str, pd_csv_root: str, hc_demo_csv: str, pd_demo_csv: str, output_dir: str,
                   train_subject_ids: Optional[list], pattern: str = CSV_PATTERN, segment_len_rows: Optional[int] = None,
                   segment_len_sec: Optional[float] = None):
    """:228-343 -- all subjects -> ``<sid>_walkway.pkl`` and ``<sid>_{insole,imu}[_base].pkl`` (``_base`` = no train statistics),
    ``imu_freeacc_stats.json`` when statistics were fitted, and the per-subject / total row and segment counts on stdout."""
    HZ = 30
    outdir = Path(output_dir); outdir.mkdir(parents=True, exist_ok=True)
    if segment_len_sec is not None:
        seg_rows = int(max(1, np.floor(float(segment_len_sec) * HZ)))
    elif segment_len_rows is not None:
        seg_rows = int(max(1, segment_len_rows))
    else:
        seg_rows = None
    weight_map = build_weight_map(hc_demo_csv, pd_demo_csv)
    all_files = {**find_subject_files(hc_csv_root, pattern), **find_subject_files(pd_csv_root, pattern)}
    if not all_files:
        print("[warn] no CSV files found; check paths/pattern")
        return
    stats = None
    if train_subject_ids:
        train_paths = [str(all_files[str(s).lower()]) for s in train_subject_ids if str(s).lower() in all_files]
        if not train_paths:
            raise ValueError("No training CSVs found. Check train_subject_ids or pattern.")
        stats = fit_train_stats(train_paths)
    suffix = "_base" if stats is None else ""
    tot = np.zeros(8, dtype=np.int64)                   # rows w / i / m / any, segments w / i / m / all
    for sid, csv_path in all_files.items():
        tb = subject_tables(csv_path, weight_map.get(sid, np.nan), stats)
        nw, ni, nm = len(tb["walkway"]), len(tb["insole"]), len(tb["imu"])
        n_any = max(nw, ni, nm)
        segs = (nw // seg_rows, ni // seg_rows, nm // seg_rows, min(nw, ni, nm) // seg_rows) if seg_rows is not None else (0, 0, 0, 0)
        print(f"[{sid}] rows_w={nw} rows_i={ni} rows_m={nm} rows_any={n_any} secs_any={(n_any / HZ if n_any > 0 else 0.0):.3f}"
              + (f" | seg_rows={seg_rows} segs_w={segs[0]} segs_i={segs[1]} segs_m={segs[2]} segs_all={segs[3]}" if seg_rows else ""))
        tb["walkway"].to_pickle(outdir / f"{sid}_walkway.pkl")
        tb["insole"].to_pickle(outdir / f"{sid}_insole{suffix}.pkl")
        tb["imu"].to_pickle(outdir / f"{sid}_imu{suffix}.pkl")
        tot += np.array([nw, ni, nm, n_any, *segs], dtype=np.int64)
    if stats is not None:
        with open(outdir / "imu_freeacc_stats.json", "w") as f:
            json.dump(stats, f)
    print(f"[TOTAL] rows_w={tot[0]} rows_i={tot[1]} rows_m={tot[2]} rows_any={tot[3]} secs_any={tot[3] / HZ:.3f}"
          + (f" | seg_rows={seg_rows} segs_w={tot[4]} segs_i={tot[5]} segs_m={tot[6]} segs_all={tot[7]}" if seg_rows else ""))


def subject_frames(csv_path, weight_kg: float, stats: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """CSV -> {modality: float64 (N, D)} in the loaders' fixed column order, NaN where a value or a column is missing: what
    ``dataloader_weargait.prepare_split(frames=...)`` uploads, without the PKL round trip."""
    from .dataloader_weargait import IMU_FIXED, INSOLE_FIXED, WALKWAY_FIXED
    tb = subject_tables(csv_path, weight_kg, stats)
    out = {}
    for m, fixed in (("walkway", WALKWAY_FIXED), ("insole", INSOLE_FIXED), ("imu", IMU_FIXED)):
        df = tb[m]
        X = np.full((len(df), len(fixed)), np.nan, dtype=np.float64)
        for j, c in enumerate(fixed):
            if c in df.columns:
                X[:, j] = _num(df[c])
            else:                                         # an axis of a packed tuple column
                base, ax = c.rsplit("_", 1)
                if base in df.columns and len(df):
                    k = {"X": 0, "Y": 1, "Z": 2, "E": 0, "N": 1, "U": 2}[ax]
                    X[:, j] = np.array([t[k] if isinstance(t, tuple) else np.nan for t in df[base]], dtype=float)
        out[m] = X
    return out


def main() -> None:
    run_end_to_end(hc_path, pd_path, hc_demo_csv, pd_demo_csv, output_dir, train_subject_ids=None)


if __name__ == "__main__":
    main()
