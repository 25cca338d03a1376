"""Device-resident WearGait data path: the B200 counterpart of ``train/data_processing/dataloader_weargait.py``.

The reference prepares a fold on the host (pandas / numpy, :181-299, :388-418), keeps every window as its own
float64 ndarray in a dict, and lets DataLoader workers stack and pin one batch at a time (:305-384, :420-455).
Here a fold is prepared ONCE on the device:

  * per-channel statistics over the finite train frames      -> ``gaitk_stats_accumulate`` / ``_finalize``  (:181-210)
  * NaN -> mean, z-score, nan_to_num, cast to fp32            -> ``gaitk_normalize_frames``                  (:212-227)
  * one fp32 frame store (sum N_frames, D) per modality stays in HBM for the whole fold; a window is just an int64
    start row (``gaitk_window_indices`` :230-237), so a batch is an index tensor that the fused stream kernels (or
    ``gaitk_window_gather`` for the dense (B, T, D) batches the unmodified trainer expects) read straight from HBM
  * the sync window intersection and the async without-replacement permutations are integer work restated exactly
    (:278-299, :311-334), and the batch order comes from torch's own DataLoader sampler machinery driven by the same
    ``torch.Generator`` protocol as the reference loaders (:420-455), so a seed gives the same batches.

Input: ``prepare_split(train, test, data_dir=...)`` reads the reference's on-disk format -- per subject three pickled
DataFrames ``<sid>_{walkway,insole,imu}.pkl`` at 30 Hz with tuple-packed accelerations (``load_subject_frames``, restating
:141-180) -- or takes ``frames[sid][modality]`` = float64 (N_frames, D) directly, in the reference's fixed column order
(WALKWAY_FIXED / INSOLE_FIXED / IMU_FIXED), NaN where a value or a whole column is missing.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

from pathlib import Path

DEFAULT_DATA_DIR = Path("data/WearGait/WearGait_preproc_SPmT_30Hz")      # dataloader_weargait.py:27
DEFAULT_MODALITIES = ("walkway", "insole", "imu")
# fixed channel order of the three streams (:29-50)
IMU_SITES = ["L_Ankle", "R_Ankle", "L_DorsalFoot", "R_DorsalFoot", "L_MidLatThigh", "R_MidLatThigh", "L_LatShank", "R_LatShank"]
WALKWAY_FIXED = ["L Foot Pressure_BW", "R Foot Pressure_BW"]
INSOLE_FIXED = ["LTotalForce_BW", "RTotalForce_BW", "SumForce_BW", "LCoP_X", "LCoP_Y", "RCoP_X", "RCoP_Y",
                "Linsole_Acc_X", "Linsole_Acc_Y", "Linsole_Acc_Z", "Rinsole_Acc_X", "Rinsole_Acc_Y", "Rinsole_Acc_Z"]
IMU_FIXED = [f"{s}_FreeAcc_{ax}" for s in IMU_SITES for ax in ("E", "N", "U")]
MODALITY_DIM = {"walkway": 2, "insole": 13, "imu": 24}
NORMALISED = ("insole", "imu")                  # walkway is used as is (build_windows_per_subject :248-253)
MIN_STD = 1e-6


# ---------------------------------------------------------------------------------------------- integer helpers
def window_indices(n_frames: int, win: int, hop: int) -> List[Tuple[int, int, int]]:
    """dataloader_weargait.py:230-237: strict full windows (id, start, stop)."""
    L = _lib.lib()
    n = int(L.gaitk_window_indices(int(n_frames), int(win), int(hop), None, 0))
    if n <= 0:
        return []
    buf = (C.c_int64 * (3 * n))()
    L.gaitk_window_indices(int(n_frames), int(win), int(hop), buf, n)
    return [(int(buf[3 * i]), int(buf[3 * i + 1]), int(buf[3 * i + 2])) for i in range(n)]


def count_windows(n_frames: int, win: int, hop: int) -> int:
    return max(int(_lib.lib().gaitk_window_indices(int(n_frames), int(win), int(hop), None, 0)), 0)


def build_subj2label(pd_ids: Sequence[str], hc_ids: Sequence[str]) -> Dict[str, int]:
    """:56-58 (PD = 1, HC = 0)."""
    return {**{s: 1 for s in pd_ids}, **{s: 0 for s in hc_ids}}


def make_fixed_balanced_folds_no_overlap(pd_ids, hc_ids, n_folds=10, per_class=8, seed=0):
    """:60-74: disjoint balanced test sets drawn with random.Random(seed)."""
    if len(pd_ids) < n_folds * per_class or len(hc_ids) < n_folds * per_class:
        raise AssertionError("Not enough subjects.")
    rng = random.Random(seed)
    pd_pool = list(pd_ids); hc_pool = list(hc_ids)
    rng.shuffle(pd_pool); rng.shuffle(hc_pool)
    used_pd = pd_pool[:n_folds * per_class]; used_hc = hc_pool[:n_folds * per_class]
    folds = []
    for f in range(n_folds):
        te = sorted(used_pd[f * per_class:(f + 1) * per_class]) + sorted(used_hc[f * per_class:(f + 1) * per_class])
        tr = sorted(s for s in (list(pd_ids) + list(hc_ids)) if s not in te)
        folds.append((tr, te))
    return folds


def _subj_from_key(k: str) -> str:
    return k.split("|", 1)[0]


# ---------------------------------------------------------------------------------------------- frame stores
class WindowStore(Mapping):
    """One modality of one split: the normalised fp32 frames of all its subjects in ONE device tensor plus the
    window table.  Behaves like the reference's ``{"sid|mod|wid": ndarray(T, D)}`` dict for the few places that
    index it by key (slow path: one small D2H copy per lookup)."""

    def __init__(self, modality: str, frames: torch.Tensor, win: int, keys: List[str], starts: np.ndarray):
        self.modality = modality
        self.frames = frames                    # (sum N_frames, D) fp32, device
        self.win = int(win)
        self._keys = keys                       # insertion order: subject order, window id ascending
        self.starts = np.asarray(starts, dtype=np.int64)       # global start row of every window
        self._pos = {k: i for i, k in enumerate(keys)}
        self._starts_dev: Optional[torch.Tensor] = None

    @property
    def dim(self) -> int:
        return int(self.frames.shape[1])

    def starts_device(self) -> torch.Tensor:
        if self._starts_dev is None:
            self._starts_dev = torch.from_numpy(self.starts).to(self.frames.device)
        return self._starts_dev

    def __len__(self): return len(self._keys)
    def __iter__(self): return iter(self._keys)
    def __contains__(self, k): return k in self._pos
    def position(self, k: str) -> int: return self._pos[k]

    def __getitem__(self, k: str) -> np.ndarray:
        s = int(self.starts[self._pos[k]])
        return self.frames[s:s + self.win].cpu().numpy()

    def gather(self, starts: torch.Tensor, enabled: bool = True) -> torch.Tensor:
        """(B,) int64 device window starts -> dense (B, T, D) fp32 (zeros when not ``enabled``: _maybe_zero)."""
        B = int(starts.numel())
        out = torch.empty(B, self.win, self.dim, dtype=torch.float32, device=self.frames.device)
        if B:
            _lib.check(_lib.lib().gaitk_window_gather(self.frames.data_ptr(), self.dim, starts.data_ptr(), B, self.win,
                                                      1 if enabled else 0, out.data_ptr(), _lib.stream_handle()),
                       "gaitk_window_gather")
        return out


# ---------------------------------------------------------------------------------------------- on-disk format (30 Hz PKLs)
def _column(df, name: str, n: int) -> np.ndarray:
    """one fixed column as float64 the way expand_* / ensure_cols leave it (:76-91 without statistics): an absent column and a
    column without a single finite value are 0.0 (so they DO enter the train statistics as zeros and z-score to (0 - mean) / std,
    exactly as in the reference); non-numeric cells -> NaN (pd.to_numeric(errors='coerce')); isolated NaNs stay NaN and become the
    train mean on the device (apply_stats :217)."""
    import pandas as pd
    if name not in df.columns:
        return np.zeros(n, dtype=np.float64)
    x = pd.to_numeric(df[name], errors="coerce").to_numpy(dtype=float)
    return x if np.isfinite(x).any() or not n else np.zeros(n, dtype=np.float64)


def load_subject_frames(data_dir, sid: str) -> Dict[str, np.ndarray]:
    """load_subject_streams + expand_insole / expand_imu + ensure_cols (:76-91, :141-180) for one subject:
    ``<sid.lower()>_{walkway,insole,imu}.pkl`` -> {modality: float64 (N_frames, D)} in the fixed column order.  A missing file
    gives zero frames; tuple columns (``Linsole_Acc``, ``<site>_FreeAcc``) are unpacked into their three axes; a missing or
    entirely non-finite column is 0.0 (see _column)."""
    import pandas as pd
    out = {}
    for m, fixed, tuples in (("walkway", WALKWAY_FIXED, ()), ("insole", INSOLE_FIXED, (("Linsole_Acc", ("X", "Y", "Z")), ("Rinsole_Acc", ("X", "Y", "Z")))),
                             ("imu", IMU_FIXED, tuple((f"{s_}_FreeAcc", ("E", "N", "U")) for s_ in IMU_SITES))):
        path = Path(data_dir) / f"{sid.lower()}_{m}.pkl"
        df = pd.read_pickle(path) if path.exists() else pd.DataFrame()
        n = len(df)
        if n:
            df = df.copy()
            for col, axes in tuples:
                if col in df.columns:
                    arr = np.vstack([np.asarray(t, dtype=float) for t in df[col].astype(object)])
                    for i, ax in enumerate(axes):
                        df[f"{col}_{ax}"] = arr[:, i]
                    df = df.drop(columns=[col])
        X = np.empty((n, len(fixed)), dtype=np.float64)
        for j, c in enumerate(fixed):
            X[:, j] = _column(df, c, n)
        out[m] = X
    return out


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def fit_stats_on_train(train_subjects: Sequence[str], frames, device="cuda") -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """:181-210 on the device: per-channel mean / std (population, floored at MIN_STD) over the finite train frames,
    float64 throughout.  Returns {modality: (mean[D], std[D])} as float64 device tensors."""
    L = _lib.lib(); st = _lib.stream_handle
    out = {}
    for m in NORMALISED:
        D = MODALITY_DIM[m]
        acc = torch.zeros(3 * D, dtype=torch.float64, device=device)
        for sid in train_subjects:
            x = torch.from_numpy(_as_f64(frames[sid][m])).to(device)
            if x.numel() == 0:
                continue
            if x.shape[1] != D:
                raise _lib.GaitkError(f"{sid}/{m}: expected {D} columns, got {x.shape[1]}")
            _lib.check(L.gaitk_stats_accumulate(x.data_ptr(), x.shape[0], D, acc.data_ptr(), st()), "gaitk_stats_accumulate")
        mean = torch.empty(D, dtype=torch.float64, device=device); std = torch.empty_like(mean)
        _lib.check(L.gaitk_stats_finalize(acc.data_ptr(), D, mean.data_ptr(), std.data_ptr(), st()), "gaitk_stats_finalize")
        out[m] = (mean, std)
    return out


def build_store(subjects: Sequence[str], frames, modality: str, stats, win: int, hop: int, device="cuda") -> WindowStore:
    """build_windows_per_subject :239-275 for one modality over a list of subjects."""
    L = _lib.lib(); st = _lib.stream_handle
    D = MODALITY_DIM[modality]
    chunks, keys, starts, base = [], [], [], 0
    for sid in subjects:
        x64 = _as_f64(frames[sid][modality])
        n = x64.shape[0]
        if n and x64.shape[1] != D:
            raise _lib.GaitkError(f"{sid}/{modality}: expected {D} columns, got {x64.shape[1]}")
        if n == 0:
            continue
        if modality in NORMALISED:
            xd = torch.from_numpy(x64).to(device)
            o = torch.empty(n, D, dtype=torch.float32, device=device)
            mean, std = stats[modality]
            _lib.check(L.gaitk_normalize_frames(xd.data_ptr(), n, D, mean.data_ptr(), std.data_ptr(), o.data_ptr(), st()),
                       "gaitk_normalize_frames")
        else:
            o = torch.from_numpy(x64.astype(np.float32)).to(device)     # .astype(np.float32) of __getitem__ :343/:361
        chunks.append(o)
        for wid, s0, _ in window_indices(n, win, hop):
            keys.append(f"{sid}|{modality}|{wid}"); starts.append(base + s0)
        base += n
    fr = torch.cat(chunks) if chunks else torch.zeros(0, D, dtype=torch.float32, device=device)
    return WindowStore(modality, fr.contiguous(), win, keys, np.asarray(starts, dtype=np.int64))


def build_sync_pairs(stores: Mapping[str, WindowStore], subjects: Sequence[str], modalities) -> List[Tuple[str, ...]]:
    """_build_index_maps :278-299 (sync half): per subject the window ids present in EVERY modality, numeric order."""
    per = {m: {} for m in modalities}
    for m in modalities:
        for k in stores[m]:
            sid, _, wid = k.split("|")
            per[m].setdefault(sid, set()).add(wid)
    pairs = []
    for sid in subjects:
        sets = [per[m].get(sid, set()) for m in modalities]
        if not all(sets):
            continue
        for wid in sorted(set.intersection(*sets), key=lambda x: int(x)):
            pairs.append(tuple(f"{sid}|{m}|{wid}" for m in modalities))
    return pairs


def prepare_split(train_subs: Sequence[str], test_subs: Sequence[str], *, data_dir=DEFAULT_DATA_DIR, win: int = 64, hop: int = 64,
                  modalities: Tuple[str, ...] = DEFAULT_MODALITIES, frames=None, device="cuda") -> Dict[str, Any]:
    """:388-418 (same call signature): statistics on train only, normalise and window train + test, build the sync index.
    Same keys in the returned dict as the reference; the stores are device-resident ``WindowStore`` objects.  ``frames``
    (optional) bypasses the PKL directory."""
    train_subs = list(train_subs); test_subs = list(test_subs)
    if frames is None:
        frames = {sid: load_subject_frames(data_dir, sid) for sid in dict.fromkeys(train_subs + test_subs)}
    stats = fit_stats_on_train(train_subs, frames, device)
    train_stores = {m: build_store(train_subs, frames, m, stats, win, hop, device) for m in modalities}
    test_stores = {m: build_store(test_subs, frames, m, stats, win, hop, device) for m in modalities}
    return {"train_subs": train_subs, "test_subs": test_subs, "stats": stats,
            "train_stores": train_stores, "test_stores": test_stores,
            "train_sync": build_sync_pairs(train_stores, train_subs, modalities),
            "test_sync": build_sync_pairs(test_stores, test_subs, modalities)}


# ---------------------------------------------------------------------------------------------- datasets (index only)
class WearGaitSyncDataset:
    """:351-363: item i = the aligned windows pairs[i] of every modality + the subject's label."""

    def __init__(self, stores: Tuple[WindowStore, ...], pairs: List[Tuple[str, ...]], subj2label: Dict[str, int]):
        self.stores = tuple(stores); self.pairs = pairs; self.subj2label = subj2label
        self.modalities = tuple(s.modality for s in self.stores)
        n = len(pairs)
        self._starts = [np.fromiter((st.starts[st.position(p[j])] for p in pairs), dtype=np.int64, count=n)
                        for j, st in enumerate(self.stores)]
        self._y = np.fromiter((subj2label[_subj_from_key(p[0])] for p in pairs), dtype=np.int64, count=n)
        self._dev = None

    def __len__(self): return len(self.pairs)

    def tables(self):
        """device copies: ([starts per modality], [labels per modality])"""
        if self._dev is None:
            d = self.stores[0].frames.device
            y = torch.from_numpy(self._y).to(d)
            self._dev = ([torch.from_numpy(s).to(d) for s in self._starts], [y] * len(self.stores))
        return self._dev

    def keys_of(self, idx: Sequence[int]):
        return [self.pairs[i] for i in idx]

    def __getitem__(self, i):                         # reference item (slow path, host copies)
        ks = self.pairs[i]
        xs = [torch.from_numpy(self.stores[j][ks[j]]) for j in range(len(self.stores))]
        return {"xs": xs, "keys": ks, "y": torch.tensor(self.subj2label[_subj_from_key(ks[0])], dtype=torch.long)}


class WearGaitMultiAsyncDataset:
    """:305-348: independent without-replacement permutation per modality (random.Random(seed).shuffle over the
    string-sorted keys), epoch length = the shortest modality, per-modality labels."""

    def __init__(self, stores: Mapping[str, WindowStore], modalities: Tuple[str, ...], subj2label: Dict[str, int], seed: int = 0):
        self.modalities = tuple(modalities); self.stores = stores; self.subj2label = subj2label
        self._keys_full = {m: sorted(stores[m].keys()) for m in self.modalities}
        self._lens_full = {m: len(self._keys_full[m]) for m in self.modalities}
        self._min_len = min(self._lens_full.values())
        # tables in sorted-key order
        self._starts_sorted = {m: np.fromiter((stores[m].starts[stores[m].position(k)] for k in self._keys_full[m]),
                                              dtype=np.int64, count=self._lens_full[m]) for m in self.modalities}
        self._y_sorted = {m: np.fromiter((subj2label[_subj_from_key(k)] for k in self._keys_full[m]), dtype=np.int64,
                                         count=self._lens_full[m]) for m in self.modalities}
        self._perms: Dict[str, List[int]] = {}
        self.reseed(seed)

    def reseed(self, seed: int):
        """:329-334 (the constructor draws the same way, :320-327)."""
        self._rng = random.Random(seed)
        for m in self.modalities:
            idxs = list(range(self._lens_full[m]))
            self._rng.shuffle(idxs)
            self._perms[m] = idxs[:self._min_len]
        self._dev = None

    def __len__(self): return self._min_len

    def tables(self):
        if self._dev is None:
            d = self.stores[self.modalities[0]].frames.device
            starts, ys = [], []
            for m in self.modalities:
                p = np.asarray(self._perms[m], dtype=np.int64)
                starts.append(torch.from_numpy(self._starts_sorted[m][p]).to(d))
                ys.append(torch.from_numpy(self._y_sorted[m][p]).to(d))
            self._dev = (starts, ys)
        return self._dev

    def keys_of(self, idx: Sequence[int]):
        return {m: [self._keys_full[m][self._perms[m][i]] for i in idx] for m in self.modalities}

    def __getitem__(self, idx):
        out = {"keys": {}, "y": {}}
        for m in self.modalities:
            k = self._keys_full[m][self._perms[m][idx]]
            out[m] = torch.from_numpy(self.stores[m][k]); out["keys"][m] = k
            out["y"][m] = torch.tensor(self.subj2label[_subj_from_key(k)], dtype=torch.long)
        return out


# ---------------------------------------------------------------------------------------------- loaders
class IndexBatch:
    """One batch as indices into the resident stores: what the fused step consumes (``win_start=``)."""
    __slots__ = ("frames", "win_start", "ys", "index")

    def __init__(self, frames, win_start, ys, index):
        self.frames = frames; self.win_start = win_start; self.ys = ys; self.index = index

    def __len__(self): return int(self.index.numel())


class _Indices:
    """the integers 0..len(dataset)-1 as a map-style dataset (batched fetch), for the order-producing DataLoader"""

    def __init__(self, dataset): self.dataset = dataset
    def __len__(self): return len(self.dataset)
    def __getitem__(self, i): return int(i)
    def __getitems__(self, idx): return list(idx)


class DeviceLoader:
    """Stands where the reference's ``DataLoader`` stands (:420-455).  The batch ORDER is produced by a real
    ``torch.utils.data.DataLoader`` over the integers 0..len-1 with the same ``batch_size / shuffle / generator``
    arguments, so the generator is consumed exactly as by the reference loaders (one base-seed draw per pass, then the
    sampler's randperm); the batch CONTENT never leaves the device.

    Iterating yields the reference's collated dict (sync: ``{"xs": [..], "y": y, "keys": [...]}``; async:
    ``{m: x_m, "y": {m: y_m}, "keys": {m: [...]}}``) with dense device tensors, so the unmodified trainer loops work
    (their ``.to(DEVICE)`` calls are no-ops).  ``index_batches()`` yields ``IndexBatch`` objects instead -- no dense
    batch is materialised, the fused stream kernels read the windows from the stores.
    """

    def __init__(self, dataset, batch_size: int, shuffle: bool, generator: torch.Generator, with_keys: bool = False):
        self.dataset = dataset; self.batch_size = int(batch_size); self.shuffle = bool(shuffle); self.generator = generator
        self.with_keys = with_keys
        self._order = torch.utils.data.DataLoader(_Indices(dataset), batch_size=self.batch_size, shuffle=self.shuffle, num_workers=0,
                                                  generator=generator, collate_fn=lambda idx: idx)

    def __len__(self): return len(self._order)

    def _index_lists(self):
        return list(self._order)                      # consumes the generator exactly like one reference pass

    def index_batches(self):
        lists = self._index_lists()
        starts, ys = self.dataset.tables()
        dev = starts[0].device
        flat = torch.tensor([i for b in lists for i in b], dtype=torch.int64).to(dev) if lists else torch.zeros(0, dtype=torch.int64, device=dev)
        if isinstance(self.dataset, WearGaitSyncDataset):
            frames = [s.frames for s in self.dataset.stores]
        else:
            frames = [self.dataset.stores[m].frames for m in self.dataset.modalities]
        same_y = all(y is ys[0] for y in ys)
        o = 0
        for b in lists:
            ix = flat[o:o + len(b)]; o += len(b)
            ws = [s.index_select(0, ix) for s in starts]
            if same_y:
                y0 = ys[0].index_select(0, ix); yy = [y0] * len(ys)
            else:
                yy = [y.index_select(0, ix) for y in ys]
            yield IndexBatch(frames, ws, yy, ix)

    def __iter__(self):
        ds = self.dataset
        sync = isinstance(ds, WearGaitSyncDataset)
        stores = list(ds.stores) if sync else [ds.stores[m] for m in ds.modalities]
        for ib in self.index_batches():
            dense = [st.gather(w) for st, w in zip(stores, ib.win_start)]
            if sync:
                out = {"xs": dense, "y": ib.ys[0]}
                out["keys"] = ds.keys_of(ib.index.tolist()) if self.with_keys else None
            else:
                out = {m: x for m, x in zip(ds.modalities, dense)}
                out["y"] = {m: y for m, y in zip(ds.modalities, ib.ys)}
                out["keys"] = ds.keys_of(ib.index.tolist()) if self.with_keys else None
            yield out


def make_sync_loaders(prep: Dict[str, Any], subj2label: Dict[str, int], *, batch_size=64, num_workers=4, seed=0,
                      modalities: Tuple[str, ...] = DEFAULT_MODALITIES, with_keys: bool = False):
    """:420-434.  ``num_workers`` is accepted for signature parity; there are no workers (nothing is copied)."""
    g = torch.Generator().manual_seed(seed)
    train_ds = WearGaitSyncDataset(tuple(prep["train_stores"][m] for m in modalities), prep["train_sync"], subj2label)
    test_ds = WearGaitSyncDataset(tuple(prep["test_stores"][m] for m in modalities), prep["test_sync"], subj2label)
    return (DeviceLoader(train_ds, batch_size, True, g, with_keys), DeviceLoader(test_ds, batch_size, False, g, with_keys))


def make_async_loaders(prep: Dict[str, Any], subj2label: Dict[str, int], *, batch_size=64, num_workers=4, seed=0,
                       modalities: Tuple[str, ...] = DEFAULT_MODALITIES, with_keys: bool = False):
    """:436-455."""
    g = torch.Generator().manual_seed(seed)
    train_ds = WearGaitMultiAsyncDataset(prep["train_stores"], modalities, subj2label, seed=seed)
    test_ds = WearGaitMultiAsyncDataset(prep["test_stores"], modalities, subj2label, seed=seed + 1)
    return (DeviceLoader(train_ds, batch_size, True, g, with_keys), DeviceLoader(test_ds, batch_size, False, g, with_keys))
