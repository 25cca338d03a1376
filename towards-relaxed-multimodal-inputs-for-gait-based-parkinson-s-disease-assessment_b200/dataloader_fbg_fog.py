"""Device-resident FoG / FBG data path: drop-in for ``train/data_processing/dataloader_fbg_fog.py`` (same names, same call
signatures, same batch dicts, same sample ORDER for a seed).

The reference keeps every clip as a host ndarray, pads / casts it per ``__getitem__`` and ships batches through DataLoader
workers + pinned H2D copies (dataloader_fbg_fog.py:124-257, 476-492).  Here a fold is prepared ONCE on the device:

  * ``ClipStore``: all clips of one modality, centred / min-max normalised / padded by ``gaitk_fog_prepare_pose`` (:93-113, 24-37
    fused, bit exact) or padded by ``gaitk_fog_prepare_sensor``, resident in HBM as one ``(n_clips * T, D)`` fp32 frame store;
  * the datasets are KEY LISTS (what the reference's oversampling / wrap-around logic manipulates, :53-90, 170-257, 375-470) --
    restated with the same draws from Python's ``random`` module in the same order, so a seed gives the same lists;
  * ``FogDeviceLoader`` takes the batch order from a real ``torch.utils.data.DataLoader`` over the integers (the reference's
    generator protocol), turns it into row indices on the device and either gathers dense batches (``gaitk_window_gather``,
    what an unmodified ``process_batch`` consumes) or hands ``(frame store, start rows, labels)`` to the fused step, whose
    stream kernels read the clips in place (``index_batches``).
"""
from __future__ import annotations

import random
from collections import defaultdict
from typing import Any, Callable, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

DEFAULT_SKELETON_LEN = 101
DEFAULT_SENSOR_LEN = 65
NUM_CLASSES = 3
MIN_STD = 1e-4
_FOG_EXCLUDED = ("SUB10", "SUB30", "SUB22")            # dataloader_fbg_fog.py:317
_DATASET_ALIASES = {"fbg": "fbg", "fog": "fog", "walk": "fbg", "turn": "fog"}     # configs.py:35-43


def normalize_dataset_name(name: str) -> str:
    key = str(name).lower()
    if key not in _DATASET_ALIASES:
        raise ValueError(f"Unknown dataset: {name}")
    return _DATASET_ALIASES[key]


# ---------------------------------------------------------------------------------------------- host helpers (:24-113)
def pad_or_trim(seq: np.ndarray, target_len: int, pad_value: float = 0.0) -> np.ndarray:
    """:24-37 -- exactly ``target_len`` frames: cut the tail or append ``pad_value`` frames."""
    n = seq.shape[0]
    if n >= target_len:
        return seq if n == target_len else seq[:target_len]
    tail = np.full((target_len - n,) + tuple(seq.shape[1:]), pad_value, dtype=seq.dtype)
    return np.concatenate([seq, tail], axis=0)


def compute_class_weights(counts: List[int]) -> torch.Tensor:
    """:39-43 -- inverse frequencies scaled to sum to NUM_CLASSES."""
    inv = 1.0 / (torch.tensor(counts, dtype=torch.float32) + 1e-8)
    return inv / inv.sum() * NUM_CLASSES


def _tail2(key: str) -> str:
    return "_".join(key.split("_")[-2:])


def _head2(key: str) -> str:
    return "_".join(key.split("_")[:2])


def group_by_subject(keys: List[str]) -> Dict[str, List[str]]:
    """:45-51 -- first token of the key -> keys, in first-seen order."""
    groups: Dict[str, List[str]] = defaultdict(list)
    for k in keys:
        groups[k.split("_")[0]].append(k)
    return groups


def build_synced_pairs(pose_map: Dict[str, List[str]], sens_map: Dict[str, List[str]]) -> List[Tuple[str, str]]:
    """:53-75 -- per subject (pose order), every (pose key, sensor key) whose last two ``_`` tokens agree."""
    pairs: List[Tuple[str, str]] = []
    for sub, pose_keys in pose_map.items():
        by_segment: Dict[str, List[str]] = defaultdict(list)
        for sk in sens_map.get(sub, []):
            by_segment[_tail2(sk)].append(sk)
        for pk in pose_keys:
            pairs.extend((pk, sk) for sk in by_segment.get(_tail2(pk), ()))
    return pairs


def _balanced_draw(groups: Mapping[Any, list], target: int) -> list:
    """``target`` uniform draws (``random.choice``) from every group, groups in insertion order -- the common core of the
    reference's oversampling blocks (:77-90, :408-416)."""
    picked = []
    for members in groups.values():
        picked.extend(random.choice(members) for _ in range(target))
    return picked


def oversample_equally(pairs: List[Tuple[str, str]], get_label: Callable[[str], int]) -> List[Tuple[str, str]]:
    """:77-90 -- every class as often as the largest one (global ``random`` state), then one shuffle."""
    by_class: Dict[int, list] = defaultdict(list)
    for pair in pairs:
        by_class[get_label(pair[0])].append(pair)
    balanced = _balanced_draw(by_class, max(len(v) for v in by_class.values()))
    random.shuffle(balanced)
    return balanced


def center_poses(pose_dict: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """:93-99 (host version; the device path fuses it into gaitk_fog_prepare_pose)."""
    return {k: a - a[:, 0:1, :] for k, a in pose_dict.items()}


def normalize_poses(pose_dict: Dict[str, np.ndarray], method: str = "minmax") -> Dict[str, np.ndarray]:
    """:101-121 (host version)."""
    if method == "minmax":
        return {k: (a - a.min(axis=(0, 1))) / (a.max(axis=(0, 1)) - a.min(axis=(0, 1)) + 1e-6) for k, a in pose_dict.items()}
    if method == "zscore":
        stack = np.vstack(list(pose_dict.values()))
        mean, std = stack.mean(axis=0), stack.std(axis=0)
        std[std < MIN_STD] = 1.0
        return {k: (a - mean) / std for k, a in pose_dict.items()}
    return pose_dict


# ---------------------------------------------------------------------------------------------- device clip store
class ClipStore(Mapping):
    """All clips of one modality, prepared on the device: ``frames`` (n_clips * T, D) fp32 -- clip i occupies rows
    [i T, (i + 1) T) -- plus key -> clip index.  ``kind='pose'``: raw (L, J, 3) float64 clips are centred on joint 0, min-max
    normalised per coordinate over the whole clip and padded / trimmed (:93-113 + :24-37 in one kernel); ``kind='plain'``:
    (L, ...) clips are padded / trimmed and cast.  Built lazily (first device access), so the key logic runs without a GPU.
    As a Mapping it behaves like the reference's ``{key: ndarray(T, ...)}`` dicts (slow path: one D2H copy per lookup)."""

    def __init__(self, clips: Mapping[str, np.ndarray], pad_length: int, kind: str = "plain", device="cuda"):
        assert kind in ("pose", "plain")
        self._keys = list(clips.keys()); self._pos = {k: i for i, k in enumerate(self._keys)}
        self._clips = clips; self.T = int(pad_length); self.kind = kind; self.device = torch.device(device)
        first = next(iter(clips.values()), None)
        self.item_shape = tuple(first.shape[1:]) if first is not None else ()
        self.D = int(np.prod(self.item_shape)) if first is not None else 0
        self._frames: Optional[torch.Tensor] = None

    def __len__(self): return len(self._keys)
    def __iter__(self): return iter(self._keys)
    def __contains__(self, k): return k in self._pos
    def position(self, k: str) -> int: return self._pos[k]

    @property
    def frames(self) -> torch.Tensor:
        if self._frames is None:
            n = len(self._keys)
            out = torch.empty(max(n, 1) * self.T, max(self.D, 1), dtype=torch.float32, device=self.device)
            if n:
                lens = np.fromiter((self._clips[k].shape[0] for k in self._keys), dtype=np.int64, count=n)
                starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
                host = np.concatenate([np.ascontiguousarray(self._clips[k], dtype=np.float64).reshape(-1, self.D) for k in self._keys])
                d_in = torch.from_numpy(host).to(self.device)
                d_st = torch.from_numpy(starts).to(self.device); d_ln = torch.from_numpy(lens).to(self.device)
                L = _lib.lib(); st = _lib.stream_handle()
                if self.kind == "pose":
                    assert self.item_shape[-1] == 3, "pose clips are (L, J, 3)"
                    _lib.check(L.gaitk_fog_prepare_pose(d_in.data_ptr(), d_st.data_ptr(), d_ln.data_ptr(), n, self.D // 3, self.T,
                                                        out.data_ptr(), st), "gaitk_fog_prepare_pose")
                else:
                    _lib.check(L.gaitk_fog_prepare_sensor(d_in.data_ptr(), d_st.data_ptr(), d_ln.data_ptr(), n, self.D, self.T,
                                                          out.data_ptr(), st), "gaitk_fog_prepare_sensor")
                torch.cuda.current_stream().synchronize()       # the temporaries above die with this scope
            self._frames = out
            self._clips = None                                  # the host copies are no longer needed
        return self._frames

    def __getitem__(self, k: str) -> np.ndarray:
        i = self._pos[k]
        return self.frames[i * self.T:(i + 1) * self.T].cpu().numpy().reshape((self.T,) + self.item_shape)

    def gather(self, rows: torch.Tensor) -> torch.Tensor:
        """(B,) int64 device clip indices -> dense (B, T, *item_shape) fp32"""
        B = int(rows.numel())
        out = torch.empty((B, self.T) + self.item_shape, dtype=torch.float32, device=self.device)
        if B:
            starts = rows * self.T
            _lib.check(_lib.lib().gaitk_window_gather(self.frames.data_ptr(), self.D, starts.data_ptr(), B, self.T, 1, out.data_ptr(),
                                                      _lib.stream_handle()), "gaitk_window_gather")
        return out


def _select(keys, selected_subjects, all_if_empty: bool) -> List[str]:
    """:131-134 / :154-157 -- the skeleton dataset takes every key only for ``None``, the sensor dataset for any empty selection"""
    if selected_subjects is None or (all_if_empty and not selected_subjects):
        return list(keys)
    return [k for k in keys if any(k.startswith(s) for s in selected_subjects)]


class _ClipDataset:
    """:124-170 -- one modality: the selected keys (a plain, replaceable list, as in the reference) over a ClipStore"""
    _all_if_empty = False

    def __init__(self, clip_dict, selected_subjects: Optional[List[str]], pad_length: int):
        self.keys = _select(clip_dict.keys(), selected_subjects, self._all_if_empty)
        if isinstance(clip_dict, ClipStore):
            assert clip_dict.T == pad_length
            self.store = clip_dict
        else:                                          # already centred / normalised host clips: pad + cast only
            self.store = ClipStore({k: clip_dict[k] for k in self.keys}, pad_length, "plain")
        self.pad_length = pad_length

    def __len__(self): return len(self.keys)

    def __getitem__(self, idx):
        key = self.keys[idx]
        return self.store[key], key

    def rows(self) -> np.ndarray:
        return np.fromiter((self.store.position(k) for k in self.keys), dtype=np.int64, count=len(self.keys))


class SkeletonDataset(_ClipDataset):
    def __init__(self, pose_dict, selected_subjects: List[str], pad_length: int = DEFAULT_SKELETON_LEN):
        super().__init__(pose_dict, selected_subjects, pad_length)

    @property
    def poses(self): return self.store


class SensorDataset(_ClipDataset):
    _all_if_empty = True

    def __init__(self, sensor_dict, selected_subjects: List[str], pad_length: int = DEFAULT_SENSOR_LEN):
        super().__init__(sensor_dict, selected_subjects, pad_length)

    @property
    def sensors(self): return self.store


class FusionDataset:
    """:170-257 -- synchronised: item i = the i-th (pose, sensor) pair (class-balanced by oversampling when ``seed`` is given);
    asynchronous: item i = pose ``i % n_pose`` with sensor ``i % n_sens``.  Labels: FBG from the per-key maps, FoG from the subject
    map.  ``tables()`` turns the CURRENT key lists into device index / label tensors (the reference code replaces
    ``pose_ds.keys`` / ``sens_ds.keys`` after construction, :366-470, so nothing is cached across replacements)."""

    def __init__(self, pose_dict, sensor_dict, subject_label_map: Dict[str, int] = None, pose_label_map: Dict[str, int] = None,
                 sensor_label_map: Dict[str, int] = None, selected_subjects: List[str] = None, synchronized: bool = False,
                 seed: int = 0, pad_skel: int = DEFAULT_SKELETON_LEN, pad_sens: int = DEFAULT_SENSOR_LEN):
        self.pose_ds = SkeletonDataset(pose_dict, selected_subjects, pad_skel)
        self.sens_ds = SensorDataset(sensor_dict, selected_subjects, pad_sens)
        self.synchronized = synchronized
        self.subject_label_map = subject_label_map; self.pose_label_map = pose_label_map; self.sensor_label_map = sensor_label_map
        if synchronized:
            pairs = build_synced_pairs(group_by_subject(self.pose_ds.keys), group_by_subject(self.sens_ds.keys))
            if seed is not None:
                random.seed(seed)
                pairs = oversample_equally(pairs, self._pose_label)
            self.pairs = pairs

    def _pose_label(self, pk: str) -> int:
        return self.pose_label_map[_head2(pk)] if self.pose_label_map is not None else self.subject_label_map[pk.split("_")[0]]

    def _sens_label(self, sk: str) -> int:
        return self.sensor_label_map[sk] if self.sensor_label_map is not None else self.subject_label_map[sk.split("_")[0]]

    def __len__(self):
        return len(self.pairs) if self.synchronized else max(len(self.pose_ds), len(self.sens_ds))

    def key_pair(self, idx: int) -> Tuple[str, str]:
        if self.synchronized:
            return self.pairs[idx]
        return self.pose_ds.keys[idx % len(self.pose_ds)], self.sens_ds.keys[idx % len(self.sens_ds)]

    def __getitem__(self, idx):                       # reference item (slow path, host copies)
        pk, sk = self.key_pair(idx)
        return {"skeleton": torch.from_numpy(self.pose_ds.store[pk]), "sensor": torch.from_numpy(self.sens_ds.store[sk]),
                "label_skeleton": torch.tensor(self._pose_label(pk), dtype=torch.long),
                "label_sensor": torch.tensor(self._sens_label(sk), dtype=torch.long)}

    def index_tables(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """host: (pose clip index, sensor clip index, pose label, sensor label) per item -- pure key logic, no device"""
        n = len(self)
        kp = [self.key_pair(i) for i in range(n)]
        ps, ss = self.pose_ds.store, self.sens_ds.store
        return (np.fromiter((ps.position(p) for p, _ in kp), dtype=np.int64, count=n),
                np.fromiter((ss.position(s) for _, s in kp), dtype=np.int64, count=n),
                np.fromiter((self._pose_label(p) for p, _ in kp), dtype=np.int64, count=n),
                np.fromiter((self._sens_label(s) for _, s in kp), dtype=np.int64, count=n))

    def tables(self):
        dev = self.pose_ds.store.device
        return tuple(torch.from_numpy(a).to(dev) for a in self.index_tables())


# ---------------------------------------------------------------------------------------------- loaders
class ClipIndexBatch:
    """One batch as rows of the resident clip stores: ``frames[s]`` (n_clips * T_s, D_s) + ``win_start[s]`` int64[B] first rows,
    ``ys[s]`` labels -- what ``FusedTrainStep.step(frames, ys, win_start=...)`` consumes; no dense batch exists."""
    __slots__ = ("frames", "win_start", "ys", "index")

    def __init__(self, frames, win_start, ys, index):
        self.frames = frames; self.win_start = win_start; self.ys = ys; self.index = index

    def __len__(self): return int(self.index.numel())


class _Indices:
    def __init__(self, dataset): self.dataset = dataset
    def __len__(self): return len(self.dataset)
    def __getitem__(self, i): return int(i)
    def __getitems__(self, idx): return list(idx)


class FogDeviceLoader:
    """Stands where the reference's ``DataLoader(FusionDataset, ...)`` stands (:476-492): same ``dataset`` attribute, same
    length, same batches in the same order (the order comes from a real DataLoader over the integers with the same
    ``shuffle`` / ``generator``), tensors already on the device."""

    def __init__(self, dataset: FusionDataset, batch_size: int, shuffle: bool, generator: torch.Generator):
        self.dataset = dataset; self.batch_size = int(batch_size); self.shuffle = bool(shuffle); self.generator = generator
        self._order = torch.utils.data.DataLoader(_Indices(dataset), batch_size=self.batch_size, shuffle=self.shuffle, num_workers=0,
                                                  generator=generator, collate_fn=lambda idx: idx)

    def __len__(self): return len(self._order)

    def index_batches(self):
        lists = list(self._order)                     # consumes the generator exactly like one reference pass
        prow, srow, ys, yt = self.dataset.tables()
        dev = prow.device
        ps, ss = self.dataset.pose_ds.store, self.dataset.sens_ds.store
        flat = torch.tensor([i for b in lists for i in b], dtype=torch.int64).to(dev) if lists else torch.zeros(0, dtype=torch.int64, device=dev)
        o = 0
        for b in lists:
            ix = flat[o:o + len(b)]; o += len(b)
            yield ClipIndexBatch([ps.frames, ss.frames], [prow.index_select(0, ix) * ps.T, srow.index_select(0, ix) * ss.T],
                                 [ys.index_select(0, ix), yt.index_select(0, ix)], ix)

    def __iter__(self):
        ps, ss = self.dataset.pose_ds.store, self.dataset.sens_ds.store
        for ib in self.index_batches():
            yield {"skeleton": ps.gather(ib.win_start[0] // ps.T), "sensor": ss.gather(ib.win_start[1] // ss.T),
                   "label_skeleton": ib.ys[0], "label_sensor": ib.ys[1]}


def create_fusion_loaders(dataset: str, reader: Any, train_subjects: List[str], eval_subjects: List[str], batch_size: int = 32,
                          synchronized: bool = False, seed: int = 0, num_workers: int = 4, pad_skel: int = DEFAULT_SKELETON_LEN,
                          pad_sens: int = DEFAULT_SENSOR_LEN, modality: str = "multimodal", device="cuda"):
    """:269-494, same signature (``num_workers`` is accepted and unused: nothing is copied per batch).  ``reader`` carries the
    RAW clip dicts; centring / normalisation / padding happen once, on the device, inside the two ``ClipStore``s that the train
    and the eval dataset share.  Every draw from ``random`` / ``random.Random(seed)`` / the torch generator happens in the
    reference's order, so the key lists, pairs and batches are the reference's."""
    dataset = normalize_dataset_name(dataset)
    random.seed(seed)
    train_subs, eval_subs = list(train_subjects), list(eval_subjects)
    if dataset == "fbg":
        subject_labels = None
        pose_labels = dict(reader.pose_label_dict)
        sens_clips, sens_labels = {}, {}
        for key, arr in reader.sensor_dict.items():           # GRF (T, trials, D) -> one clip per trial (:303-313)
            if arr.ndim == 3:
                for i in range(arr.shape[1]):
                    sens_clips[f"{key}_{i}"] = arr[:, i, :]; sens_labels[f"{key}_{i}"] = reader.sensor_label_dict[key]
            else:
                sens_clips[key] = arr; sens_labels[key] = reader.sensor_label_dict[key]
        pose_heads = {_head2(k) for k in reader.pose_dict}; sens_heads = {_head2(k) for k in sens_clips}
        wanted = {"skeleton": lambda s: s in pose_heads, "sensor": lambda s: s in sens_heads}.get(
            modality, lambda s: s in pose_heads or s in sens_heads)
        kept = [s for s in train_subs if wanted(s)]
        if len(kept) != len(train_subs):
            print(f"[WARN] dropping train subjects missing {modality} data: {set(train_subs) - set(kept)}")
        train_subs = kept
    else:
        subject_labels = {s: (l[0] if isinstance(l, (list, tuple)) else int(l)) for s, l in reader.labels_dict.items()
                          if s not in _FOG_EXCLUDED}
        pose_labels = sens_labels = None
        sens_clips = reader.sensor_dict
    pose_store = ClipStore(reader.pose_dict, pad_skel, "pose", device)
    sens_store = ClipStore(sens_clips, pad_sens, "plain", device)

    def dataset_for(subs, ds_seed):
        return FusionDataset(pose_store, sens_store, subject_labels, pose_labels, sens_labels, subs, synchronized=synchronized,
                             seed=ds_seed, pad_skel=pad_skel, pad_sens=pad_sens)
    train_ds = dataset_for(train_subs, None if synchronized else seed)
    eval_ds = dataset_for(eval_subs, seed)

    if modality == "multimodal" and not synchronized:
        # equalise the two training key lists by drawing extras for the shorter one (:366-377)
        pk, sk = train_ds.pose_ds.keys, train_ds.sens_ds.keys
        if len(pk) != len(sk):
            rng = random.Random(seed)
            if len(pk) < len(sk):
                train_ds.pose_ds.keys = pk + rng.choices(pk, k=len(sk) - len(pk))
            else:
                train_ds.sens_ds.keys = sk + rng.choices(sk, k=len(pk) - len(sk))

    if modality in ("skeleton", "sensor"):
        # class-balanced eval keys of the one modality in use (:380-421)
        part = eval_ds.pose_ds if modality == "skeleton" else eval_ds.sens_ds
        label_of = eval_ds._pose_label if modality == "skeleton" else eval_ds._sens_label
        by_class: Dict[int, List[str]] = defaultdict(list)
        for k in part.keys:
            by_class[label_of(k)].append(k)
        keys = _balanced_draw(by_class, max(len(v) for v in by_class.values()))
        random.shuffle(keys)
        part.keys = keys

    if modality == "multimodal" and not synchronized:
        # subject-balanced asynchronous eval: every eval subject `target` times in both modalities (:424-470)
        unit = _head2 if dataset == "fbg" else (lambda k: k.split("_")[0])
        pose_by, sens_by = defaultdict(list), defaultdict(list)
        for k in eval_ds.pose_ds.keys:
            pose_by[unit(k)].append(k)
        for k in eval_ds.sens_ds.keys:
            sens_by[unit(k)].append(k)
        target = max(max(len(pose_by[s]) for s in eval_subs), max(len(sens_by[s]) for s in eval_subs))
        new_pose, new_sens = [], []
        for s in eval_subs:
            gp, gs = pose_by.get(s, []), sens_by.get(s, [])
            if not gp or not gs:
                raise ValueError(f"Subject {s} lacks data for one modality")
            for _ in range(target):                    # the two draws alternate (one random stream feeds both lists)
                new_pose.append(random.choice(gp)); new_sens.append(random.choice(gs))
        random.shuffle(new_pose); random.shuffle(new_sens)
        eval_ds.pose_ds.keys = new_pose; eval_ds.sens_ds.keys = new_sens

    g = torch.Generator()
    g.manual_seed(seed)
    return (FogDeviceLoader(train_ds, batch_size, True, g), FogDeviceLoader(eval_ds, batch_size, False, g))
