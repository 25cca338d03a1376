"""Host-side integer logic of the device-resident data path (dataloader_weargait.py of the package) against the
golden vectors dumped from the reference's own prepare_split / datasets / loaders.  No GPU: the stores hold dummy CPU
frames, only the window tables, index maps, permutations and the loader ORDER are checked."""
import numpy as np
import pytest
import torch

from conftest import load_golden

MODS = ("walkway", "insole", "imu")


@pytest.fixture(scope="module")
def dl():
    import gaitk
    return gaitk.dataloader_weargait


@pytest.fixture(scope="module")
def fold(dl):
    g = load_golden("data_path")
    sids = [str(s) for s in g["sids"]]
    train = [str(s) for s in g["train"]]; test = [str(s) for s in g["test"]]

    def store(subs, m):
        keys, starts, base = [], [], 0
        for s in subs:
            n = g[f"raw/{s}/{m}"].shape[0]
            for wid, s0, _ in dl.window_indices(n, 64, 64):
                keys.append(f"{s}|{m}|{wid}"); starts.append(base + s0)
            base += n
        return dl.WindowStore(m, torch.zeros(base, dl.MODALITY_DIM[m]), 64, keys, np.array(starts, dtype=np.int64))

    prep = {"train_stores": {m: store(train, m) for m in MODS}, "test_stores": {m: store(test, m) for m in MODS}}
    prep["train_sync"] = dl.build_sync_pairs(prep["train_stores"], train, MODS)
    prep["test_sync"] = dl.build_sync_pairs(prep["test_stores"], test, MODS)
    return g, sids, prep, dl.build_subj2label(sids[:3], sids[3:])


def test_window_indices_match_reference(dl):
    g = load_golden("data_path")
    for i, (n, w, h) in enumerate(g["win_cases"]):
        got = np.array(dl.window_indices(int(n), int(w), int(h)), dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(got, g[f"win_{i}"]), (n, w, h)


def test_store_keys_and_sync_pairs_match_reference(fold):
    g, sids, prep, _ = fold
    for split in ("train", "test"):
        for m in MODS:
            assert sorted(prep[f"{split}_stores"][m].keys()) == [str(k) for k in g[f"{split}_keys/{m}"]]
        got = [[p[0].split("|")[0], p[0].split("|")[2]] for p in prep[f"{split}_sync"]]
        assert got == [[str(a), str(b)] for a, b in g[f"{split}_sync"]]
    # the subject with fewer insole frames than one window is dropped from the sync map
    assert not any(p[0].startswith(sids[4] + "|") for p in prep["train_sync"])


def test_async_permutations_match_reference(fold, dl):
    g, sids, prep, s2l = fold
    ds = dl.WearGaitMultiAsyncDataset(prep["train_stores"], MODS, s2l, seed=43)
    assert np.array_equal(np.array([ds._perms[m] for m in MODS]), g["async_perm_seed43"])
    ds.reseed(44)
    assert np.array_equal(np.array([ds._perms[m] for m in MODS]), g["async_perm_seed44"])
    ks = ds.keys_of([3])
    assert [ks[m][0] for m in MODS] == [str(k) for k in g["async_item3_keys"]]
    _, ys = ds.tables()
    assert [int(y[3]) for y in ys] == [int(v) for v in g["async_item3_y"]]


def test_folds_match_reference(fold, dl):
    g, sids, _, _ = fold
    f = dl.make_fixed_balanced_folds_no_overlap(sids[:3], sids[3:], n_folds=1, per_class=1, seed=43)
    assert f[0][1] == [str(s) for s in g["fold0_test"]]
    f3 = dl.make_fixed_balanced_folds_no_overlap(sids[:3], sids[3:], n_folds=3, per_class=1, seed=7)
    assert [[",".join(tr), ",".join(te)] for tr, te in f3] == [[str(a), str(b)] for a, b in g["folds3_seed7"]]


def test_sync_loader_order_matches_reference(fold, dl):
    """two epochs of (train pass, test pass) on loaders that share one generator, as the trainer runs them"""
    g, sids, prep, s2l = fold
    tr, te = dl.make_sync_loaders(prep, s2l, batch_size=4, num_workers=0, seed=43)
    assert len(tr) == 3 and len(te) == 2
    for ep in range(2):
        for nm, ld in (("train", tr), ("test", te)):
            ks, ys = [], []
            for ib in ld.index_batches():
                ks += [p[0].split("|")[0] + "|" + p[0].split("|")[2] for p in ld.dataset.keys_of(ib.index.tolist())]
                ys += ib.ys[0].tolist()
                # the three streams of a sync batch are the same windows
                for j, m in enumerate(MODS):
                    st = ld.dataset.stores[j]
                    want = [st.starts[st.position(p[j])] for p in ld.dataset.keys_of(ib.index.tolist())]
                    assert ib.win_start[j].tolist() == want
            assert ks == [str(k) for k in g[f"loader_sync/{nm}_ep{ep}_keys"]], (ep, nm)
            assert ys == g[f"loader_sync/{nm}_ep{ep}_y"].tolist()


def test_async_loader_order_matches_reference(fold, dl):
    g, sids, prep, s2l = fold
    tr, te = dl.make_async_loaders(prep, s2l, batch_size=4, num_workers=0, seed=43)
    for ep in range(1, 3):
        tr.dataset.reseed(43 + ep)
        for nm, ld in (("train", tr), ("test", te)):
            ks = {m: [] for m in MODS}; ys = {m: [] for m in MODS}
            for ib in ld.index_batches():
                kk = ld.dataset.keys_of(ib.index.tolist())
                for j, m in enumerate(MODS):
                    ks[m] += kk[m]; ys[m] += ib.ys[j].tolist()
                    st = ld.dataset.stores[m]
                    assert ib.win_start[j].tolist() == [st.starts[st.position(k)] for k in kk[m]]
            for m in MODS:
                assert ks[m] == [str(k) for k in g[f"loader_async/{nm}_ep{ep}_keys/{m}"]], (ep, nm, m)
                assert ys[m] == g[f"loader_async/{nm}_ep{ep}_y/{m}"].tolist()


def test_loader_order_equals_a_real_torch_dataloader(dl):
    """independent of the golden file: same generator protocol as torch.utils.data.DataLoader on this torch version"""
    n = 37
    class DS(torch.utils.data.Dataset):
        def __len__(self): return n
        def __getitem__(self, i): return i
    st = dl.WindowStore("walkway", torch.zeros(n * 64, 2), 64, [f"s|walkway|{i}" for i in range(n)], np.arange(n) * 64)
    pairs = [(f"s|walkway|{i}",) for i in range(n)]
    ds = dl.WearGaitSyncDataset((st,), pairs, {"s": 1})
    g1 = torch.Generator().manual_seed(5); g2 = torch.Generator().manual_seed(5)
    ref = torch.utils.data.DataLoader(DS(), batch_size=8, shuffle=True, num_workers=0, generator=g1)
    ref_te = torch.utils.data.DataLoader(DS(), batch_size=8, shuffle=False, num_workers=0, generator=g1)
    mine = dl.DeviceLoader(ds, 8, True, g2); mine_te = dl.DeviceLoader(ds, 8, False, g2)
    for _ in range(3):
        assert [b.tolist() for b in ref] == [ib.index.tolist() for ib in mine.index_batches()]
        assert [b.tolist() for b in ref_te] == [ib.index.tolist() for ib in mine_te.index_batches()]


@pytest.mark.parametrize("n,bs,seed", [(1, 4, 0), (5, 5, 1), (64, 7, 2), (129, 64, 3)])
def test_loader_order_equals_torch_dataloader_for_other_sizes(dl, n, bs, seed):
    class DS(torch.utils.data.Dataset):
        def __len__(self): return n
        def __getitem__(self, i): return i
    st = dl.WindowStore("walkway", torch.zeros(n * 64, 2), 64, [f"s|walkway|{i}" for i in range(n)], np.arange(n) * 64)
    ds = dl.WearGaitSyncDataset((st,), [(f"s|walkway|{i}",) for i in range(n)], {"s": 0})
    g1 = torch.Generator().manual_seed(seed); g2 = torch.Generator().manual_seed(seed)
    ref = torch.utils.data.DataLoader(DS(), batch_size=bs, shuffle=True, num_workers=0, generator=g1)
    mine = dl.DeviceLoader(ds, bs, True, g2)
    assert len(mine) == len(ref)
    for _ in range(2):
        a = [b.tolist() for b in ref]; b = [ib.index.tolist() for ib in mine.index_batches()]
        assert a == b
        assert sorted(i for x in b for i in x) == list(range(n))          # every window exactly once per epoch


def test_mask_table_arithmetic_matches_the_reference_formulas():
    """evaluation.py turns per-batch hit counts into the numbers eval_with_mask / eval_one_epoch return
    (weargait_train.py:322-433): sync = micro accuracy over windows, async = mean over batches of the per-batch
    accuracy ((pred == y).float().mean().item() * 100), macro over the enabled streams."""
    import gaitk
    ev = gaitk.evaluation
    counts = np.array([[3, 2, 1, 3, 2, 1, 3, 2, 1, 3], [1, 1, 1, 1, 1, 1, 1, 0, 2, 1]], dtype=np.int32)   # two batches
    sizes = np.array([4, 3], dtype=np.int64)
    for i, (name, mask) in enumerate(ev.MASK_COMBOS.items()):
        got = ev._mask_result(counts, sizes, False, i, mask)
        assert got == 100.0 * float(counts[:, i].sum()) / 7.0
        res = ev._mask_result(counts, sizes, True, i, mask)
        want = {}
        for s, nm in enumerate(("walkway", "insole", "imu")):
            if mask[s]:
                accs = [torch.tensor([1.0] * int(c) + [0.0] * int(n - c)).mean().item() * 100 for c, n in zip(counts[:, 7 + s], sizes)]
                want[nm] = sum(accs) / 2
        want["macro_enabled"] = sum(want.values()) / len(want)
        assert res == want, (name, res, want)


def test_pkl_reader_matches_the_reference_loader(tmp_path):
    """load_subject_frames (the reference's on-disk format: three pickled 30 Hz DataFrames per subject, tuple-packed
    accelerations) against the reference's own load_subject_streams + expand_* + ensure_cols, when a reference copy exists."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import ref_harness as H
    R = H.load_reference()
    if R is None:
        pytest.skip("no reference copy (oracle/_ref or /root/reference)")
    import pandas as pd
    import gaitk
    DW = R.DW
    sids, _ = H.write_synthetic_weargait(tmp_path, n_per_class=2, seed=3, frames=(80, 160))
    d = tmp_path / "data" / "WearGait" / "WearGait_preproc_SPmT_30Hz"
    # damage one subject the way real recordings are damaged: a missing IMU site, an all-NaN insole column, a NaN run, no walkway
    imu = pd.read_pickle(d / f"{sids[1]}_imu.pkl").drop(columns=["R_LatShank_FreeAcc"]); imu.to_pickle(d / f"{sids[1]}_imu.pkl")
    ins = pd.read_pickle(d / f"{sids[1]}_insole.pkl"); ins["RCoP_Y"] = np.nan; ins.loc[3:9, "LCoP_X"] = np.nan; ins.to_pickle(d / f"{sids[1]}_insole.pkl")
    (d / f"{sids[2]}_walkway.pkl").unlink()
    for sid in sids:
        got = gaitk.dataloader_weargait.load_subject_frames(d, sid)
        st = DW.load_subject_streams(d, sid)
        ref_w = DW.ensure_cols(st["walkway"], DW.WALKWAY_FIXED).to_numpy(dtype=float)
        ref_i = DW.expand_insole(st["insole"]); ref_m = DW.expand_imu(st["imu"])
        assert got["walkway"].shape == ref_w.shape and np.array_equal(got["walkway"], ref_w)
        assert np.array_equal(got["insole"], ref_i.to_numpy(dtype=float), equal_nan=True), sid
        assert np.array_equal(got["imu"], ref_m.to_numpy(dtype=float), equal_nan=True), sid
    assert list(gaitk.dataloader_weargait.INSOLE_FIXED) == list(DW.INSOLE_FIXED) and list(gaitk.dataloader_weargait.IMU_FIXED) == list(DW.IMU_FIXED)


def test_fog_pairing_and_oversampling_match_the_reference():
    """A5: build_synced_pairs / oversample_equally / FusionDataset / create_fusion_loaders key logic (dataloader_fbg_fog.py:53-90,
    170-257, 269-470) against lists the reference itself produced (fog_loaders golden): key lists of both datasets after all the
    oversampling, the synchronised pairs, and the (pose, sensor, labels) index tables -- host logic only, no device."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import ref_harness as H
    import gaitk
    DF = gaitk.dataloader_fbg_fog
    g = load_golden("fog_loaders")
    pm = DF.group_by_subject(["A_x_1", "A_y_2", "B_x_1", "A_x_3"]); sm = DF.group_by_subject(["A_x_1", "A_q_x_1", "B_x_1", "B_z_1", "C_x_1"])
    assert np.array_equal(np.array(DF.build_synced_pairs(pm, sm)), g["pairs_small"])
    for cname, dataset, sync, modality, pad_skel, pad_sens in H.FOG_LOADER_CASES:
        reader, subs = H.synthetic_fog_reader(dataset, seed=5)
        tr_s, ev_s = H.fog_loader_split(subs, dataset)
        tr, ev = DF.create_fusion_loaders(dataset, reader, tr_s, ev_s, batch_size=7, synchronized=sync, seed=43, num_workers=0,
                                          pad_skel=pad_skel, pad_sens=pad_sens, modality=modality)
        for nm, ld in (("train", tr), ("eval", ev)):
            ds = ld.dataset
            assert list(g[f"{cname}/{nm}_pose_keys"]) == ds.pose_ds.keys, (cname, nm)
            assert list(g[f"{cname}/{nm}_sens_keys"]) == ds.sens_ds.keys, (cname, nm)
            assert int(g[f"{cname}/{nm}_len"]) == len(ds)
            if sync:
                assert [tuple(p) for p in g[f"{cname}/{nm}_pairs"]] == [tuple(p) for p in ds.pairs], (cname, nm)
            prow, srow, ys, yt = ds.index_tables()
            assert len(prow) == len(ds) and prow.max() < len(ds.pose_ds.store) and srow.max() < len(ds.sens_ds.store)
            # labels of the items in dataset order == the labels the reference's un-shuffled eval loader delivered
            if nm == "eval":
                assert np.array_equal(ys, g[f"{cname}/eval_ep0/ys"]) and np.array_equal(yt, g[f"{cname}/eval_ep0/yt"]), cname
        assert len(tr) == len(g[f"{cname}/train_ep0/bs"]) and len(ev) == len(g[f"{cname}/eval_ep0/bs"])
    assert np.allclose(DF.compute_class_weights([10, 20, 30]).numpy(), (lambda w: w / w.sum() * 3)(1 / np.array([10, 20, 30.0])), rtol=1e-6)
    a = np.arange(12.0).reshape(6, 2)
    assert DF.pad_or_trim(a, 6) is a and DF.pad_or_trim(a, 4).shape == (4, 2) and np.array_equal(DF.pad_or_trim(a, 8)[6:], np.zeros((2, 2)))


def test_weargait_etl_matches_the_reference(tmp_path):
    """f3: raw CSVs -> 30 Hz PKLs (preprocess_weargait.run_end_to_end, :228-343) against the reference's own ETL on the same synthetic
    recordings -- fold-agnostic (`*_base.pkl`) and with train statistics -- frame by frame, and the direct CSV -> frame-matrix path
    against the PKL round trip through the loader."""
    import contextlib, io, json, sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import ref_harness as H
    if H.load_reference() is None:
        pytest.skip("no reference copy (oracle/_ref or /root/reference)")
    import pandas as pd
    from data_processing import preprocess_weargait as REF
    import gaitk
    from importlib import import_module
    ETL = import_module(gaitk.__name__ + ".preprocess_weargait")
    hc, pdr, hcd, pdd, sids = H.write_synthetic_weargait_csvs(tmp_path, n_per_class=2, seed=4)
    for mode, train in (("base", None), ("fold", [sids[0], sids[2]])):
        outs = {}
        for tag, mod in (("ref", REF), ("new", ETL)):
            o = tmp_path / f"{mode}_{tag}"
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                mod.run_end_to_end(hc, pdr, hcd, pdd, str(o), train_subject_ids=train, segment_len_rows=64)
            outs[tag] = (o, buf.getvalue())
        assert outs["ref"][1] == outs["new"][1]                        # the printed per-subject / total counts
        files = sorted(p.name for p in outs["ref"][0].iterdir())
        assert files == sorted(p.name for p in outs["new"][0].iterdir()) and len(files) >= 3 * len(sids)
        for f in files:
            if f.endswith(".json"):
                a = json.loads((outs["ref"][0] / f).read_text()); b = json.loads((outs["new"][0] / f).read_text())
                assert a.keys() == b.keys() and all(np.allclose(a[k], b[k], rtol=1e-13, atol=0) for k in a)
                continue
            a = pd.read_pickle(outs["ref"][0] / f); b = pd.read_pickle(outs["new"][0] / f)
            assert list(a.columns) == list(b.columns) and len(a) == len(b) > 50, f
            for c in a.columns:
                if a[c].dtype == object:
                    xa = np.array([np.asarray(t, dtype=float) for t in a[c]]); xb = np.array([np.asarray(t, dtype=float) for t in b[c]])
                else:
                    xa = a[c].to_numpy(dtype=float); xb = b[c].to_numpy(dtype=float)
                assert np.allclose(xa, xb, rtol=1e-13, atol=0, equal_nan=True), (f, c)
    # CSV -> frame matrices directly == CSV -> PKL -> loader
    wmap = ETL.build_weight_map(hcd, pdd)
    files = {**ETL.find_subject_files(hc), **ETL.find_subject_files(pdr)}
    d = tmp_path / "base_new"
    for sid in sids:
        direct = ETL.subject_frames(files[sid.lower()], wmap[sid.lower()])
        for nm in ("insole", "imu"):                                    # the loader reads <sid>_<m>.pkl: give it the base tables
            (d / f"{sid.lower()}_{nm}.pkl").write_bytes((d / f"{sid.lower()}_{nm}_base.pkl").read_bytes())
        via = gaitk.dataloader_weargait.load_subject_frames(d, sid)
        for m in ("walkway", "insole", "imu"):
            a = np.nan_to_num(direct[m], nan=0.0) if m != "walkway" else direct[m]
            # the loader maps wholly missing columns to 0 (ensure_cols); isolated NaNs stay NaN in both
            for j in range(via[m].shape[1]):
                col_d, col_v = direct[m][:, j], via[m][:, j]
                if np.isnan(col_d).all():
                    assert (col_v == 0).all()
                else:
                    assert np.array_equal(col_d, col_v, equal_nan=True), (sid, m, j)
