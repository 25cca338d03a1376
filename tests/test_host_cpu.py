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
    declared = sorted(set(re.findall(r"\b(gaitk_[a-z_]+)\s*\(", hdr)))
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
