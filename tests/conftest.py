import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import json
    import numpy as np
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "meta" in d:
        d["meta"] = json.loads(str(d["meta"]))
    return d


def sub(d, prefix):
    """entries of a flat golden dict under 'prefix/' with the prefix stripped"""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in d.items() if k.startswith(prefix + "/")}
