"""oracle/build_ref.py -- stage the UNMODIFIED reference under oracle/_ref/ so that it travels to the GPU box.

Test / measurement infrastructure only (never imported by the product package).  The reference is pure Python, has no
installable package (`train/setup.py` is a bare find_packages() without the model code) and `/root/reference` does not
exist on the GPU box; this recipe is the "install": it copies the .py files of the three directories the hot path
imports from (train/, data/WearGait/, const/) byte for byte into the git-ignored, NOT gpurun-ignored `oracle/_ref/`, and
records their sha256 in `oracle/_ref/MANIFEST.json`.  Nothing under oracle/_ref is ever committed.

    python oracle/build_ref.py            # (re)stage; no-op when /root/reference is absent and _ref exists

Consumers (the only ones): oracle/ref_harness.py -> bench.py --impl reference, bench.py's rank-0 `cuda_reference` leg,
tests/ (trainer-level drop-in tests, oracle cross-checks).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from pathlib import Path

REF = Path(os.environ.get("GAIT_REFERENCE", "/root/reference"))
DST = Path(__file__).resolve().parent / "_ref"
SUBDIRS = ("train", "data/WearGait", "const")


def stage(force: bool = False) -> Path | None:
    if not REF.exists():
        if DST.exists():
            return DST
        print(f"[build_ref] {REF} not present and {DST} not staged: reference arm unavailable", file=sys.stderr)
        return None
    manifest = {}
    if DST.exists():
        shutil.rmtree(DST)
    for sub in SUBDIRS:
        for src in sorted((REF / sub).rglob("*.py")):
            if "__pycache__" in src.parts:
                continue
            rel = src.relative_to(REF)
            out = DST / rel
            out.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(src, out)
            manifest[str(rel)] = hashlib.sha256(src.read_bytes()).hexdigest()
    if (REF / "requirements.txt").exists():
        shutil.copyfile(REF / "requirements.txt", DST / "requirements.txt")
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(REF), "files": manifest}, indent=1))
    print(f"[build_ref] staged {len(manifest)} reference files under {DST}")
    return DST


if __name__ == "__main__":
    stage(force=True)
