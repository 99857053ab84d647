"""Vendor the UNMODIFIED reference package into oracle/_ref/ so that the reference arm of bench.py can run the real
reference modules on the GPU box's host cores (SURVEY.md §8d "CPU baseline").

    python -m oracle.make_ref            (build container only: reads /root/reference)

Copies /root/reference/aimnet (pure Python + its YAML / D3 table data files) byte for byte into oracle/_ref/aimnet and
writes oracle/_ref/MANIFEST.json (sha256 of every copied file + the reference commit if known).  oracle/_ref/ is listed
in .gitignore (reference sources never enter the history) but not in .gpurunignore (it travels to the GPU box like the
built .so).  TEST / BASELINE INFRASTRUCTURE ONLY: nothing under aimnetcentral_b200/ imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("AIMNET_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
SKIP_DIRS = {"__pycache__"}


def vendored_available() -> bool:
    return os.path.isfile(os.path.join(DST, "aimnet", "__init__.py")) and os.path.isfile(os.path.join(DST, "MANIFEST.json"))


def make(force: bool = False) -> str | None:
    """Returns the vendored root, or None when the reference tree is absent (GPU box: the prebuilt copy is used)."""
    src_pkg = os.path.join(SRC, "aimnet")
    if not os.path.isdir(src_pkg):
        return DST if vendored_available() else None
    files = []
    for root, dirs, names in os.walk(src_pkg):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for n in sorted(names):
            if n.endswith((".pyc", ".pyo")):
                continue
            files.append(os.path.relpath(os.path.join(root, n), SRC))
    manifest = {}
    for rel in files:
        with open(os.path.join(SRC, rel), "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    commit = None
    try:
        commit = subprocess.run(["git", "-C", SRC, "rev-parse", "HEAD"], capture_output=True, text=True, timeout=10).stdout.strip() or None
    except Exception:  # noqa: BLE001
        pass
    man_path = os.path.join(DST, "MANIFEST.json")
    if not force and vendored_available():
        try:
            if json.load(open(man_path)).get("files") == manifest:
                return DST
        except Exception:  # noqa: BLE001
            pass
    if os.path.isdir(os.path.join(DST, "aimnet")):
        shutil.rmtree(os.path.join(DST, "aimnet"))
    for rel in files:
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), out)
    with open(man_path, "w") as f:
        json.dump({"source": SRC, "commit": commit, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    root = make(force="--force" in sys.argv)
    print(root if root else "reference tree not found and no vendored copy present")
