"""oracle/_ref/ = a VERBATIM copy of the reference's Python sources (git-ignored, NOT gpurun-ignored: it travels to the
GPU box like a built .so).  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # needs /root/reference (the build container)

Used by (a) `bench.py --impl reference`: the unmodified reference models on the box's host cores (`kind: "reference"`)
and, with `--ref-device cuda` / the `reference_gpu` record of the default line, through stock ATen on the same B200;
(b) tests/test_gpu_dropin.py: the unmodified model files, training loop and criterion running on top of the dropin/
shims.  Nothing under the product package imports it.  No reference SOURCE is committed: this script is the recipe."""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PCNBR_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
PARTS = ("models", "Training", "data_processing", "train.py")


def available() -> bool:
    return os.path.isdir(os.path.join(DST, "models", "utils"))


def build(force: bool = False) -> str | None:
    """Copy PARTS from the reference checkout; returns DST, or None when no reference is present (GPU box: the
    prebuilt copy is used)."""
    if not os.path.isdir(os.path.join(REF, "models")):
        return DST if available() else None
    if available() and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for part in PARTS:
        src = os.path.join(REF, part)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DST, part), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.md"))
        else:
            shutil.copy2(src, os.path.join(DST, part))
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
