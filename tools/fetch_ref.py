#!/usr/bin/env python
"""Place the UNMODIFIED reference next to the repo for the CPU arm of bench.py (`--impl reference`).

    python tools/fetch_ref.py [--src /root/reference]

The reference (ahmdtaha/distributed_sigmoid_loss) is five pure-Python files with no setup.py, so `pip install
--target baseline/_ref /root/reference` has nothing to build; this recipe copies the *.py files verbatim into
`baseline/_ref/` instead. That directory is git-ignored (never part of the history) but NOT gpurun-ignored, so it
travels to the GPU box with the tree exactly like the built `.so`. Nothing else in the repo reads it:
`bench.py --impl reference` imports `DDPSigmoidLoss` from there and runs it through its own public API
(distributed_sigmoid_loss.py:8-48) on the box's host cores. A manifest with the sha256 of every file is written
beside the copies so a reader can check they are byte-identical to upstream.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("distributed_sigmoid_loss.py", "distributed_utils.py", "rwightman_sigmoid_loss.py",
         "test_distributed_sigmoid_loss.py", "test_sigmoid_loss_variants.py")


def fetch(src: str = "/root/reference", quiet: bool = False) -> bool:
    """Returns True if baseline/_ref holds the reference afterwards (False: the source tree is not present here, e.g.
    on the GPU box, and no earlier copy exists)."""
    if not os.path.isdir(src):
        return os.path.exists(os.path.join(DEST, FILES[0]))
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for name in FILES:
        s = os.path.join(src, name)
        if not os.path.exists(s):
            continue
        d = os.path.join(DEST, name)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"reference: {len(manifest)} files -> {DEST}")
    return FILES[0] in manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if fetch(a.src) else 1)
