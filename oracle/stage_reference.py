#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY.  Stages the UNMODIFIED reference (pure Python: it has no setup.py / pyproject, so
`pip install --target` cannot install it) from /root/reference into oracle/_ref/EvDeblurNeRF so that it travels to the GPU
box with the snapshot (oracle/_ref/ is git-ignored, NOT gpurun-ignored) and `bench.py --impl reference` / the eager-GPU
baseline leg can time the reference's OWN code (`kind: "reference"`) instead of the oracle port.

    python oracle/stage_reference.py            # no-op with a message when /root/reference is absent (GPU box)

Only the Python sources on the render / loss path are copied, byte for byte (a MANIFEST with sha256 sums is written beside
them); nothing under oracle/_ref is ever committed, imported by the product, or edited."""
import hashlib
import json
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "EvDeblurNeRF")
TREES = ("networks", "utils", "data")
FILES = ("options.py", "run_nerf.py")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"stage_reference: {SRC} not present (GPU box); using what is already under {DST}" if os.path.isdir(DST)
                  else f"stage_reference: {SRC} not present and nothing staged")
        return os.path.isdir(DST)
    manifest = {}
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for tree in TREES:
        for root, _, files in os.walk(os.path.join(SRC, tree)):
            for f in files:
                if f.endswith(".py"):
                    s = os.path.join(root, f)
                    rel = os.path.relpath(s, SRC)
                    d = os.path.join(DST, rel)
                    os.makedirs(os.path.dirname(d), exist_ok=True)
                    shutil.copyfile(s, d)
                    manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(SRC, f), "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"stage_reference: {len(manifest)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
