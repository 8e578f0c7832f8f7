#!/usr/bin/env python
"""Stage the LIVE reference under ``oracle/_ref/`` (git-ignored, but shipped to the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY — nothing under ``mhaq_b200/`` imports it.

The reference (aifoundry-org/MHAQ) is pure Python: there is nothing to compile.  "Building"
it means copying its package tree, unmodified, from where it lies (``/root/reference``) to
``oracle/_ref/`` so that the same files travel to the GPU box, where ``/root/reference`` does
not exist.  Outputs go ONLY into ``oracle/_ref/`` (listed in .gitignore: the reference's
sources never enter this repository's history).  What is staged:

* ``src/**/*.py``      — the reference package (ops, layers, GDNSQQuant, ModelHelper,
                         PotentialLoss, model_stats, calibration, in-tree models, ...)
* ``config/*.yaml``    — the experiment configs BASELINE.json names
* ``MANIFEST.json``    — source path + sha256 of every staged file (so a test can assert the
                         staged copy is the unmodified reference)

Run by ``__graft_entry__.build()`` when ``/root/reference`` is present; a no-op otherwise
(the GPU box uses the files staged here).  ``oracle/ref_loader.py`` imports the staged tree.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = os.environ.get("MHAQ_REFERENCE", "/root/reference")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(src_root: str = SRC_DEFAULT, dest: str = DEST, verbose: bool = True) -> bool:
    """Copy the reference tree; returns True if a staged copy exists afterwards."""
    if not os.path.isdir(os.path.join(src_root, "src", "quantization")):
        if verbose:
            print(f"make_ref: {src_root} not present; keeping whatever is staged in {dest}")
        return os.path.isdir(os.path.join(dest, "src", "quantization"))
    manifest = {}
    for sub, exts in (("src", (".py",)), ("config", (".yaml", ".yml"))):
        for dirpath, dirnames, filenames in os.walk(os.path.join(src_root, sub)):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            for fn in filenames:
                if not fn.endswith(exts):
                    continue
                s = os.path.join(dirpath, fn)
                rel = os.path.relpath(s, src_root)
                d = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not os.path.exists(d) or _sha(d) != _sha(s):
                    shutil.copyfile(s, d)
                manifest[rel] = _sha(d)
    # drop staged files that no longer exist in the reference
    for dirpath, _, filenames in os.walk(dest):
        for fn in filenames:
            rel = os.path.relpath(os.path.join(dirpath, fn), dest)
            if rel != "MANIFEST.json" and rel not in manifest and not rel.endswith(".pyc"):
                os.remove(os.path.join(dirpath, fn))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src_root, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"make_ref: staged {len(manifest)} reference files under {dest}")
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    sys.exit(0 if ok else 1)
