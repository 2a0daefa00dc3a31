#!/usr/bin/env python
"""Recipe: vendor the UNMODIFIED reference (pure Python) into oracle/_ref/ so that it can travel to the GPU box.

  python oracle/make_ref.py [--src /root/reference]

oracle/_ref/ is build output: listed in .gitignore (never committed), NOT in .gpurunignore (it ships with the
snapshot, like the built .so).  Only `*.py` / `*.yml` files are taken; nothing is edited.  A manifest with the
sha256 of every file is written next to them so tests can tell that the copy is pristine.
TEST INFRASTRUCTURE: only tests/, __graft_entry__ and bench.py's reference arms may import from oracle/_ref.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".yml", ".yaml")
SKIP_DIRS = {".git", "__pycache__", "data_preprocessing"}


def vendor(src="/root/reference", dst=DST):
    if not os.path.isdir(src):
        return None
    manifest = {}
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            out_dir = os.path.join(dst, rel) if rel != "." else dst
            os.makedirs(out_dir, exist_ok=True)
            sp, dp = os.path.join(root, f), os.path.join(out_dir, f)
            shutil.copyfile(sp, dp)
            with open(sp, "rb") as fh:
                manifest[os.path.normpath(os.path.join(rel, f))] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    out = vendor(a.src)
    print(out if out else "reference tree %s not present: nothing vendored" % a.src)
    sys.exit(0 if out else 1)
