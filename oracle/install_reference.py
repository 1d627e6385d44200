#!/usr/bin/env python3
"""Recipe: put the UNMODIFIED reference next to the repo for the CPU arm of bench.py.

The reference (H2muller/CROPSR) is two Python files with no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` answers "Directory '/root/reference' is not
installable" (recorded in DESIGN.md).  What an install would do -- put the module files where
bench.py can run them -- is done here by a byte-for-byte copy of

    /root/reference/CROPSR.py  /root/reference/cropsr_functions.py   ->   baseline/_ref/

`baseline/_ref/` is git-ignored (no reference source enters the history) but NOT gpurun-ignored:
it travels to the GPU box with the snapshot, where /root/reference does not exist.  The copy is
verified against the sha256 of the sources and described in baseline/_ref/MANIFEST.json.
TEST / BENCH INFRASTRUCTURE ONLY: nothing under cropsr_b200/ reads it.

usage: python oracle/install_reference.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("CROPSR.py", "cropsr_functions.py")


def sha256(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def install(src="/root/reference", dest=DEST):
    """Returns the manifest, or None when the reference sources are not on this machine."""
    if not all(os.path.exists(os.path.join(src, f)) for f in FILES):
        return None
    os.makedirs(dest, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dest, f))
        a, b = sha256(os.path.join(src, f)), sha256(os.path.join(dest, f))
        if a != b:
            raise RuntimeError(f"{f}: copy differs from the source")
        manifest["files"][f] = a
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return manifest


def installed(dest=DEST):
    """The manifest of an intact install, else None."""
    try:
        with open(os.path.join(dest, "MANIFEST.json")) as fh:
            manifest = json.load(fh)
        for f, digest in manifest["files"].items():
            if sha256(os.path.join(dest, f)) != digest:
                return None
        return manifest
    except Exception:
        return None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    m = install(ap.parse_args().src)
    if m is None:
        sys.exit("reference sources not found")
    print(json.dumps(m, indent=1))
