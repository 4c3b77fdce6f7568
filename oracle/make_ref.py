#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: a git-ignored, UNMODIFIED copy of the reference's hot-path packages so that the
reference's own CPU classes can be timed on the GPU box (``bench.py --impl reference``), where
``/root/reference`` does not exist.  TEST / BASELINE INFRASTRUCTURE ONLY -- nothing under ``gpu_se_b200/`` reads it.

    python oracle/make_ref.py            (build container; ``__graft_entry__.build()`` runs it when the reference is there)

Copies ``filter/``, ``gaussian_sum_dist/`` and ``model/`` byte for byte from ``/root/reference`` (no file is edited;
``MANIFEST.json`` records the sha256 of every file copied) and adds two EMPTY stub modules, ``cupy.py`` and ``osqp.py``:
the reference imports both at module top (filter/particle.py:6, gaussian_sum_dist/MultivariateGaussianSum.py:3,
controller.py:6) although its numpy code paths never call them, and neither is installed in this image.
``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never enter the history) but not in
``.gpurunignore`` (it travels to the GPU box like the built ``.so``).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("filter", "gaussian_sum_dist", "model")
STUB = '"""Empty stand-in: the reference imports this module but its numpy code paths never use it."""\nfloat32 = "stub"\n'


def make(reference_root="/root/reference", dest=DEST, quiet=False):
    if not os.path.isdir(os.path.join(reference_root, "filter")):
        raise RuntimeError("reference tree not found at %s" % reference_root)
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    manifest = {}
    for pkg in PACKAGES:
        src = os.path.join(reference_root, pkg)
        for dirpath, dirnames, files in os.walk(src):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(dirpath, f)
                rel = os.path.relpath(s, reference_root)
                d = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    for stub in ("cupy.py", "osqp.py"):
        with open(os.path.join(dest, stub), "w") as fh:
            fh.write(STUB)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference_root, "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print("oracle/_ref: %d reference files copied unmodified from %s" % (len(manifest), reference_root))
    return dest


if __name__ == "__main__":
    make(*(sys.argv[1:2]))
