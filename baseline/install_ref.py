#!/usr/bin/env python
"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box).

The reference is plain Python with no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing
to build ("neither 'setup.py' nor 'pyproject.toml' found"); what an install would give -- its modules, importable -- is a verbatim
copy of the source files of the render path.  bench.py's reference arm (`--impl reference`, `cpu_baseline`, `gpu_eager_baseline`)
imports them from there and runs the reference's own NeRFRenderer.render_rays / AutoRFMix code, untouched.  A sha256 manifest is
written beside the files so that the run records exactly which bytes were timed.

  python baseline/install_ref.py        (build container only: needs /root/reference)"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference/src"
FILES = ["renderer.py", "utils.py", "model_codenerf.py", "model_autorf.py", "model_supnerf.py"]


def install(verbose=True):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    json.dump({"source": SRC, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("staged", len(FILES), "reference modules under", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
