"""Copy the unmodified reference tree into baseline/_ref/ so that it travels to the GPU box.

    python tools/install_reference.py [--src /root/reference]

The reference is a plain Python source tree (no setup.py / pyproject.toml: `pip install /root/reference` has
nothing to build), so "installing" it is a copy of its importable packages and entry scripts.  baseline/_ref/ is
git-ignored (the reference's sources never enter this repo's history) but not gpurun-ignored.  The example images
and the analysis notebooks' assets are left out (1.5 MB of PNGs that nothing on the codec path imports).

Consumers: tests/refload.py (GPU tests of the patched REAL reference), bench.py --impl reference and the
`aten_cuda_baseline` leg of the default bench line.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SKIP_DIRS = {".git", "examples", "analysis", "__pycache__"}


def install(src: str = "/root/reference", dest: str = DEST) -> str:
    if not os.path.isdir(os.path.join(src, "models")):
        raise FileNotFoundError(f"{src} does not look like the reference tree")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    n = 0
    for cur, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(cur, src)
        os.makedirs(os.path.join(dest, rel), exist_ok=True)
        for f in files:
            if f.endswith((".py", ".yaml", ".yml", ".txt", ".md", ".sh")):
                shutil.copy2(os.path.join(cur, f), os.path.join(dest, rel, f))
                n += 1
    with open(os.path.join(dest, "INSTALLED_FROM"), "w") as fh:
        fh.write(f"{src}\n{n} files, verbatim copy by tools/install_reference.py\n")
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    print(install(a.src))
    sys.exit(0)
