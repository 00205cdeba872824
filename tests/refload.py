"""Locate and import the UNMODIFIED reference tree (test / bench infrastructure only — the product never imports it).

Search order: $GBCODEC_REF, /root/reference (build container), <repo>/baseline/_ref (verbatim copy made by
tools/install_reference.py; git-ignored, travels to the GPU box).  The reference's top-level packages are called
`models`, `utils`, `datasets`, `data`, `configs`: they are imported with the tree at the front of sys.path and
removed from sys.modules again by `release()`, so that they cannot shadow anything else in the test process.
`pycocotools` (absent from the image) is replaced by a stub: datasets/coco_dataset.py:14 imports it at module level
and the codec path never calls it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_TOP = ("models", "utils", "datasets", "data", "configs")


def find() -> Optional[str]:
    if os.environ.get("GBCODEC_NO_REFERENCE"):          # tests of the fall-back paths
        return None
    for cand in (os.environ.get("GBCODEC_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models")) and os.path.isfile(os.path.join(cand, "models", "fusion_head.py")):
            return cand
    return None


def _stub_pycocotools() -> None:
    if "pycocotools" in sys.modules:
        return
    try:
        importlib.import_module("pycocotools.coco")
        return
    except Exception:
        pass
    pk, coco, ev = types.ModuleType("pycocotools"), types.ModuleType("pycocotools.coco"), types.ModuleType("pycocotools.cocoeval")
    coco.COCO = object
    ev.COCOeval = object
    pk.coco, pk.cocoeval = coco, ev
    sys.modules["pycocotools"], sys.modules["pycocotools.coco"], sys.modules["pycocotools.cocoeval"] = pk, coco, ev


def release() -> None:
    ref = find()
    while ref and ref in sys.path:
        sys.path.remove(ref)
    for m in [m for m in sys.modules if m.split(".")[0] in _TOP]:
        mod = sys.modules[m]
        f = getattr(mod, "__file__", None) or ""
        if ref and f.startswith(ref):
            del sys.modules[m]


class Reference:
    """`with Reference() as ref:` -> ref.fusion_head, ref.pose_estimator, ref.models, ref.config, ref.coco_dataset (lazy)."""

    def __init__(self):
        self.path = find()

    @property
    def present(self) -> bool:
        return self.path is not None

    def __enter__(self):
        if not self.path:
            raise FileNotFoundError("reference tree not found ($GBCODEC_REF, /root/reference, baseline/_ref)")
        release()
        _stub_pycocotools()
        sys.path.insert(0, self.path)
        return self

    def __exit__(self, *exc):
        release()
        return False

    def module(self, name: str):
        return importlib.import_module(name)

    @property
    def fusion_head(self):
        return self.module("models.fusion_head")

    @property
    def pose_estimator(self):
        return self.module("models.pose_estimator")

    @property
    def models(self):
        return self.module("models")

    @property
    def config(self):
        return self.module("configs.config")

    @property
    def coco_dataset(self):
        return self.module("datasets.coco_dataset")

    @property
    def postprocess(self):
        return self.module("utils.postprocess")

    @property
    def losses(self):
        return self.module("models.losses")
