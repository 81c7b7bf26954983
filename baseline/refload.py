"""Loader of the staged reference modules (baseline/_ref/, written by tools/stage_reference.py; git-ignored).

Measurement / test infrastructure only - the package never imports this. Two ways to load the reference's flat modules:

  load_reference()            the UNMODIFIED modules (hourglass, loss, transform, evaluate, train, ...): the reference arm
                              of bench.py, the cuDNN "library bar" and the bf16-autocast comparator of the parity tests.
  load_reference_drivers()    the reference's DRIVERS (train.py, evaluate.py, optim.py, utils.py, config.py; the two
                              files with the one-token squeeze fix from baseline/_ref/patched/) with the three shim
                              modules of INTEGRATION.md section 1 ahead of them on sys.path, i.e. the reference's own
                              loops running over the B200 kernels.

Both return a dict name -> module and leave sys.modules / sys.path as they found them (the flat names `hourglass`,
`loss`, ... would otherwise collide between the two flavours).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SHIMS = os.path.join(os.path.dirname(HERE), "real_time_helmet_detection_b200", "shims")
FLAT = ("hourglass", "loss", "transform", "evaluate", "train", "optim", "utils", "config", "data", "main")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "MANIFEST.json"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    """imgaug / torchsummary are not installed in this image and only serve the dataloader / a summary print."""
    try:
        import imgaug  # noqa: F401
    except Exception:
        ia = _stub("imgaug")
        ia.augmenters = _stub("imgaug.augmenters")
        ia.augmentables = _stub("imgaug.augmentables")
        ia.augmentables.bbs = _stub("imgaug.augmentables.bbs", BoundingBox=object, BoundingBoxesOnImage=object)
    try:
        import torchsummary  # noqa: F401
    except Exception:
        _stub("torchsummary", summary=lambda *a, **k: None)


def _load(paths, names):
    if not available():
        raise FileNotFoundError("baseline/_ref is not staged: run `python tools/stage_reference.py` in the build container")
    install_stubs()
    saved = {n: sys.modules.pop(n) for n in FLAT if n in sys.modules}
    old_path = list(sys.path)
    sys.path[:0] = paths
    try:
        for n in names:
            importlib.import_module(n)
    finally:
        sys.path[:] = old_path
        loaded = {n: sys.modules.pop(n) for n in FLAT if n in sys.modules}
        sys.modules.update(saved)
    return loaded


def load_reference(names=("hourglass", "loss", "transform", "evaluate", "train", "optim")):
    return _load([REF], names)


def load_reference_drivers(names=("train", "evaluate", "optim", "config")):
    return _load([SHIMS, os.path.join(REF, "patched"), REF], names)
