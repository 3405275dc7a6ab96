"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY -- used by oracle/gen_golden.py to produce tests/golden/*.npz and
by tests that are skipped when /root/reference is absent (it is absent on the GPU box).
Viz / IO packages the reference imports at module level but never touches on the hot path
are replaced by MagicMock stubs.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("CLID_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "open3d", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "roma", "skimage",
    "skimage.measure", "pypose", "natsort", "plyfile", "laspy", "evo", "pyquaternion",
    "rerun", "dtyper", "wandb",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def load():
    """Returns a namespace with the reference's Config, Decoder, NeuralPoints, Mapper,
    LocalPointCloudMap, loss and tools modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = MagicMock(name=name)
    # the reference uses top-level package names `model` and `utils`
    for taken in ("model", "utils"):
        mod = sys.modules.get(taken)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REFERENCE_ROOT):
            raise RuntimeError(f"module name {taken!r} already imported from elsewhere")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from types import SimpleNamespace

    from model.decoder import Decoder
    from model.local_point_cloud_map import LocalPointCloudMap
    from model.neural_points import NeuralPoints
    from utils import loss as ref_loss
    from utils import tools as ref_tools
    from utils.config import Config
    from utils.mapper import Mapper

    return SimpleNamespace(
        Config=Config, Decoder=Decoder, NeuralPoints=NeuralPoints, Mapper=Mapper,
        LocalPointCloudMap=LocalPointCloudMap, loss=ref_loss, tools=ref_tools,
    )
