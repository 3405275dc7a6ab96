"""clid_slam_b200 — B200-native implementation of CLID-SLAM's neural-SDF training/query hot path.

    from clid_slam_b200 import Decoder, NeuralPoints, Mapper
    clid_slam_b200.install()        # make `from model.decoder import Decoder` etc. resolve here

Host code mirrors the reference's Python API (model.decoder.Decoder,
model.neural_points.NeuralPoints, utils.mapper.Mapper, utils.loss, part of utils.tools and
utils.data_sampler); the arithmetic runs in libclid_sdf.so (hand-written sm_100a kernels behind the C
ABI in include/clid_sdf.h).  See DESIGN.md and INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_OVERLAY = {
    "model.decoder": "clid_slam_b200.model.decoder",
    "model.neural_points": "clid_slam_b200.model.neural_points",
    "utils.mapper": "clid_slam_b200.utils.mapper",
    "utils.loss": "clid_slam_b200.utils.loss",
    "utils.data_sampler": "clid_slam_b200.utils.data_sampler",
    "model.local_point_cloud_map": "clid_slam_b200.model.local_point_cloud_map",
}
_FEEDERS = ("utils.data_sampler", "model.local_point_cloud_map")


def install(replace_sampler: bool = True) -> None:
    """Alias the reference's module names to this package in ``sys.modules`` so that an unmodified
    ``slam.py`` (``from model.neural_points import NeuralPoints`` ...) picks up the B200 path.

    Call it before the reference's modules are imported, with the reference tree on ``sys.path``
    (its ``utils.config``, ``utils.tools``, dataset / tracker / mesher code keeps being used as is).
    ``replace_sampler=False`` leaves the per-frame feeders (``utils.data_sampler``,
    ``model.local_point_cloud_map``) on the reference's own code; both implementations produce the
    same samples for the same torch seed (tests/test_host_logic.py)."""
    for ref_name, ours in _OVERLAY.items():
        if ref_name in _FEEDERS and not replace_sampler:
            continue
        module = importlib.import_module(ours)
        sys.modules[ref_name] = module
        parent, _, child = ref_name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, module)


def __getattr__(name):  # lazy re-exports keep `import clid_slam_b200` cheap
    if name == "Decoder":
        from .model.decoder import Decoder

        return Decoder
    if name == "NeuralPoints":
        from .model.neural_points import NeuralPoints

        return NeuralPoints
    if name == "Mapper":
        from .utils.mapper import Mapper

        return Mapper
    if name == "Config":
        from .config import Config

        return Config
    raise AttributeError(name)
