"""Put the UNMODIFIED reference (DUTRobot/CLID-SLAM, /root/reference) under baseline/_ref so that it travels
to the GPU box with the repo snapshot (baseline/_ref is git-ignored, not gpurun-ignored).

The base recipe -- `pip install --no-index --target baseline/_ref /root/reference` -- does not apply: the
reference has no setup.py / pyproject.toml (it is run from its checkout, `python3 slam.py cfg.yaml`), so pip
refuses it ("does not appear to be a Python project").  The tree is pure Python; what the reference arm of
bench.py needs is its `model/`, `utils/` and `config/` directories, copied verbatim.  Nothing under
baseline/_ref is tracked by git and nothing in the product imports it.

    python -m baseline.install_ref            # no-op when /root/reference is absent (the GPU box)
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CLID_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
PARTS = ("model", "utils", "config")


def install(verbose: bool = False) -> bool:
    """True when baseline/_ref holds the reference afterwards."""
    if not os.path.isdir(os.path.join(REF_SRC, "model")):
        return os.path.isdir(os.path.join(REF_DST, "model"))
    for part in PARTS:
        src, dst = os.path.join(REF_SRC, part), os.path.join(REF_DST, part)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(REF_DST, "SOURCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {REF_SRC}/{{{','.join(PARTS)}}} (DUTRobot/CLID-SLAM); not part of this repository\n")
    if verbose:
        print("reference copied to", REF_DST)
    return True


if __name__ == "__main__":
    print("installed" if install(verbose=True) else "reference not available")
