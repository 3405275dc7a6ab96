"""Compile libclid_sdf.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m clid_slam_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# developer knob: CLID_VARIANT=name builds/loads libclid_sdf_name.so (A/B kernel experiments with
# CLID_NVCC_FLAGS); unset, the product library libclid_sdf.so
VARIANT = os.environ.get("CLID_VARIANT", "")
LIB_PATH = os.path.join(PKG_DIR, f"libclid_sdf_{VARIANT}.so" if VARIANT else "libclid_sdf.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]
OBJ_DIR = os.path.join(CSRC, "_build_" + VARIANT if VARIANT else "_build")


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libclid_sdf.so cannot be built")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))


def _extra_flags():
    """Developer knob: CLID_NVCC_FLAGS="-DCLID_WALK_BATCH=2 ..." is appended to every compile."""
    return os.environ.get("CLID_NVCC_FLAGS", "").split()


def _compile_one(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + _extra_flags() + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-c", src, "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a (one nvcc process per translation unit, in parallel)
    and link them into libclid_sdf.so.  Only stale objects are recompiled."""
    from concurrent.futures import ThreadPoolExecutor

    stamp = os.path.join(OBJ_DIR, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS + _extra_flags())
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    src_time = max(os.path.getmtime(s) for s in sources())
    if not os.path.isdir(OBJ_DIR):
        # a shipped tree (the GPU box): objects do not travel, the library does
        if not force and not _extra_flags() and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(hdr_time, src_time):
            return LIB_PATH
        os.makedirs(OBJ_DIR, exist_ok=True)
    flags_changed = not os.path.exists(stamp) or open(stamp).read() != flags_now
    todo, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (force or flags_changed or not os.path.exists(obj)
                or os.path.getmtime(obj) < max(hdr_time, os.path.getmtime(src))):
            todo.append(src)
    if not todo and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(o) for o in objs):
        return LIB_PATH
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda s: _compile_one(s, verbose), todo))
        with open(stamp, "w") as fh:
            fh.write(flags_now)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"] + objs + ["-o", LIB_PATH]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
