"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/clid_sdf.h declares, validates arguments, and the product refuses CPU tensors."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from clid_slam_b200 import _lib
    from clid_slam_b200 import build as _build

    _build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from clid_slam_b200 import _lib

    header = open(os.path.join(ROOT, "include", "clid_sdf.h")).read()
    declared = set(re.findall(r"CLID_API\s+[\w\s\*]+?\b(clid_\w+)\s*\(", header))
    assert declared, "no CLID_API declarations found"
    assert declared == set(_lib.exported_symbols()), "ctypes signature table out of sync with the header"
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_error_string(lib):
    assert lib.clid_version() == 5
    rc = lib.clid_query_forward(None, None, None, None, 5, 0, None, None)
    assert rc == -1
    assert b"NULL" in lib.clid_last_error()


def test_zero_queries_is_a_noop(lib):
    from clid_slam_b200 import _lib

    m, out = _lib.ClidMap(), _lib.ClidQueryOut()
    assert lib.clid_query_forward(C.byref(m), None, None, None, 0, 0, C.byref(out), None) == 0
    assert lib.clid_query_forward(C.byref(m), None, None, None, -3, 0, C.byref(out), None) == -1


def test_struct_sizes_match_the_header(lib):
    """Compile a tiny C program against the header and compare sizeof() of EVERY struct with its ctypes mirror."""
    import subprocess
    import tempfile

    from clid_slam_b200 import _lib

    header = open(os.path.join(ROOT, "include", "clid_sdf.h")).read()
    names = re.findall(r"^\}\s*(Clid\w+);", header, flags=re.M)
    assert len(names) >= 15, names
    missing = [n for n in names if not hasattr(_lib, n)]
    assert not missing, f"structs without a ctypes mirror: {missing}"
    body = "\n".join(f'  printf("%zu\\n", sizeof({n}));' for n in names)
    src = "#include <stdio.h>\n#include \"clid_sdf.h\"\nint main(void) {\n" + body + "\n  return 0;\n}\n"
    with tempfile.TemporaryDirectory() as tmp:
        c_path, exe = os.path.join(tmp, "s.c"), os.path.join(tmp, "s")
        open(c_path, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c_path, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    for name, size in zip(names, sizes):
        assert size == C.sizeof(getattr(_lib, name)), f"{name}: C {size} bytes, ctypes {C.sizeof(getattr(_lib, name))}"


def test_product_refuses_cpu_tensors():
    from clid_slam_b200.config import Config
    from clid_slam_b200.model.neural_points import NeuralPoints

    cfg = Config()
    cfg.device = "cpu"
    cfg.buffer_size = 1009
    npm = NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(1)
    pts = torch.rand(500, 3) * 10
    npm.update(pts, torch.zeros(3), torch.eye(3), 0)  # host-side map maintenance works anywhere
    assert npm.count() > 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        npm.query_feature(pts[:4])


def test_install_aliases_reference_module_names():
    """clid_slam_b200.install() makes the names slam.py imports resolve to this package."""
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import clid_slam_b200; clid_slam_b200.install()\n"
        "from model.neural_points import NeuralPoints\n"
        "from model.decoder import Decoder\n"
        "from utils.mapper import Mapper\n"
        "from utils.loss import sdf_bce_loss\n"
        "assert NeuralPoints.__module__ == 'clid_slam_b200.model.neural_points'\n"
        "assert Mapper.__module__ == 'clid_slam_b200.utils.mapper' and Decoder.__module__.startswith('clid_slam_b200')\n"
        "from model.local_point_cloud_map import LocalPointCloudMap\n"
        "assert LocalPointCloudMap.__module__ == 'clid_slam_b200.model.local_point_cloud_map'\n"
        "assert sys.modules['utils.data_sampler'].__name__ == 'clid_slam_b200.utils.data_sampler'\n"
        "print('ok')\n" % ROOT
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_map_maintenance_entries_validate_their_arguments(lib):
    """SURVEY 8f entry points: sizes and NULL / range checks happen on the host, before any launch."""
    from clid_slam_b200 import _lib

    small, big = lib.clid_scan_workspace_bytes(1), lib.clid_scan_workspace_bytes(4_000_000)
    assert 16 <= small < big < 1 << 20, "the scan workspace grows with the element count (one entry per 2048 elements)"
    assert lib.clid_voxel_keys(None, None, 10, 0.4, None, None, None) == -1 and b"NULL" in lib.clid_last_error()
    assert lib.clid_map_insert_probe(None, None) == -1
    a = _lib.ClidInsertArgs()
    a.n = 5
    assert lib.clid_map_insert_probe(C.byref(a), None) == -1 and b"NULL" in lib.clid_last_error()
    w = _lib.ClidWindowArgs()
    assert lib.clid_local_window_select(C.byref(w), None) == -1 and b"m = 0" in lib.clid_last_error()
    r = _lib.ClidWindowRows()
    assert lib.clid_local_window_gather(C.byref(r), None) == -1
    assert lib.clid_table_store(None, None, 3, 0, None, 0, None) == -1
    assert lib.clid_table_store(None, None, 0, 0, None, 0, None) == -1, "a NULL table is refused even for zero rows"
    assert lib.clid_compact_rows(None, 3, None, None, None, 1, None) == -1
    rs = _lib.ClidRaySampleArgs()
    assert lib.clid_ray_samples(C.byref(rs), None) == 0, "zero scan points: nothing to do"
    rs.n_points = 4
    assert lib.clid_ray_samples(C.byref(rs), None) == -1 and b"NULL" in lib.clid_last_error()
    assert lib.clid_ray_labels(None, None, None, 4, 8, 4, None, None, None) == -1
    assert lib.clid_flag_ranks(None, 4, None, None, 0, None) == -1
    assert lib.clid_pool_filter_select(None, 4, None, 1.0, 0, 0, None, None, None, 0, None) == -1
