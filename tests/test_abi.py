"""The C-ABI library loads and exports exactly the symbols include/gsx.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

from gradslam_b200 import _C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gsx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(gsx_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_list_the_same_symbols():
    assert _declared() == set(_C.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_C.LIB_PATH), "libgsx.so was not built"
    handle = ctypes.CDLL(_C.LIB_PATH)
    for name in sorted(_declared()):
        assert hasattr(handle, name), name


def test_host_only_entry_points():
    lib = _C.lib()
    assert lib.gsx_version() == 200
    n = lib.gsx_fusion_workspace_bytes(8, 480, 640)
    assert n >= 8 * 480 * 640 * 16
    assert 0 < lib.gsx_fusion_workspace_stats_offset(8, 480, 640) < n
    # argument validation happens before any launch, so it is testable without a GPU
    rc = lib.gsx_backproject_normals_fwd(None, 0, None, 0, None, 0, 1, 1, 4, 4, None, None, None, None, None)
    assert rc != 0 and b"null" in lib.gsx_last_error()


def test_no_oracle_import_in_product_package():
    """The product never routes through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "gradslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "gsx_oracle" not in src and "import oracle" not in src, os.path.join(dirpath, f)
