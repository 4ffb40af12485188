"""The C-ABI library loads without a GPU and exports every symbol include/troute_b200.h declares.  CPU only:
no compute entry point is called (those need a device and fail loudly without one)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "troute_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from troute_b200 import _lib
    return _lib


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("trt_network_create", "trt_route", "trt_upload_forcing", "trt_run", "trt_download_results",
                 "trt_mc_segment_batch", "trt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    assert os.path.exists(built_lib.LIB_PATH)
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(handle, n)]
    assert not missing, missing


def test_ctypes_table_matches_header(built_lib):
    bound = sorted(name for name, _, _ in built_lib.SYMBOLS)
    assert bound == declared_symbols()


def test_version_and_error_string(built_lib):
    L = built_lib.lib()
    assert L.trt_version() >= 100
    assert isinstance(L.trt_last_error(), bytes)


def test_no_cpu_fallback_when_library_missing(monkeypatch, built_lib):
    """The product path must fail loudly if the CUDA extension is absent."""
    monkeypatch.setattr(built_lib, "_lib", None)
    monkeypatch.setattr(built_lib, "LIB_PATH", "/nonexistent/libtroute_b200.so")
    with pytest.raises(ImportError):
        built_lib.lib()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "t-route_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
