"""The C-ABI shared library loads on a CPU-only machine and exports every symbol include/gnm.h declares
(no compute calls here - those need a GPU)."""
import ctypes
import os
import re

from graph_neural_mapping_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gnm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib.build()
    handle = ctypes.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(handle, n), "libgnm.so does not export %s" % n
    # the Python binding table covers the same set (plus nothing undeclared)
    bound = set(lib.SIGNATURES) | {"gnm_error_string"}
    assert bound == set(names), (sorted(bound - set(names)), sorted(set(names) - bound))


def test_abi_version_and_error_strings():
    l = lib.load()
    assert l.gnm_abi_version() == lib.ABI_VERSION
    assert l.gnm_error_string(0) == b"ok"
    assert b"bad argument" in l.gnm_error_string(-1)
    header = open(os.path.join(ROOT, "include", "gnm.h")).read()
    assert "#define GNM_ABI_VERSION %d" % lib.ABI_VERSION in header


def test_argument_checks_need_no_gpu():
    """Entry points validate arguments before touching the device: negative sizes are rejected with
    GNM_ERR_BAD_ARG, empty problems return GNM_OK without a launch."""
    l = lib.load()
    assert l.gnm_aggregate(None, None, -1, None, 0, None, None, 0, 4, 0, None, None, None) == -1
    assert l.gnm_aggregate(None, None, 0, None, 0, None, None, 0, 4, 0, None, None, None) == 0
    assert l.gnm_linear(None, 0, 0, 4, None, 0, 0, None, None, None, None, 0, 4, None, None, None) == 0
    assert l.gnm_linear_bwd(None, 0, None, 0, None, None, 0, None, None, None, None, None, 0, None, 0, None, None, 0,
                            None, 10, 100, 4, None, None) == -2
