"""The C-ABI library loads and exports every symbol include/dssmd_b200.h declares
(no compute calls here: this container has no GPU)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from dssmd_b200.build import build_extension
    return ctypes.CDLL(build_extension())


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dssmd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dssmd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for needed in ("dssmd_corr_fwd", "dssmd_corr_bwd", "dssmd_warp_fwd", "dssmd_warp_bwd",
                   "dssmd_lrwarp_error_fwd", "dssmd_last_error", "dssmd_version"):
        assert needed in syms


def test_every_declared_symbol_is_exported(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), f"{name} declared in include/dssmd_b200.h but not exported"


def test_version_and_error_string(built_lib):
    built_lib.dssmd_version.restype = ctypes.c_int
    built_lib.dssmd_last_error.restype = ctypes.c_char_p
    assert built_lib.dssmd_version() == 1
    assert isinstance(built_lib.dssmd_last_error(), bytes)


def test_binding_table_matches_header():
    """dssmd_b200/_lib.py binds exactly the compute entry points of the header,
    with as many arguments as the prototypes have."""
    from dssmd_b200 import _lib
    text = open(os.path.join(ROOT, "include", "dssmd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(dssmd_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    compute = {k: v for k, v in protos.items()
               if k not in ("dssmd_version", "dssmd_last_error", "dssmd_warp_bwd_workspace_bytes")}
    assert set(compute) == set(_lib.SIGNATURES)
    for name, args in compute.items():
        assert len(_lib.SIGNATURES[name]) == len(args.split(",")), name


def test_invalid_arguments_fail_without_touching_the_gpu(built_lib):
    """Argument validation happens before any CUDA call, so it is testable here."""
    built_lib.dssmd_last_error.restype = ctypes.c_char_p
    vp = ctypes.c_void_p
    rc = built_lib.dssmd_corr_fwd(vp(0), vp(0), vp(0), 1, 1, 1, 1, 1, 0, vp(0))
    assert rc == 1 and b"null" in built_lib.dssmd_last_error()
    rc = built_lib.dssmd_corr_fwd(vp(16), vp(16), vp(16), 1, 0, 1, 1, 1, 0, vp(0))
    assert rc == 1 and b"positive" in built_lib.dssmd_last_error()
    rc = built_lib.dssmd_corr_fwd(vp(16), vp(16), vp(16), 1, 1, 1, 1, 1, 99, vp(0))
    assert rc == 1 and b"dtype" in built_lib.dssmd_last_error()
    rc = built_lib.dssmd_warp_fwd(vp(16), vp(16), vp(16), vp(0), vp(0), 1, 3, 4, 4, 7, 0, vp(0))
    assert rc == 1 and b"padding_mode" in built_lib.dssmd_last_error()
    rc = built_lib.dssmd_warp_fwd(vp(16), vp(16), vp(16), vp(0), vp(0), 1, 3, 4, 4, 1, 1, vp(0))
    assert rc == 2  # bf16 warp: unsupported by design
