"""The C-ABI library builds for sm_100a on a GPU-less box, loads, and exports every symbol include/coma_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from coma_b200 import build
    return build.build(force=False)


def _declared():
    hdr = open(os.path.join(ROOT, "include", "coma_b200.h")).read()
    return sorted(set(re.findall(r"COMA_API [^;(]*?\b(coma_\w+)\s*\(", hdr)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 13 and "coma_pair_accumulate_f32" in names and "coma_orient_accumulate_f32" in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/coma_b200.h but not exported"
    lib.coma_b200_version.restype = ctypes.c_int
    assert lib.coma_b200_version() >= 100
    lib.coma_b200_launch_count.restype = ctypes.c_int64
    assert lib.coma_b200_launch_count() == 0


def test_python_binding_covers_header(lib_path):
    from coma_b200 import _lib
    bound = set(_lib.SIGNATURES) | {"coma_b200_version", "coma_b200_last_error", "coma_b200_launch_count", "coma_b200_last_kernel"}
    assert bound == set(_declared())
    assert _lib.load() is not None


def test_bad_arguments_are_rejected_without_a_gpu(lib_path):
    from coma_b200 import _lib
    lib = _lib.load()
    rc = lib.coma_pair_accumulate_f32(None, None, 1, 1, 1, 0.1, 0.1, None, None, None)
    assert rc == -1 and b"null pointer" in lib.coma_b200_last_error()
    rc = lib.coma_orient_accumulate_f32(None, None, 1, 1, 1, None, 1, 0.2, 1e-10, None, None, None, None, None)
    assert rc == -1


def test_binary_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
