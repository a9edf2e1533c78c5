"""CPU-only checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol that
include/coach_b200.h declares; the ctypes table in coach_b200/_lib.py covers the same set.  No compute calls."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "coach_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from coach_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), "libcoach_b200.so does not export %s" % name
    assert lib.cb200_abi_version() == 1


def test_ctypes_table_matches_header():
    from coach_b200 import _lib
    assert sorted(_lib.PROTOTYPES.keys()) == header_symbols()


def test_only_sm100a_code_in_library():
    import subprocess
    from coach_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_errors_surface_as_valueerror():
    # argument validation happens before any CUDA call, so this is safe without a GPU
    from coach_b200 import _lib
    lib = _lib.load()
    with pytest.raises(ValueError):
        _lib.check(lib.cb200_per_init(None, None, None, None, 8, None))
    with pytest.raises(ValueError):
        _lib.check(lib.cb200_per_sample(1, 1, 12, 1, 4, 4, 0.4, 1, None, None, None))   # size not a power of two
    assert b"power of 2" in lib.cb200_last_error()
