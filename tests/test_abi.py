"""CPU: the C-ABI library builds, loads, and exports every symbol include/sednet_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "sednet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sed_[a-z0-9_]+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as g
    lib_path = g.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.sed_version.restype = ctypes.c_int
    assert lib.sed_version() >= 100


def test_binding_table_matches_header():
    from sednet_b200.src import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.sed_error_string(-1).decode() == "invalid argument"
    assert lib.sed_sednet_workspace_bytes(1, 1024, 32) > 0


def test_sm100a_sass_present():
    import subprocess
    import __graft_entry__ as g
    out = subprocess.run(["cuobjdump", "-lelf", g.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sed-net_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
