"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports exactly what
include/rvs_b200.h declares (no compute call is made here: there is no GPU in this tier)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import orc


@pytest.fixture(scope="module")
def az():
    import __graft_entry__ as ge
    ge.build()
    import alphazero_reversi_b200 as m
    return m


def _header_functions():
    txt = open(os.path.join(orc.ROOT, "include", "rvs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rvs_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(az):
    names = _header_functions()
    assert len(names) >= 24
    L = az._lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in rvs_b200.h but not exported"
    assert sorted(az._lib.PROTOTYPES) == names  # the binding covers the whole header, nothing else
    out = subprocess.run(["nm", "-D", "--defined-only", az._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (rvs_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == names, "library exports symbols the header does not declare (or vice versa)"


def test_struct_layouts_match_header(az):
    assert C.sizeof(az._lib.EngineConfig) == 56
    assert C.sizeof(az._lib.EngineStats) == 104


def test_sm100a_only(az):
    out = subprocess.run(["cuobjdump", "-lelf", az._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_fails_loudly_without_gpu(az):
    import numpy as np
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(az.RvsError):
        az.ReversiGame().get_valid_moves()
    with pytest.raises(az.RvsError):
        az.Engine(4, 10, 1)
    with pytest.raises(az.RvsError):
        az.board_ops.perft(3)


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/"""
    pkg = os.path.join(orc.ROOT, "alphazero-reversi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in src and "import orc" not in src and "rvs_oracle.h" not in src, f
