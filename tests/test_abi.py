"""The C-ABI library builds for sm_100a, loads, exports every symbol include/cellula_b200.h
declares, and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

import cellula_b200
from cellula_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cellula_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    cellula_b200.build()
    assert os.path.exists(_lib.LIB_PATH)
    assert _lib.lib().cb200_version().decode().endswith("sm_100a")


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 25
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in cellula_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cellula_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_init_fails_loudly_without_gpu():
    with pytest.raises(cellula_b200.Cb200Error, match="no CUDA device|no CPU path"):
        cellula_b200.Context(0)
