"""CPU checks of the C-ABI library: it builds, loads, exports every symbol that
include/c3cuda.h declares, and fails loudly (no CPU fallback) without a device."""

import ctypes
import os
import re

import pytest

from cogent3_b200 import _lib
from cogent3_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "c3cuda.h")).read()
    return sorted(set(re.findall(r"C3_API\s+[\w\s\*]+?\b(c3_\w+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = header_symbols()
    assert len(declared) >= 25
    assert sorted(_lib.SIGNATURES) == declared


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(raw, name), name


def test_abi_version(lib):
    assert lib.c3_abi_version() == 1


def test_no_cpu_fallback(lib):
    """without a CUDA device engine creation must raise, never compute on the CPU"""
    import numpy as np

    from cogent3_b200.engine import LikelihoodEngine
    from cogent3_b200.patterns import compress_alignment
    from cogent3_b200.tree import Topology

    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    topo, _ = Topology.from_newick("(a,b,c);")
    rows = np.eye(4)
    pats = compress_alignment(topo, np.array([[0, 1], [0, 2], [3, 1]]), rows)
    with pytest.raises(_lib.C3Error):
        LikelihoodEngine(pats)


def test_missing_library_raises(tmp_path):
    with pytest.raises(_lib.C3Error):
        _lib.load(str(tmp_path / "nope.so"))
