"""CPU tests of the drop-in boundary: the C-ABI library loads and exports exactly the symbols
include/nabu_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'nabu_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nabu_[a-z0-9_]+)\s*\(', text)))


def test_library_is_built_and_exports_every_header_symbol():
    from nabu_b200 import lib
    assert os.path.exists(lib.LIB_PATH), 'run python -m nabu_b200.build'
    cdll = ctypes.CDLL(lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(cdll, n), 'missing export %s' % n


def test_ctypes_signatures_cover_the_header():
    from nabu_b200 import lib
    assert sorted(lib.SIGNATURES) == _header_symbols()
    l = lib.load()
    assert l.nabu_version() >= 100
    assert l.nabu_kernel_launches() == 0


def test_no_cpu_fallback():
    import torch
    from nabu_b200 import lib
    with pytest.raises(lib.NabuError):
        lib.ptr(torch.zeros(4))            # CPU tensors are refused, not silently computed on
    assert lib.load().nabu_gemm_workspace_bytes() > 0


def test_product_does_not_import_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'nabu_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(import|from)\s+oracle\b', src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
