"""The C-ABI library loads and exports every symbol include/mosh2.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from moshpp_b200 import build, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def library():
    build.build_library()
    return lib.load_library()


def test_header_symbols_are_exported(library):
    header = open(os.path.join(ROOT, 'include', 'mosh2.h')).read()
    declared = set(re.findall(r'\b(mosh2_[a-z0-9_]+)\s*\(', header))
    declared -= {'mosh2_stageii_func'}
    assert declared == set(lib.EXPORTED_SYMBOLS), declared ^ set(lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(library, name), name
    version = int(re.search(r'#define MOSH2_VERSION (\d+)', header).group(1))
    assert library.mosh2_version() == version == lib.ABI_VERSION


def test_ctypes_struct_layout_matches_header(library):
    o = lib.Options()
    library.mosh2_default_options(C.byref(o))
    assert (o.wt_data, o.wt_poseB, o.wt_velo, o.wt_extrap_dmpl, o.num_train_markers) == (400.0, 1.6, 2.5, 6.0, 46.0)
    assert (o.delta_0, o.e3_first, o.e3, o.maxiter) == (0.5, 1e-3, 1e-2, 100)
    assert C.sizeof(lib.ModelDesc) % 8 == 0 and C.sizeof(lib.Result) == 8 * 8


def test_no_cpu_fallback(library, cases):
    """Without a GPU the product must fail loudly, not fall back to a CPU path."""
    if library.mosh2_device_count() > 0:
        pytest.skip('a GPU is present')
    case = cases('C4')
    with pytest.raises(lib.Mosh2Error, match='no CUDA device'):
        lib.Model(case['pack'])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'moshpp_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), fn


def test_stale_library_is_rejected(library, tmp_path, monkeypatch):
    """A library of another ABI version must not load silently (the ctypes structs would no longer match)."""
    monkeypatch.setattr(lib, 'ABI_VERSION', lib.ABI_VERSION + 1)
    with pytest.raises(lib.Mosh2Error, match='implements ABI'):
        lib.load_library(lib.default_library_path())
