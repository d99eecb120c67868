"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/sph.h declares, and the product fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from salva_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_bound_and_exported():
    names = _declared_symbols()
    assert len(names) >= 20
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libsalva_b200.so does not export %s" % n
        assert n in _lib.SYMBOLS, "python binding misses %s" % n
    assert set(_lib.SYMBOLS) == set(names)


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.WorldDesc) == 15 * 4
    assert C.sizeof(_lib.ForceDesc) == 4 + 8 * 4
    assert C.sizeof(_lib.StepStats) == 12 * 4 + 4 * 4 + 2 * 4 + 3 * 8 + 4 + 3 * 4 + 8 + 4 * 4


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import, link or execute anything under oracle/."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s*[\"<].*oracle|liboracle|CDLL\(.*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "salva_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), "%s references the oracle" % os.path.join(dirpath, f)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from salva_b200 import LiquidWorld, SphError
    with pytest.raises(SphError) as e:
        LiquidWorld(particle_radius=0.05)
    assert e.value.status == 2  # SPH_ERR_CUDA
