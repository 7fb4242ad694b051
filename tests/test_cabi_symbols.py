"""CPU-only: the C-ABI shared library loads and exports exactly what include/cosmomap2_b200.h
declares, and the ctypes table in cosmomap2_b200/_cabi.py covers every declaration."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cosmomap2_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|int64_t|double|const char \*)\s*(cm2_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        nargs = 0 if args in ("", "void") else len(args.split(","))
        decls[m.group(2)] = nargs
    return decls


def test_header_declares_functions():
    decls = declared_functions()
    assert len(decls) >= 30
    for name in ("cm2_pointing_apply", "cm2_pointing_apply_t", "cm2_amatvec_white", "cm2_bd_apply",
                 "cm2_weights_moments", "cm2_m2_apply", "cm2_pcg_update_xr"):
        assert name in decls


def test_library_exports_every_declared_symbol():
    from cosmomap2_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "library does not export %s" % name
    assert _cabi.lib.cm2_version() >= 100


def test_ctypes_table_matches_header():
    from cosmomap2_b200 import _cabi
    decls = declared_functions()
    assert set(decls) == set(_cabi.SIGNATURES), (set(decls) ^ set(_cabi.SIGNATURES))
    for name, nargs in decls.items():
        assert len(_cabi.SIGNATURES[name][1]) == nargs, name


def test_no_cpu_fallback():
    """Without a GPU every compute entry of the Python surface must refuse loudly."""
    import numpy as np
    import pytest
    import torch
    import cosmomap2_b200 as cm
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        cm.ProcessTimeSamples(np.arange(10), 10)
    with pytest.raises(RuntimeError):
        cm.SparseLO(10, 10, np.arange(10))
    with pytest.raises(RuntimeError):
        cm.cg(None, np.ones(3))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under cosmomap2_b200/ may reference it."""
    pkg = os.path.join(ROOT, "cosmomap2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "from .. import oracle" not in txt
