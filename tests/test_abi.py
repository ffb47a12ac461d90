"""The C-ABI library loads and exports every symbol include/pgbp_b200.h declares; the ctypes table of
the host mirror lists exactly those symbols.  No compute call is made (runs without a GPU)."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pgbp_b200  # noqa: E402
from pgbp_b200 import _lib  # noqa: E402

HEADER = os.path.join(ROOT, "include", "pgbp_b200.h")
PKG = os.path.join(ROOT, "phylogaussianbeliefprop.jl_b200")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int32_t|int64_t)\s+(pgbp_\w+)\s*\(", txt)))


def test_header_and_ctypes_table_agree():
    decl = declared_symbols()
    assert len(decl) >= 40
    assert sorted(_lib.SIGNATURES) == decl


@pytest.mark.parametrize("libname", ["libpgbp_b200.so", "libpgbp_emul.so"])
def test_library_exports_every_declared_symbol(libname):
    path = os.path.join(PKG, "lib", libname)
    if not os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("pgbp_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(emul=libname.endswith("emul.so"))
    dll = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(dll, name), name
    dll.pgbp_abi_version.restype = ctypes.c_int32
    assert dll.pgbp_abi_version() == 2
    lib = pgbp_b200.Library(path)  # binds restype / argtypes of every symbol
    buf = ctypes.create_string_buffer(64)
    assert lib.pgbp_last_error(buf, 64) == 0


def test_product_has_no_cpu_fallback(tmp_path):
    # a missing CUDA library is an ImportError, never a silent switch to another implementation
    with pytest.raises(ImportError):
        pgbp_b200.Library(str(tmp_path / "libpgbp_b200.so"))
    src = open(os.path.join(PKG, "api.py")).read() + open(os.path.join(PKG, "_lib.py")).read() + \
        open(os.path.join(PKG, "drivers.py")).read() + open(os.path.join(PKG, "sharding.py")).read()
    assert "oracle" not in src.replace("C oracle", "") and "libpgbp_emul" not in src
