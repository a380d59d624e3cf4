"""The C-ABI library loads on a CPU-only box and exports every symbol include/oobleck_b200.h declares."""
import os
import re

from oobleck_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "oobleck_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oob_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(L.exported_symbols())


def test_library_exports_all_symbols():
    lib = L.load()  # raises if missing; AttributeError if a symbol is not exported
    assert lib.oob_version() >= 100
    for name in declared_symbols():
        assert hasattr(lib, name)


def test_binding_argument_counts_match_the_prototypes():
    """Every prototype in the header has as many parameters as the ctypes signature in lib.py."""
    src = open(os.path.join(ROOT, "include", "oobleck_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"\s+", " ", src)
    protos = dict(re.findall(r"\b(oob_[a-z0-9_]+) ?\(([^()]*)\) ?;", src))
    assert set(protos) == set(L._SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(L._SIGNATURES[name][1]), (name, n, len(L._SIGNATURES[name][1]))
