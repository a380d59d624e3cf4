"""The ctypes mirror in oobleck_b200/lib.py must lay every struct out exactly as include/oobleck_b200.h does: compile a
tiny C program against the header with the host compiler and compare sizeof / offsetof of every field."""
import ctypes as C
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oobleck_b200 import lib as L  # noqa: E402

PAIRS = [("oob_planes", L.Planes), ("oob_gemm_epilogue", L.GemmEpilogue), ("oob_dims", L.OobDims),
         ("oob_layer_params", L.OobLayerParams), ("oob_block_ctx", L.OobBlockCtx), ("oob_bwd_scratch", L.OobBwdScratch),
         ("oob_head_ctx", L.OobHeadCtx)]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs the host C compiler")
def test_ctypes_structs_match_the_header(tmp_path):
    header = open(os.path.join(ROOT, "include", "oobleck_b200.h")).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "oobleck_b200.h"', "int main(void) {"]
    for cname, cls in PAIRS:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S)
        assert body, f"{cname} not found in the header"
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            assert re.search(r"\b%s\b" % re.escape(fname), body.group(1)), f"{cname}.{fname} is not in the header"
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in PAIRS:
        assert int(out[cname]) == C.sizeof(cls), (cname, out[cname], C.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
