"""CPU tier: the C-ABI library builds, loads and exports every symbol include/snaq_b200.h
declares; without a GPU the product fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import snaq_jl_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "snaq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(snaq_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = S.load_library()
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/snaq_b200.h but not exported"
    assert sorted(S.engine.EXPORTS) == syms
    assert b"sm_100a" in lib.snaq_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the failure path cannot be exercised")
    with pytest.raises(S.SnaqError, match="no CUDA device|CUDA"):
        S.Engine(0)


def test_product_does_not_reference_the_oracle():
    """The product tree must not import, include, link or execute anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "snaq.jl_b200")
    bad = re.compile(r"(^\s*(from|import)\s+(oracle|tests)\b)|(#include\s+\"[^\"]*(oracle|hostsim|tests)[^\"]*\")|"
                     r"(libsnaq_oracle|libhostsim|oracle/|tests/hostsim/)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                m = bad.search(txt)
                assert m is None, f"{f} references test infrastructure: {m.group(0)!r}"
    # and the shipped library links neither
    import subprocess
    out = subprocess.run(["ldd", S.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hostsim" not in out


def test_only_sm100a_code_in_the_library():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", S.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
