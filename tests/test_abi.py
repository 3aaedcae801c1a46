"""CPU: the C-ABI library loads and exports every symbol include/lbzip2_b200.h
declares; without a GPU the product fails loudly instead of falling back."""
import os
import re

import pytest

import lbzip2_b200
from lbzip2_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lbzip2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", hdr)
    return sorted(set(names) - {"combine_crc"})


def test_library_exports_every_declared_symbol():
    L = lbzip2_b200.load_library()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), "missing export: " + name
    assert set(api.EXPORTS) <= set(decl) | {"lbz_version"}


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lbzip2_b200.LbzError):
        lbzip2_b200.Engine(device=0, level=1, max_chunks=1)
    with pytest.raises(lbzip2_b200.LbzError):
        lbzip2_b200.Decoder(device=0, max_blocks=1, in_cap=1 << 16)


def test_product_is_not_built_with_the_host_emulation():
    # tests/simt_emul compiles the decompressor's kernels for the host to check their logic where
    # there is no GPU; the product build must never define that switch nor export emulation symbols
    mk = open(os.path.join(ROOT, "lbzip2_b200", "csrc", "Makefile")).read()
    assert "UB_EMUL" not in mk and "simt_emul" not in mk
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", lbzip2_b200.lib_path()], capture_output=True, text=True).stdout
    assert "emu_" not in syms
    for root, _, files in os.walk(os.path.join(ROOT, "lbzip2_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "emul" not in open(os.path.join(root, f)).read(), f


def test_product_does_not_link_oracle():
    # the product sources never mention the oracle; only tests/bench may load it
    for root, _, files in os.walk(os.path.join(ROOT, "lbzip2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "liboracle" not in txt and "bz_oracle" not in txt and "orclib" not in txt, f
