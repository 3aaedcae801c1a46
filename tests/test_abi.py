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
    names += re.findall(r"\bextern\s+[a-z0-9_]+\s+([a-z_][a-z0-9_]*)\s*\[", hdr)      # data symbols (crc_table)
    return sorted(set(names) - {"combine_crc", "fn"})


def test_library_exports_every_declared_symbol():
    L = lbzip2_b200.load_library()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), "missing export: " + name
    assert set(api.EXPORTS) <= set(decl) | {"lbz_version"}


def test_crc_table_is_the_bzip2_table():
    """crc_table (reference src/decode.h:70, src/crctab.c): CRC-32/BZIP2, poly 0x04C11DB7, MSB first."""
    import ctypes as C
    import bz2
    L = lbzip2_b200.load_library()
    tab = (C.c_uint32 * 256).in_dll(L, "crc_table")
    assert tab[0] == 0 and tab[1] == 0x04C11DB7 and tab[255] == 0xB1F740B4
    # the first block's CRC follows "BZh9" and the 6-byte block magic: check the table against libbz2
    data = b"crc check through an independent implementation"
    crc = 0xFFFFFFFF
    for c in data:
        crc = ((crc << 8) & 0xFFFFFFFF) ^ tab[(crc >> 24) ^ c]
    z = bz2.compress(data)
    assert z[4:10] == bytes([0x31, 0x41, 0x59, 0x26, 0x53, 0x59])
    assert int.from_bytes(z[10:14], "big") == crc ^ 0xFFFFFFFF


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


def test_header_is_plain_c_and_cxx(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 and as C++17 on its own (no CUDA or torch
    types in any signature), and a C translation unit that uses every streaming-session entry point must
    link against the library."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "lbzip2_b200.h"\n'
                 "int use(lbz_decoder *d, const uint8_t *p, size_t n, uint8_t *o, size_t cap) {\n"
                 "  size_t taken = 0, got = 0; lbz_dstream_info inf;\n"
                 "  if (lbz_decoder_open_stream(d, 0) != LBZ_OK) return -1;\n"
                 "  if (lbz_decoder_feed(d, p, n, 1, &taken)) return -1;\n"
                 "  return lbz_decoder_next(d, o, cap, &got, &inf) == LBZ_NEED_INPUT;\n"
                 "}\n")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + inc, "-c", str(c), "-o", str(tmp_path / "t.o")])
    cxx = tmp_path / "t.cpp"
    cxx.write_text('#include "lbzip2_b200.h"\nint main() { return lbz_version() == nullptr; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + inc, str(cxx), "-L" + os.path.join(ROOT, "lbzip2_b200"),
                           "-lbz2b200", "-Wl,-rpath," + os.path.join(ROOT, "lbzip2_b200"), "-o", str(tmp_path / "t")])
