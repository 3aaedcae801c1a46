"""CPU: the DECODE oracle (oracle/bz_unoracle.c) against the committed decode goldens -- the
compiled reference CLI's verdict, error kind and output hash for 291 inputs
(tests/golden/make_decode_golden.py) -- and, where oracle/_ref is present, against the reference
CLI itself on fresh mutations."""
import hashlib
import json
import os

import numpy as np
import pytest

import orclib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "decode")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]


def test_oracle_matches_decode_goldens():
    kinds = set()
    for c in MANIFEST:
        z = open(os.path.join(GOLD, c["file"]), "rb").read()
        st, out, si = orclib.orc_decompress(z, cap=max(48 << 20, c["out_len"] + 16))
        assert orclib.ERR_NAMES[st] == c["status"], (c["file"], c["name"])
        assert len(out) == c["out_len"] and hashlib.sha256(out).hexdigest() == c["out_sha256"], c["file"]
        assert si.num_blocks == c["num_blocks"]
        kinds.add(c["status"])
    assert len(kinds) >= 12


def test_oracle_stage_functions_compose():
    """retrieve -> ibwt -> unrle of every block of an accepted multi-block golden reproduces the file."""
    c = [c for c in MANIFEST if c["name"] == "libbz2 text -1"][0]
    z = open(os.path.join(GOLD, c["file"]), "rb").read()
    pos, out = 32, b""
    for _ in range(c["num_blocks"]):
        assert z[pos // 8: pos // 8 + 7] is not None
        bi, bwt = orclib.orc_retrieve(z, pos + 80)
        assert bi.status == 0
        st, raw, crc = orclib.orc_unrle(orclib.orc_ibwt(bwt, bi.bwt_idx, bi.rand))
        assert st == 0
        stored = int.from_bytes(bytes(np.packbits(np.unpackbits(np.frombuffer(z, np.uint8))[pos + 48: pos + 80])), "big")
        assert crc == stored
        out += raw.tobytes()
        pos = bi.end_bit
    assert hashlib.sha256(out).hexdigest() == c["out_sha256"]


def test_oracle_matches_reference_cli_on_fresh_mutations():
    if not os.path.exists(os.path.join(orclib.REF_DIR, "lbzip2")):
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(2718)
    small = [open(os.path.join(GOLD, c["file"]), "rb").read() for c in MANIFEST
             if 8 < os.path.getsize(os.path.join(GOLD, c["file"])) < 5000]
    for i in range(150):
        z = bytearray(small[int(rng.integers(len(small)))])
        if i % 2:
            z = z[: int(rng.integers(4, len(z)))]
        else:
            bit = int(rng.integers(32, 8 * len(z)))
            z[bit >> 3] ^= 0x80 >> (bit & 7)
        z = bytes(z)
        rc, out, err = orclib.ref_cli_decompress(z)
        st, got, si = orclib.orc_decompress(z, cap=48 << 20)
        assert (rc == 0) == (st == 0), (rc, st, err)
        if rc == 0:
            assert got == out
        else:
            assert orclib.ERR_TEXT[st] in err, (st, err)
