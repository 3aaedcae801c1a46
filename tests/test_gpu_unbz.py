"""GPU parity of the batch DECOMPRESSOR through the C ABI
(lbz_decompress_stream, lbz_scan_blocks, lbz_decoder_read) against the decode oracle and the
committed goldens, whose expectations are the compiled reference CLI's behaviour
(tests/golden/make_decode_golden.py)."""
import bz2
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import orclib
import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "decode")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
NOEMIT = 0xFFFFFFFFFFFFFFFF


def load(c):
    return open(os.path.join(GOLD, c["file"]), "rb").read()


@pytest.fixture(scope="module")
def dec():
    import lbzip2_b200
    d = lbzip2_b200.Decoder(device=0, max_blocks=24, in_cap=8 << 20, out_cap=96 << 20)
    yield d
    d.close()


def brute_scan(z):
    bits = np.unpackbits(np.frombuffer(z + b"\0" * ((-len(z)) % 4), np.uint8))
    pat = np.unpackbits(np.frombuffer(bytes.fromhex("314159265359"), np.uint8))
    if len(bits) < 48:
        return []
    return [int(p) for p in np.flatnonzero(bits[: len(bits) - 47] == pat[0]) if (bits[p:p + 48] == pat).all()]


def test_scan_matches_brute_force(dec):
    for c in MANIFEST[:60]:
        z = load(c)
        assert dec.scan(z) == brute_scan(z), c["file"]
    z = b"BZh9" + bytes.fromhex("314159265359")
    assert dec.scan(z) == [32] and dec.scan(z[:-1]) == []
    # every bit alignment
    rng = np.random.default_rng(3)
    magic = int("314159265359", 16)
    for shift in range(0, 70, 3):
        body = bytearray(rng.integers(0, 256, 64, dtype=np.uint8))
        v = int.from_bytes(body, "big")
        pos = 40 + shift
        v &= ~(((1 << 48) - 1) << (512 - pos - 48))
        v |= magic << (512 - pos - 48)
        z = v.to_bytes(64, "big")
        assert dec.scan(z) == brute_scan(z)


def test_goldens_status_and_output(dec):
    for c in MANIFEST:
        z = load(c)
        st, out, info = dec.decompress(z, cap=max(48 << 20, c["out_len"] + 16))
        name = orclib.ERR_NAMES[st] if 0 <= st < 20 else str(st)
        assert name == c["status"], (c["file"], c["name"], name, c["status"])
        assert len(out) == c["out_len"], (c["file"], c["name"])
        assert hashlib.sha256(out).hexdigest() == c["out_sha256"], (c["file"], c["name"])
        assert info.num_blocks == c["num_blocks"], (c["file"], c["name"])
        if st == 0:
            assert info.num_streams == c["num_streams"] and info.garbage == c["garbage"], c["file"]


def test_goldens_one_block_per_wave():
    import lbzip2_b200
    d = lbzip2_b200.Decoder(device=0, max_blocks=1, in_cap=1 << 20)
    for c in MANIFEST:
        if c["num_blocks"] < 2 and c["status"] != "OK":
            continue
        z = load(c)
        st, out, info = d.decompress(z, cap=max(48 << 20, c["out_len"] + 16))
        assert (orclib.ERR_NAMES[st] if st < 20 else st) == c["status"], c["file"]
        assert hashlib.sha256(out).hexdigest() == c["out_sha256"], c["file"]
        assert info.waves >= info.num_blocks
    d.close()


def test_stages_match_oracle(dec):
    checked = 0
    for c in MANIFEST:
        if c["status"] != "OK" or c["num_blocks"] == 0:
            continue
        z = load(c)
        st, out, info = dec.decompress(z, cap=c["out_len"] + 16)
        assert st == 0
        for slot in range(dec.last_wave_blocks):
            b = dec.block(slot)
            if b.out_off == NOEMIT:
                continue
            bi, bwt = orclib.orc_retrieve(z, b.pos + 80)
            assert (bi.status, bi.block_size, bi.bwt_idx, bi.rand, bi.end_bit) == (
                b.status, b.block_size, b.bwt_idx, b.rand, b.end_bit)
            assert (dec.array(1, slot, b.block_size) == bwt).all()
            txt = orclib.orc_ibwt(bwt, bi.bwt_idx, bi.rand)
            assert (dec.array(2, slot, b.block_size) == txt).all()
            rst, raw, crc = orclib.orc_unrle(txt)
            assert rst == 0 and len(raw) == b.out_len and crc == b.crc
            checked += 1
    assert checked > 40


def test_output_capacity_and_bomb(dec):
    bomb = [c for c in MANIFEST if c["name"] == "ref:ch255.bz2"][0]
    z = load(bomb)
    st, out, info = dec.decompress(z, cap=1000)
    assert st == 100 and len(out) == 0
    st, out, info = dec.decompress(z, cap=bomb["out_len"])
    assert st == 0 and hashlib.sha256(out).hexdigest() == bomb["out_sha256"]


def test_false_candidates_are_rejected(dec):
    c = [c for c in MANIFEST if c["name"] == "planted magic"][0]
    st, out, info = dec.decompress(load(c), cap=c["out_len"] + 16)
    assert st == 0 and hashlib.sha256(out).hexdigest() == c["out_sha256"]
    assert info.candidates > info.num_blocks and info.false_candidates >= 1


def test_roundtrip_with_own_compressor_and_libbz2():
    """Level-9 blocks of every BASELINE shape: our compressor's output and libbz2's (bit-aligned
    blocks) decode to the input; compared with the oracle's decode of the same bytes."""
    import lbzip2_b200
    rng = np.random.default_rng(9)
    fib = bytes(synth.fib(1_200_000))
    inputs = {
        "text": synth.text(2_500_000, seed=21),
        "random": bytes(rng.integers(0, 256, 1_900_000, dtype=np.uint8)),
        "runs+fib": b"a" * 1_500_000 + fib,
        "empty": b"",
        "one": b"x",
    }
    eng = lbzip2_b200.Engine(device=0, level=9, max_chunks=8)
    d = lbzip2_b200.Decoder(device=0, max_blocks=8, in_cap=8 << 20)
    for name, raw in inputs.items():
        for z in (eng.compress_stream(raw), bz2.compress(raw, 9)):
            st, out, info = d.decompress(z, cap=len(raw) + 64)
            assert st == 0 and out == raw, name
            ost, oout, osi = orclib.orc_decompress(z, cap=len(raw) + 64)
            assert ost == 0 and info.num_blocks == osi.num_blocks and info.num_streams == osi.num_streams
    eng.close()
    d.close()


def test_full_size_roundtrip_properties():
    """100 MB-class check by properties: decompress(compress(x)) == x by sha256, block count and
    every block CRC as the compressor reported them."""
    import lbzip2_b200
    raw = synth.text(30_000_000, seed=77)
    eng = lbzip2_b200.Engine(device=0, level=9, max_chunks=40)
    z = eng.compress_stream(raw)
    eng.close()
    d = lbzip2_b200.Decoder(device=0, max_blocks=80, in_cap=16 << 20, out_cap=128 << 20)
    st, out, info = d.decompress(z, cap=len(raw) + 64)
    assert st == 0 and len(out) == len(raw)
    assert hashlib.sha256(out).digest() == hashlib.sha256(raw).digest()
    assert info.waves == 1 and info.num_streams == 1
    # a smaller decoder takes several waves and must give the same bytes
    d2 = lbzip2_b200.Decoder(device=0, max_blocks=7, in_cap=16 << 20, out_cap=48 << 20)
    st2, out2, info2 = d2.decompress(z, cap=len(raw) + 64)
    assert st2 == 0 and out2 == out and info2.waves > 5 and info2.num_blocks == info.num_blocks
    d.close()
    d2.close()


def test_fuzz_against_oracle(dec):
    """Damaged inputs (truncations, bit flips, overwritten spans): status, surviving output and
    block count equal the oracle's, which is pinned on the reference CLI for such inputs."""
    rng = np.random.default_rng(99)
    small = [load(c) for c in MANIFEST if 8 < os.path.getsize(os.path.join(GOLD, c["file"])) < 6000]
    seen = {}
    for i in range(500):
        z = bytearray(small[int(rng.integers(len(small)))])
        kind = i % 4
        if kind == 0:
            z = z[: int(rng.integers(4, len(z)))]
        elif kind == 3:
            a, b = sorted(int(x) for x in rng.integers(4, len(z), 2))
            z[a:b] = bytes(rng.integers(0, 256, b - a, dtype=np.uint8))
        else:
            for _ in range(1 + kind):
                bit = int(rng.integers(32, 8 * len(z)))
                z[bit >> 3] ^= 0x80 >> (bit & 7)
        z = bytes(z)
        ost, oout, osi = orclib.orc_decompress(z, cap=48 << 20)
        st, out, info = dec.decompress(z, cap=48 << 20)
        assert st == ost and out == oout and info.num_blocks == osi.num_blocks, (st, ost, z.hex()[:80])
        seen[st] = seen.get(st, 0) + 1
    assert len(seen) >= 10, seen


def test_default_capacity_grows_with_the_output(dec):
    """Decoder.decompress without `cap`: inputs that expand far beyond any fixed ratio (10 MB of zeros
    are a few hundred bytes of .bz2) decode wave by wave into a growing output."""
    import bz2
    data = b"\0" * 10_000_000 + b"tail"
    z = bz2.compress(data, 9)
    assert len(z) * 64 < len(data)
    st, out, info = dec.decompress(z)
    assert st == 0 and out == data


def _pieces(z, rng, lo, hi):
    pos = 0
    while pos < len(z):
        k = int(rng.integers(lo, hi + 1))
        yield z[pos:pos + k]
        pos += k


def test_streaming_session_equals_whole_file_session(dec):
    """lbz_decoder_open_stream / lbz_decoder_feed / lbz_decoder_next: every golden in pieces of random
    sizes through a window of 300 KB -- status, output and counters of the whole-file session (which the
    goldens pin to the reference CLI); then 60 MB of text (66 blocks, 17 MB compressed) and 20 MB of
    incompressible bytes at -9 through windows of 6 MB and 3 MB: the window has to slide many times."""
    import lbzip2_b200
    rng = np.random.default_rng(21)
    win = lbzip2_b200.Decoder(device=0, max_blocks=8, in_cap=300_000, out_cap=96 << 20)
    try:
        for c in MANIFEST:
            z = load(c)
            if len(z) > 300_000 and c["status"] != "OK":
                continue
            cap = max(48 << 20, c["out_len"] + 16)
            st, out, info = win.decompress_pieces(_pieces(z, rng, 1 if len(z) < 2000 else 500, 7 if len(z) < 2000 else 70000), cap)
            name = orclib.ERR_NAMES[st] if 0 <= st < 20 else str(st)
            assert name == c["status"], (c["file"], c["name"], name, c["status"])
            assert len(out) == c["out_len"] and hashlib.sha256(out).hexdigest() == c["out_sha256"], (c["file"], c["name"])
            assert info.num_blocks == c["num_blocks"], (c["file"], c["name"])
            if st == 0:
                assert info.num_streams == c["num_streams"] and info.garbage == c["garbage"], c["file"]
    finally:
        win.close()
    eng = lbzip2_b200.Engine(device=0, level=9, max_chunks=80)
    for data, window in ((synth.text(60_000_000, offset=17), 6 << 20), (synth.random_bytes(20_000_000, seed=17), 3 << 20)):
        z = eng.compress_stream(data)
        assert len(z) > 2 * window
        d = lbzip2_b200.Decoder(device=0, max_blocks=16, in_cap=window, out_cap=64 << 20)
        st, out, info = d.decompress_pieces(_pieces(z, rng, 100_000, 2_000_000), 64 << 20)
        d.close()
        assert st == 0 and out == data
        assert info.end_bit == 8 * len(z) and info.num_streams == 1
    eng.close()
