"""CPU checks of the DECOMPRESSOR's device logic through the sequential SIMT emulation
(tests/simt_emul): the very source nvcc compiles for the GPU, run on the host, against the decode
oracle and the committed goldens (expectations = the compiled reference CLI, see
tests/golden/make_decode_golden.py).  The GPU twin of this file is tests/test_gpu_unbz.py."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import emulib
import orclib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "decode")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]


def load(c):
    return open(os.path.join(GOLD, c["file"]), "rb").read()


def test_mtf_front_matches_list_model():
    L = emulib.lib()
    rng = np.random.default_rng(1)
    lst = list(rng.permutation(256))
    words = (C.c_uint32 * 64)(*[sum(int(lst[4 * k + j]) << (8 * j) for j in range(4)) for k in range(64)])
    ranks = list(rng.integers(1, 256, 3000)) + list(rng.integers(1, 8, 3000)) + [255, 254, 4, 3, 1, 252, 251, 8, 7]
    for r in ranks:
        r = int(r)
        c = L.emu_mtf_front(words, r)
        want = lst.pop(r)
        lst.insert(0, want)
        assert c == want
        got = [(words[i >> 2] >> (8 * (i & 3))) & 255 for i in range(256)]
        assert got == [int(x) for x in lst]


def test_crc_shift_is_crc_of_concatenation():
    L = emulib.lib()
    O = orclib.oracle()

    def gf_mul(a, b):
        r = 0
        for i in range(31, -1, -1):
            r = ((r << 1) & 0xFFFFFFFF) ^ (0x04C11DB7 if r & 0x80000000 else 0)
            if (b >> i) & 1:
                r ^= a
        return r
    pw = [0x100]
    for _ in range(63):
        pw.append(gf_mul(pw[-1], pw[-1]))
    pwa = (C.c_uint32 * 64)(*pw)
    rng = np.random.default_rng(2)
    for la, lb in [(1, 1), (3, 1000), (77, 0), (0, 5), (4096, 123457)]:
        a = rng.integers(0, 256, la, dtype=np.uint8)
        b = rng.integers(0, 256, lb, dtype=np.uint8)
        za = np.zeros(1, np.uint8)
        ra = O.orc_crc_update(0, orclib._ptr(a if la else za, orclib.u8p), la)
        rb = O.orc_crc_update(0, orclib._ptr(b if lb else za, orclib.u8p), lb)
        ab = np.concatenate([a, b])
        whole = O.orc_crc_update(0xFFFFFFFF, orclib._ptr(ab, orclib.u8p), la + lb)
        got = L.emu_gf_shift(0xFFFFFFFF, la + lb, pwa) ^ L.emu_gf_shift(ra, lb, pwa) ^ rb
        assert got == whole


def brute_scan(z):
    bits = np.unpackbits(np.frombuffer(z + b"\0" * ((-len(z)) % 4), np.uint8))
    pat = np.unpackbits(np.frombuffer(bytes.fromhex("314159265359"), np.uint8))
    hits = []
    first = np.flatnonzero(bits[: len(bits) - 47] == pat[0]) if len(bits) >= 48 else []
    for p in first:
        if (bits[p:p + 48] == pat).all():
            hits.append(int(p))
    return hits


def test_scan_finds_magics_at_every_bit_offset():
    d = emulib.EmuDecoder(max_blocks=2, in_cap=1 << 20)
    rng = np.random.default_rng(3)
    magic = int("314159265359", 16)
    for shift in list(range(0, 33)) + [47, 63, 64, 65]:
        body = bytearray(rng.integers(0, 256, 64, dtype=np.uint8))
        v = int.from_bytes(body, "big")
        nb = 8 * len(body)
        pos = 40 + shift
        v &= ~(((1 << 48) - 1) << (nb - pos - 48))
        v |= magic << (nb - pos - 48)
        z = v.to_bytes(len(body), "big")
        assert d.scan(z) == brute_scan(z)
        assert pos in d.scan(z)
    for c in MANIFEST[:40]:
        z = load(c)
        assert d.scan(z) == brute_scan(z), c["file"]
    # magic at the very end, and cut by the end of the input
    z = b"BZh9" + bytes.fromhex("314159265359")
    assert d.scan(z) == brute_scan(z) == [32]
    assert d.scan(z[:-1]) == []
    d.close()


@pytest.mark.parametrize("max_blocks,reverse", [(8, 0), (1, 0), (3, 1)])
def test_goldens_status_and_output(max_blocks, reverse):
    emulib.lib().emu_set_reverse(reverse)
    d = emulib.EmuDecoder(max_blocks=max_blocks, in_cap=1 << 20)
    try:
        for c in MANIFEST:
            z = load(c)
            st, out, info = d.decompress(z, cap=max(48 << 20, c["out_len"] + 16))
            name = orclib.ERR_NAMES[st] if 0 <= st < 20 else str(st)
            assert name == c["status"], (c["file"], c["name"], name, c["status"])
            assert len(out) == c["out_len"], (c["file"], c["name"])
            assert hashlib.sha256(out).hexdigest() == c["out_sha256"], (c["file"], c["name"])
            assert info.num_blocks == c["num_blocks"], (c["file"], c["name"])
            if st == 0:
                assert info.num_streams == c["num_streams"] and info.garbage == c["garbage"], c["file"]
    finally:
        emulib.lib().emu_set_reverse(0)
        d.close()


def test_stages_match_oracle():
    """Last column, inverse BWT text and CRC of every block of the accepted goldens."""
    d = emulib.EmuDecoder(max_blocks=16, in_cap=1 << 20)
    checked = 0
    for c in MANIFEST:
        if c["status"] != "OK" or c["num_blocks"] == 0:
            continue
        z = load(c)
        st, out, info = d.decompress(z, cap=c["out_len"] + 16)
        assert st == 0
        for slot in range(d.last_wave_blocks):
            b = d.block(slot)
            if b.out_off == 0xFFFFFFFFFFFFFFFF:
                continue
            bi, bwt = orclib.orc_retrieve(z, b.pos + 80)
            assert (bi.status, bi.block_size, bi.bwt_idx, bi.rand, bi.end_bit) == (
                b.status, b.block_size, b.bwt_idx, b.rand, b.end_bit)
            assert (d.array(1, slot, b.block_size) == bwt).all()
            txt = orclib.orc_ibwt(bwt, bi.bwt_idx, bi.rand)
            assert (d.array(2, slot, b.block_size) == txt).all()
            rst, raw, crc = orclib.orc_unrle(txt)
            assert rst == 0 and len(raw) == b.out_len and crc == b.crc
            checked += 1
    assert checked > 40
    d.close()


def test_output_capacity_and_bomb():
    bomb = [c for c in MANIFEST if c["name"] == "ref:ch255.bz2"]
    assert bomb
    z = load(bomb[0])
    d = emulib.EmuDecoder(max_blocks=2, in_cap=1 << 16)
    st, out, info = d.decompress(z, cap=1000)
    assert st == 100 and len(out) == 0          # LBZ_ERR_OUTCAP, nothing claimed
    st, out, info = d.decompress(z, cap=bomb[0]["out_len"])
    assert st == 0 and hashlib.sha256(out).hexdigest() == bomb[0]["out_sha256"]
    d.close()


def test_false_candidates_are_rejected():
    c = [c for c in MANIFEST if c["name"] == "planted magic"][0]
    d = emulib.EmuDecoder(max_blocks=8, in_cap=1 << 20)
    st, out, info = d.decompress(load(c), cap=c["out_len"] + 16)
    assert st == 0 and hashlib.sha256(out).hexdigest() == c["out_sha256"]
    assert info.candidates > info.num_blocks and info.false_candidates >= 1
    d.close()


def _mutations(count, seed):
    rng = np.random.default_rng(seed)
    small = [load(c) for c in MANIFEST if 8 < os.path.getsize(os.path.join(GOLD, c["file"])) < 6000]
    for i in range(count):
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
        yield bytes(z)


def test_fuzz_against_oracle():
    """Damaged inputs: same status, same surviving output and same block count as the oracle
    (which is pinned on the reference CLI for exactly this kind of input, tools/pin_unoracle.py)."""
    d = emulib.EmuDecoder(max_blocks=6, in_cap=1 << 16)
    seen = {}
    for z in _mutations(int(os.environ.get("LBZ_FUZZ", "700")), 99):
        ost, oout, osi = orclib.orc_decompress(z, cap=48 << 20)
        st, out, info = d.decompress(z, cap=48 << 20)
        assert st == ost, (orclib.ERR_NAMES[st] if st < 20 else st, orclib.ERR_NAMES[ost] if ost < 20 else ost, z.hex()[:80])
        assert out == oout and info.num_blocks == osi.num_blocks
        seen[st] = seen.get(st, 0) + 1
    assert len(seen) >= 10, seen
    d.close()


# ---- streaming session: lbz_decoder_open_stream / lbz_decoder_feed / lbz_decoder_next -------------
def _pieces(z, rng, lo, hi):
    pos = 0
    while pos < len(z):
        k = int(rng.integers(lo, hi + 1))
        yield z[pos:pos + k]
        pos += k


@pytest.mark.parametrize("max_blocks,piece", [(8, (1, 7)), (3, (200, 5000)), (2, (60000, 90000))])
def test_streaming_session_equals_whole_file_session(max_blocks, piece):
    """Every golden fed in pieces of random sizes (single bytes up to more than a block) through a
    window far smaller than the larger files: status, output, block/stream counts and the absolute
    end position must equal the whole-file session's (which the goldens pin to the reference CLI)."""
    rng = np.random.default_rng(7 + max_blocks)
    whole = emulib.EmuDecoder(max_blocks=max_blocks, in_cap=1 << 20)
    # the window: a few blocks' worth (the largest compressed block of the goldens is < 200 KB)
    win = emulib.EmuDecoder(max_blocks=max_blocks, in_cap=300_000)
    cases = MANIFEST if piece[0] > 100 else [c for c in MANIFEST if os.path.getsize(os.path.join(GOLD, c["file"])) < 3000]
    try:
        for c in cases:
            z = load(c)
            if len(z) > 300_000 and c["status"] != "OK":
                continue                      # truncated giants: covered by the whole-file tests
            cap = max(48 << 20, c["out_len"] + 16)
            st0, out0, info0 = whole.decompress(z, cap=cap)
            trace = []
            st, out, info = win.decompress_pieces(_pieces(z, rng, *piece), cap, trace=trace)
            assert st == st0, (c["file"], c["name"], st, st0, trace[-5:])
            assert out == out0, (c["file"], c["name"])
            assert (info.num_blocks, info.num_streams, info.garbage) == (info0.num_blocks, info0.num_streams, info0.garbage), c["file"]
            if st == 0:                   # where the walk stands after an error depends on how far ahead it could look
                assert info.end_bit == info0.end_bit, (c["file"], c["name"], info.end_bit, info0.end_bit)
    finally:
        whole.close()
        win.close()


def test_streaming_session_window_slides_over_a_large_file():
    """A 2.6 MB stream of 40 small blocks + a concatenated second stream through a 160 KB window (the incompressible block alone takes 100 KB of it):
    the consumed front must be dropped many times, output and counters stay those of the file."""
    import synth
    data = synth.text(1_200_000, offset=3) + synth.random_bytes(90_000, seed=3) + b"\0" * 500_000
    z = orclib.orc_stream(data, 1)[0] + orclib.orc_stream(data[:250_000], 2)[0]
    whole = emulib.EmuDecoder(max_blocks=4, in_cap=len(z) + 64)
    st0, out0, info0 = whole.decompress(z, cap=len(data) + 300_000)
    whole.close()
    assert st0 == 0 and out0 == data + data[:250_000]
    win = emulib.EmuDecoder(max_blocks=4, in_cap=160_000)
    rng = np.random.default_rng(11)
    trace = []
    st, out, info = win.decompress_pieces(_pieces(z, rng, 1000, 40000), 48 << 20, trace=trace, greedy=False)
    assert st == 0 and out == out0
    st, out, _ = win.decompress_pieces(_pieces(z, rng, 1000, 40000), 48 << 20)
    win.close()
    assert st == 0 and out == out0
    assert (info.num_blocks, info.num_streams, info.end_bit) == (info0.num_blocks, info0.num_streams, info0.end_bit)
    assert trace.count(101) >= 5          # it really had to wait for input again and again


def test_streaming_session_fuzz_against_whole_file():
    d0 = emulib.EmuDecoder(max_blocks=6, in_cap=1 << 16)
    d1 = emulib.EmuDecoder(max_blocks=6, in_cap=1 << 16)
    rng = np.random.default_rng(5)
    for z in _mutations(int(os.environ.get("LBZ_FUZZ", "700")) // 2, 123):
        st0, out0, info0 = d0.decompress(z, cap=48 << 20)
        st, out, info = d1.decompress_pieces(_pieces(z, rng, 1, 300), 48 << 20, greedy=bool(len(z) & 1))
        assert (st, out, info.num_blocks) == (st0, out0, info0.num_blocks), z.hex()[:80]
        if st == 0:                       # where the walk stands after an error depends on how far ahead it could look
            assert info.end_bit == info0.end_bit, z.hex()[:80]
    d0.close()
    d1.close()
