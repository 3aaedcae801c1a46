"""CPU: collect() decides the block split on the host (engine.cu rle_feed) -- no device call.
Checked against the oracle's one-shot RLE1 stage (pinned on the reference, tests/test_oracle_vs_ref.py)
with the input offered in random pieces, which is how the -u mode of the reference scheduler feeds
one block (src/compress.c:160-187): the bytes taken and the full flag must not depend on the cuts."""
import ctypes as C

import numpy as np

import lbzip2_b200
import orclib
import synth


def feed_in_pieces(L, data, mbs, rng):
    st = C.create_string_buffer(L.encoder_alloc_size(mbs))
    L.encoder_init(st, mbs, 8)
    pos, full = 0, 0
    while pos < len(data) and not full:
        n = int(rng.integers(1, max(2, min(len(data) - pos, 3 * mbs)) + 1))
        piece = data[pos:pos + n]
        buf = C.create_string_buffer(piece, len(piece))
        left = C.c_size_t(len(piece))
        full = L.collect(st, buf, C.byref(left))
        took = len(piece) - left.value
        assert 0 <= took <= len(piece)
        if not full:
            assert left.value == 0, "a block that is not full takes everything it is offered"
        pos += took
    return pos, full


def cases():
    rng = np.random.default_rng(42)
    out = []
    for _ in range(300):
        n = int(rng.integers(1, 4000))
        kind = rng.integers(0, 4)
        if kind == 0:      # long runs, lengths around the 4 / 259 thresholds
            parts = []
            while sum(len(p) for p in parts) < n:
                parts.append(bytes([int(rng.integers(0, 3))]) * int(rng.choice([1, 2, 3, 4, 5, 6, 258, 259, 260, 263, 518, 519, 600, 1200])))
            data = b"".join(parts)[:n]
        elif kind == 1:
            data = bytes(rng.integers(0, 2, n, dtype=np.uint8))
        elif kind == 2:
            data = synth.text(n, offset=int(rng.integers(0, 50)))
        else:
            data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        cap = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 17, 100, 259, 260, 1000, 3000]))
        out.append((data, cap))
    return out


def test_split_matches_the_oracle_for_any_piece_boundaries():
    L = lbzip2_b200.load_library()
    rng = np.random.default_rng(7)
    for data, cap in cases():
        a = np.frombuffer(data, dtype=np.uint8)
        block = np.zeros(cap + 8, np.uint8)
        used = np.zeros(256, np.uint8)
        nblock, consumed, crc = C.c_uint32(0), C.c_size_t(0), C.c_uint32(0)
        OL = orclib.oracle()
        u8p = C.POINTER(C.c_uint8)
        full = OL.orc_rle1(a.ctypes.data_as(u8p), a.size, cap, block.ctypes.data_as(u8p), C.byref(nblock), C.byref(consumed),
                           used.ctypes.data_as(u8p), C.byref(crc))
        for _ in range(3):
            got_pos, got_full = feed_in_pieces(L, data, cap, rng)
            assert got_pos == consumed.value, (cap, len(data), got_pos, consumed.value)
            # the oracle's one-shot call reports "full" only when the block closed inside the input
            if full:
                assert got_full == 1
            else:
                assert got_pos == len(data)
