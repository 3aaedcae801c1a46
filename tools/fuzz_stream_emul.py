"""Long fuzz campaign of the decoder's STREAMING session (lbz_decoder_open_stream / feed / next, through
tests/simt_emul) against the whole-file session of the same source: damaged and spliced goldens, fed in
pieces of random sizes, greedy or starving, through windows of random sizes (down to barely more than the
largest block of the input).  Status, output and block count must agree; the end position when the file is accepted.
usage: python tools/fuzz_stream_emul.py [iterations] [seed] [max_file_bytes]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emulib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
maxb = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
GOLD = os.path.join(ROOT, "tests", "golden", "decode")
M = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
files = [open(os.path.join(GOLD, c["file"]), "rb").read() for c in M]
files = [f for f in files if 8 < len(f) <= maxb]
rng = np.random.default_rng(seed)
whole = emulib.EmuDecoder(max_blocks=8, in_cap=1 << 21)
wins = {}
seen, bad, t0 = {}, 0, time.time()


def pieces(z, lo, hi):
    pos = 0
    while pos < len(z):
        k = int(rng.integers(lo, hi + 1))
        yield z[pos:pos + k]
        pos += k


for i in range(iters):
    z = bytearray(files[int(rng.integers(len(files)))])
    kind = int(rng.integers(7))
    if kind == 0:
        z = z[: int(rng.integers(4, len(z)))]
    elif kind == 1:
        a, b = sorted(int(x) for x in rng.integers(4, len(z), 2))
        z[a:b] = bytes(rng.integers(0, 256, b - a, dtype=np.uint8))
    elif kind == 2:
        o = files[int(rng.integers(len(files)))]
        z = z[: int(rng.integers(4, len(z)))] + o[int(rng.integers(0, len(o))):]
    elif kind == 3:
        o = files[int(rng.integers(len(files)))]
        z = z + o                                        # concatenated streams (or garbage after the first)
    elif kind == 4:
        pass                                             # undamaged
    else:
        for _ in range(int(rng.integers(1, 5))):
            bit = int(rng.integers(32, 8 * len(z)))
            z[bit >> 3] ^= 0x80 >> (bit & 7)
    z = bytes(z)
    st0, out0, info0 = whole.decompress(z, cap=48 << 20)
    if st0 == 100:                                       # LBZ_ERR_OUTCAP of the one-call form: its capacity is per file, the session's per wave
        continue
    mb = int(rng.integers(1, 9))
    # windows: from generous down to ~the largest block of these goldens (< 200 KB compressed)
    cap = int(rng.choice([210_000, 260_000, 400_000, 1 << 20]))
    key = (mb, cap)
    if key not in wins:
        wins[key] = emulib.EmuDecoder(max_blocks=mb, in_cap=cap)
    lo, hi = [(1, 5), (1, 300), (50, 5000), (3000, 90000)][int(rng.integers(4))]
    try:
        st, out, info = wins[key].decompress_pieces(pieces(z, lo, hi), 48 << 20, greedy=bool(rng.integers(2)))
    except AssertionError as ex:
        st, out, info = -1, b"", None
        print("FAILED CALL", ex, z.hex()[:120], key, (lo, hi), flush=True)
    ok = st == st0 and out == out0 and info is not None and info.num_blocks == info0.num_blocks and \
        (st != 0 or (info.end_bit == info0.end_bit and info.num_streams == info0.num_streams and info.garbage == info0.garbage))
    seen[st0] = seen.get(st0, 0) + 1
    if not ok:
        bad += 1
        print("MISMATCH", i, st, st0, len(out), len(out0), key, (lo, hi), z.hex()[:160], flush=True)
        if bad > 5:
            break
    if (i + 1) % 500 == 0:
        print("%d iterations, %d mismatches, statuses %s, %.0f s" % (i + 1, bad, dict(sorted(seen.items())), time.time() - t0), flush=True)
print("done: %d iterations, %d mismatches" % (i + 1, bad))
