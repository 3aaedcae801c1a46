"""Small compress + decompress run for compute-sanitizer (memcheck / racecheck): smoke()'s stream plus a
streaming decoder session through a window smaller than the file.  usage: compute-sanitizer --tool memcheck
python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
import lbzip2_b200

data = synth.text(700_000, seed=7) + b"\x00" * 5000 + bytes(np.random.default_rng(1).integers(0, 256, 240_000, dtype=np.uint8))
eng = lbzip2_b200.Engine(device=0, level=1, max_chunks=12)
z = eng.compress_stream(data)
eng.close()
dec = lbzip2_b200.Decoder(device=0, max_blocks=4, in_cap=1 << 20)
st, back, info = dec.decompress(z, cap=len(data) + 64)
assert st == 0 and back == data
dec.close()
win = lbzip2_b200.Decoder(device=0, max_blocks=3, in_cap=260_000)
st, back, info = win.decompress_pieces((z[i:i + 37_000] for i in range(0, len(z), 37_000)), 48 << 20)
assert st == 0 and back == data and info.end_bit == 8 * len(z)
win.close()
print("sanitize smoke ok: %d -> %d bytes -> back, %d blocks; streaming session through a 260 KB window" % (len(data), len(z), info.num_blocks))
