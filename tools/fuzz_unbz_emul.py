"""Long fuzz campaign of the decompressor's device logic (through tests/simt_emul) against the decode
oracle: truncations, bit flips, overwritten spans, spliced streams, on small and multi-block goldens.
usage: python tools/fuzz_unbz_emul.py [iterations] [seed] [max_file_bytes]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emulib, orclib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
maxb = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
GOLD = os.path.join(ROOT, "tests", "golden", "decode")
M = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
files = [open(os.path.join(GOLD, c["file"]), "rb").read() for c in M]
files = [f for f in files if 8 < len(f) <= maxb]
rng = np.random.default_rng(seed)
d = emulib.EmuDecoder(max_blocks=int(rng.integers(1, 9)), in_cap=1 << 20)
seen, bad, t0 = {}, 0, time.time()
for i in range(iters):
    z = bytearray(files[int(rng.integers(len(files)))])
    kind = int(rng.integers(6))
    if kind == 0:
        z = z[: int(rng.integers(4, len(z)))]
    elif kind == 1:
        a, b = sorted(int(x) for x in rng.integers(4, len(z), 2))
        z[a:b] = bytes(rng.integers(0, 256, b - a, dtype=np.uint8))
    elif kind == 2:                      # splice the tail of another file in at a bit-ish position
        o = files[int(rng.integers(len(files)))]
        z = z[: int(rng.integers(4, len(z)))] + o[int(rng.integers(0, len(o))):]
    elif kind == 3:                      # duplicate a span
        a, b = sorted(int(x) for x in rng.integers(4, len(z), 2))
        z = z[:b] + z[a:b] + z[b:]
    else:
        for _ in range(int(rng.integers(1, 5))):
            bit = int(rng.integers(32, 8 * len(z)))
            z[bit >> 3] ^= 0x80 >> (bit & 7)
    z = bytes(z)
    ost, oout, osi = orclib.orc_decompress(z, cap=48 << 20)
    st, out, info = d.decompress(z, cap=48 << 20)
    if st != ost or out != oout or info.num_blocks != osi.num_blocks:
        bad += 1
        name = "/tmp/fuzz_fail_%d_%d.bz2" % (seed, i)
        open(name, "wb").write(z)
        print("MISMATCH", name, "emu", st, len(out), info.num_blocks, "oracle", ost, len(oout), osi.num_blocks, flush=True)
    seen[st] = seen.get(st, 0) + 1
print("iterations %d mismatches %d statuses %s  %.0fs" % (iters, bad, {orclib.ERR_NAMES[k] if k < 20 else k: v for k, v in sorted(seen.items())}, time.time() - t0))
