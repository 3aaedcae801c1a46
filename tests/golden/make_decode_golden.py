"""Generate the committed DECODE goldens (dev container only: needs /root/reference + oracle/_ref).

Inputs: every distinct .bz2 fixture of the reference (tests/*.bz2, tests/suite/manual-expand),
seeded truncations / bit flips of them, and streams made here by the reference CLI and by
libbz2 (bit-aligned blocks, several blocks, concatenations, trailing garbage, block
randomisation).  Expectation for each = what the compiled reference `lbzip2 -d -n1` does:
exit status, error text -> status name, sha256 and length of its output.

usage: python tests/golden/make_decode_golden.py
"""
import bz2, glob, hashlib, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib, synth

OUT = os.path.join(HERE, "decode")
os.makedirs(OUT, exist_ok=True)
for f in glob.glob(os.path.join(OUT, "*.bz2")):
    os.unlink(f)
cases = {}


def add(name, z):
    z = bytes(z)
    key = hashlib.sha1(z).hexdigest()[:16]
    if key in cases:
        return
    rc, out, err = orclib.ref_cli_decompress(z)
    status = "OK"
    if rc != 0:
        hit = [k for k, t in orclib.ERR_TEXT.items() if t in err]
        assert len(hit) == 1, (name, err)
        status = orclib.ERR_NAMES[hit[0]]
    st, got, si = orclib.orc_decompress(z, cap=max(1 << 22, 300 * len(z)))
    if st == 100:
        st, got, si = orclib.orc_decompress(z, cap=1 << 30)
    assert orclib.ERR_NAMES[st] == status, (name, status, st)
    if rc == 0:
        assert got == out, name
    fn = key + ".bz2"
    open(os.path.join(OUT, fn), "wb").write(z)
    cases[key] = dict(file=fn, name=name, status=status, out_len=len(out) if rc == 0 else len(got),
                      out_sha256=hashlib.sha256(out if rc == 0 else got).hexdigest(),
                      num_blocks=int(si.num_blocks), num_streams=int(si.num_streams), garbage=int(si.garbage))


files = sorted(glob.glob("/root/reference/tests/*.bz2")) + sorted(
    glob.glob("/root/reference/tests/suite/manual-expand/*.bz2"))
for f in files:
    add("ref:" + os.path.basename(f), open(f, "rb").read())

rng = np.random.default_rng(2024)
small = [open(f, "rb").read() for f in files if 8 < os.path.getsize(f) < 5000]
for i in range(260):
    z = bytearray(small[int(rng.integers(len(small)))])
    if i % 3 == 0:
        z = z[: int(rng.integers(4, len(z)))]
        add("trunc", z)
    else:
        bit = int(rng.integers(32, 8 * len(z)))
        z[bit >> 3] ^= 0x80 >> (bit & 7)
        add("flip", z)

# streams made here
text = synth.text(330_000, seed=5)
rnd = bytes(np.random.default_rng(3).integers(0, 256, 120_000, dtype=np.uint8))
runs = (b"a" * 7000 + b"abcd" * 300 + b"\x00" * 300 + b"zzzz" + b"y" * 259 + b"q" * 5) * 6
a1 = orclib.ref_cli(text, 1)               # lbzip2: 4 byte-aligned blocks
a2 = bz2.compress(text, 1)                 # libbz2: bit-aligned blocks
a3 = bz2.compress(rnd, 1)
a4 = orclib.ref_cli(runs, 9)
a5 = bz2.compress(b"abab" * 40, 9)         # periodic block
a6 = bz2.compress(b"x" * 50_000, 9)        # one run
add("lbzip2 text -1", a1)
add("libbz2 text -1", a2)
add("libbz2 random -1", a3)
add("lbzip2 runs -9", a4)
add("periodic", a5)
add("single run", a6)
add("concat", a2 + a4 + a5)
add("concat+garbage", a5 + a6 + b"\x00garbage")
add("garbage BZ", a5 + b"BZ")
add("garbage BZh0", a5 + b"BZh0xx")
add("empty+stream", orclib.ref_cli(b"", 9) + a5)
add("level mismatch", b"BZh1" + bz2.compress(text[:250_000], 3)[4:])   # block larger than the header allows
for cut in (len(a2) - 1, len(a2) - 4, len(a2) - 5, len(a2) - 11, len(a2) // 2, 15, 10):
    add("libbz2 cut %d" % cut, a2[:cut])
z = bytearray(a2); z[len(z) // 2] ^= 0x10
add("libbz2 flip", z)
# A block magic planted INSIDE a block: a false candidate for the scanner.  The selector list may
# be longer than the groups that follow need (src/decode.c:631), so 30 extra unary selector codes
# that spell 0x314159265359 (+ a closing 0 bit) are spliced into a libbz2 block and the count
# field is raised; CRCs are unaffected.
def plant_magic(z):
    bits = "".join("{:08b}".format(b) for b in z)
    p = 32 + 48 + 32 + 1 + 24
    big = bits[p:p + 16]; p += 16 + 16 * big.count("1")
    assert int(bits[p:p + 3], 2) >= 3
    p += 3
    nsel_at = p
    nsel = int(bits[p:p + 15], 2); p += 15
    for _ in range(nsel):
        while bits[p] == "1":
            p += 1
        p += 1
    pat = "{:048b}".format(0x314159265359) + "0"
    extra = pat.count("0")                      # every 0 closes one unary code
    out = bits[:nsel_at] + "{:015b}".format(nsel + extra) + bits[nsel_at + 15:p] + pat + bits[p:]
    # the old padding after the trailer is dropped and redone
    end = out.rindex("{:048b}".format(0x177245385090)) + 48 + 32
    out = out[:end] + "0" * ((-end) % 8)
    return int(out, 2).to_bytes(len(out) // 8, "big")


planted = plant_magic(bz2.compress(text[:6000], 9))
assert bz2.decompress(planted) == text[:6000]
add("planted magic", planted)
add("magic in trailing garbage", a5 + b"\x00" + bytes.fromhex("314159265359") + a5[10:40])

json.dump(dict(cases=sorted(cases.values(), key=lambda c: c["file"])), open(os.path.join(OUT, "manifest.json"), "w"), indent=0)
tot = sum(os.path.getsize(os.path.join(OUT, c["file"])) for c in cases.values())
from collections import Counter
print(len(cases), "cases,", tot, "bytes;", Counter(c["status"] for c in cases.values()))
