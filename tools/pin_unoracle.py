"""Pin the DECODE oracle (oracle/bz_unoracle.c) against the compiled reference CLI
(dev container only: needs /root/reference + oracle/_ref).

  * every .bz2 fixture of the reference (tests/*.bz2, tests/suite/manual-expand/*.bz2):
    accept/reject, the error text, and the output bytes of accepted files
  * truncations and bit flips of a sample of them (same comparison)
  * round trips: reference-compressed fixtures decode to the original

usage: python tools/pin_unoracle.py [--mutations N]
"""
import argparse, bz2, glob, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import orclib

ap = argparse.ArgumentParser()
ap.add_argument("--mutations", type=int, default=400)
ap.add_argument("--roundtrips", type=int, default=150)
a = ap.parse_args()


def compare(name, z):
    rc, out, err = orclib.ref_cli_decompress(z)
    st, got, si = orclib.orc_decompress(z, cap=max(1 << 22, 300 * len(z)))
    if st == 100:
        st, got, si = orclib.orc_decompress(z, cap=1 << 30)
    ok = True
    if (rc == 0) != (st == 0):
        ok = False
    elif rc == 0:
        ok = (got == out)
    else:
        ok = orclib.ERR_TEXT.get(st, "?") in err
    if not ok:
        print("MISMATCH", name, "ref rc=%d err=%r len=%d | oracle %s len=%d" % (
            rc, err.strip()[-60:], len(out), orclib.ERR_NAMES[st] if st < 20 else st, len(got)))
    return ok, st


t0 = time.time()
files = sorted(glob.glob("/root/reference/tests/*.bz2")) + sorted(
    glob.glob("/root/reference/tests/suite/manual-expand/*.bz2"))
bad = 0
kinds = {}
for f in files:
    z = open(f, "rb").read()
    ok, st = compare(os.path.basename(f), z)
    kinds[st] = kinds.get(st, 0) + 1
    bad += not ok
print("fixtures=%d mismatches=%d statuses=%s" % (len(files), bad, {orclib.ERR_NAMES[k]: v for k, v in sorted(kinds.items())}))

rng = np.random.default_rng(11)
small = [f for f in files if os.path.getsize(f) < 200000]
mbad = 0
mk = {}
for i in range(a.mutations):
    f = small[int(rng.integers(len(small)))]
    z = bytearray(open(f, "rb").read())
    if len(z) < 8:
        continue
    if i % 2 == 0:
        z = z[: int(rng.integers(4, len(z)))]
        what = "trunc%d" % len(z)
    else:
        bit = int(rng.integers(32, 8 * len(z)))
        z[bit >> 3] ^= 0x80 >> (bit & 7)
        what = "flip%d" % bit
    ok, st = compare(os.path.basename(f)[:16] + ":" + what, bytes(z))
    mk[st] = mk.get(st, 0) + 1
    mbad += not ok
print("mutations=%d mismatches=%d statuses=%s" % (a.mutations, mbad, {orclib.ERR_NAMES[k]: v for k, v in sorted(mk.items())}))

rbad = 0
comp = []
for s in ("fuzz-collect", "manual-compress", "fuzz-divbwt"):
    comp += sorted(glob.glob("/root/reference/tests/suite/%s/*.bz2" % s))
for f in comp[:: max(1, len(comp) // a.roundtrips)]:
    raw = bz2.decompress(open(f, "rb").read())
    for lv in (1, 9):
        z = orclib.ref_cli(raw, lv)
        st, got, si = orclib.orc_decompress(z, cap=len(raw) + 1024)
        if st != 0 or got != raw:
            rbad += 1
            print("ROUNDTRIP", f, lv, st)
print("roundtrip mismatches=%d   %.1fs" % (rbad, time.time() - t0))
sys.exit(1 if bad or mbad or rbad else 0)
