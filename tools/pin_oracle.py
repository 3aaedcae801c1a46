"""Pin the oracle restatement against the compiled reference on the reference's
own compress fixtures (dev container only: needs /root/reference + oracle/_ref).

usage: python tools/pin_oracle.py [--levels 9,1] [--limit N]
"""
import argparse, bz2, glob, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import orclib

ap = argparse.ArgumentParser()
ap.add_argument("--levels", default="9")
ap.add_argument("--limit", type=int, default=0)
ap.add_argument("--suite", default="fuzz-collect,manual-compress,fuzz-divbwt")
a = ap.parse_args()
files = []
for s in a.suite.split(","):
    files += sorted(glob.glob("/root/reference/tests/suite/%s/*.bz2" % s))
if a.limit:
    files = files[: a.limit]
bad = periodic = blocks = 0
t0 = time.time()
for lv in [int(x) for x in a.levels.split(",")]:
    for f in files:
        raw = bz2.decompress(open(f, "rb").read())
        got, infos = orclib.orc_stream(raw, lv)
        want = orclib.ref_cli(raw, lv)
        blocks += len(infos)
        ties = [i for i in infos if i.tie_count > 1]
        if got != want:
            if ties and len(got) == len(want):
                periodic += 1
                print("PERIODIC", os.path.basename(f)[:12], lv, len(raw), [(i.nblock, i.tie_count, i.bwt_idx) for i in ties])
            else:
                bad += 1
                print("MISMATCH", f, lv, len(raw), len(got), len(want))
        elif ties:
            print("periodic-but-equal", os.path.basename(f)[:12], lv, [(i.nblock, i.tie_count, i.bwt_idx) for i in ties])
print("files=%d blocks=%d mismatches=%d periodic-exceptions=%d  %.1fs" % (len(files), blocks, bad, periodic, time.time() - t0))
