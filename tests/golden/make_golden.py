"""Generate the committed golden vectors (dev container only).

For every selected input the UNMODIFIED reference CLI (oracle/_ref/lbzip2, built
from /root/reference by oracle/Makefile) produces the expected .bz2; we store
  tests/golden/inputs/<name>.bz2   the raw input, recompressed with Python's bz2
                                   (storage only -- not a reference fixture copy)
  tests/golden/manifest.json       per (input, level): sha256 + length of the
                                   reference output, block count, and whether the
                                   input produces an exactly periodic block
Inputs: the decompressed payloads of the reference's own compress fixtures
(tests/suite/{fuzz-collect,manual-compress} in full, a spread of fuzz-divbwt)
plus synthetic inputs from tests/synth.py.

usage: python tests/golden/make_golden.py
"""
import bz2, glob, hashlib, json, os, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib, synth  # noqa: E402
import numpy as np  # noqa: E402

REF = "/root/reference/tests/suite"
cases = {}


def add(name, raw, levels):
    cases[name] = (raw, levels)


for f in sorted(glob.glob(REF + "/fuzz-collect/*.bz2")):
    add("fc-" + os.path.basename(f)[:10], bz2.decompress(open(f, "rb").read()), [9, 1])
for f in sorted(glob.glob(REF + "/manual-compress/*.bz2")):
    add("mc-" + os.path.basename(f)[:10], bz2.decompress(open(f, "rb").read()), [9, 2])
div = sorted(glob.glob(REF + "/fuzz-divbwt/*.bz2"))
periodic_ids = ("080828bf", "50c6d8e1", "6e84db59", "bc503b2b", "e075462d", "ed12e1a6", "de736e80", "df254dab")
sel = [f for i, f in enumerate(div) if i % 8 == 0 or os.path.basename(f).startswith(periodic_ids)]
sel += sorted(div, key=os.path.getsize)[-12:]
for f in sorted(set(sel)):
    add("fd-" + os.path.basename(f)[:10], bz2.decompress(open(f, "rb").read()), [9])
rng = np.random.default_rng(42)
add("syn-text-1M", synth.text(1_000_000), [9, 1])
add("syn-text-2p5M", synth.text(2_500_000, offset=3), [9])
add("syn-random-300k", synth.random_bytes(300_000), [9, 1])
add("syn-fib-400k", synth.fib(400_000), [9, 1])
add("syn-runs-fib-2M", synth.runs_and_fib(2_000_000), [9])
add("syn-zeros-1M", b"\0" * 1_000_000, [9, 1])
add("syn-lowent", bytes(rng.choice([65, 66, 67, 10], 500_000, p=[0.7, 0.2, 0.09, 0.01]).astype(np.uint8)), [9, 3])
add("syn-allbytes", bytes(range(256)) * 700, [9, 1])
add("syn-1byte", b"x", [9])
add("syn-2bytes", b"xy", [9])
add("syn-run259x3", b"q" * (259 * 3), [9])
add("syn-mixed", synth.text(150_000, offset=9) + b"\xff" * 3000 + synth.random_bytes(90_000, seed=5) + b"ab" * 40_000, [9, 1])

os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
manifest = {}
for name, (raw, levels) in sorted(cases.items()):
    with open(os.path.join(HERE, "inputs", name + ".bz2"), "wb") as f:
        f.write(bz2.compress(raw, 9))
    ent = {"len": len(raw), "sha256": hashlib.sha256(raw).hexdigest(), "levels": {}}
    for lv in levels:
        ref = orclib.ref_cli(raw, lv)
        orc, infos = orclib.orc_stream(raw, lv)
        periodic = any(i.tie_count > 1 for i in infos)
        ent["levels"][str(lv)] = {
            "ref_sha256": hashlib.sha256(ref).hexdigest(), "ref_len": len(ref), "blocks": len(infos),
            "periodic": periodic, "oracle_equals_ref": orc == ref}
        if orc != ref and not periodic:
            raise SystemExit("oracle differs from reference on aperiodic input %s -%d" % (name, lv))
    manifest[name] = ent
json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
tot = sum(os.path.getsize(p) for p in glob.glob(os.path.join(HERE, "inputs", "*.bz2")))
print("%d inputs, %d (input,level) cases, %.1f kB stored" % (len(manifest), sum(len(e["levels"]) for e in manifest.values()), tot / 1e3))
