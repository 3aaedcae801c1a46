"""Decode-only probe for profiling: compress SIZE MB of the bench text once, then decompress it
REPS times through lbz_decompress_ex (resident input, device output).  Prints per-stage device ms.
usage: python tools/decode_probe.py [size_mb] [reps] [text|random|runs_fib] [concat]"""
import ctypes as C
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: F401
import synth
import lbzip2_b200
from lbzip2_b200 import api

size_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
concat = int(sys.argv[4]) if len(sys.argv) > 4 else 1     # decode the stream this many times over, as one concatenated file
n = size_mb * 1_000_000
t0 = time.time()
raw = synth.text(n) if kind == "text" else synth.random_bytes(n) if kind == "random" else synth.runs_and_fib(n)
eng = lbzip2_b200.Engine(device=0, level=9, max_chunks=min(128, n // 900000 + 1))
z = eng.compress_stream(raw)
eng.close()
print("input %d MB %s -> %d bytes (%.1fs)" % (size_mb, kind, len(z), time.time() - t0), flush=True)
z = z * concat
n = n * concat
L = lbzip2_b200.load_library()
nblk = 2 * (n // 900000 + 2)
dec = lbzip2_b200.Decoder(device=0, max_blocks=nblk, in_cap=len(z) + 64, out_cap=n + (1 << 20))
h_z = L.lbz_host_alloc(len(z))
C.memmove(h_z, z, len(z))
dec.load(h_z, len(z))
for i in range(reps):
    t = time.perf_counter()
    st, ln, info = dec.decompress_ptr(h_z, len(z), None, n + 64, api.D_RESIDENT_INPUT | api.D_DEVICE_OUTPUT)
    dt = (time.perf_counter() - t) * 1e3
    assert st == 0 and ln == n, (st, ln)
    print("rep %d: %.2f ms wall, %.2f ms device, %.0f MB/s, blocks %d, stages %s" % (
        i, dt, dec.last_ms, n / 1e6 / (dec.last_ms / 1e3), info.num_blocks,
        {k: round(v, 2) for k, v in dec.stage_ms().items()}), flush=True)
ok = hashlib.sha256(dec.array(api.DA_OUT, 0, n).tobytes()).digest() == hashlib.sha256(raw * concat).digest()
print("output sha256 equals input:", ok, "launches per call:", dec.launches // reps, "device MB:", dec.device_bytes >> 20)
dec.close()
