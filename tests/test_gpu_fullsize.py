"""BASELINE.json configs 3-5 at their stated size, bit-exact against the reference CLI.

  config 4: 1 GB of uniform random bytes (MTF / Huffman stress)
  config 5: 1 GB = 500 MB of one byte value + 500 MB of the Fibonacci word (tests/fib.c:28-33)
  config 3: one GPU's share of the 10 GB text job over 8 GPUs (1.25 GB: chunks are sharded
            round-robin and every shard is compressed on its own)

The stream is produced through the C ABI (lbz_compress_stream: pinned-size batches of 560
chunks inside one call) and compared by sha256 with `oracle/_ref/lbzip2 -9` on the same bytes."""
import hashlib
import os
import subprocess

import pytest

import orclib
import synth

pytestmark = pytest.mark.gpu

import lbzip2_b200  # noqa: E402

MB = 1_000_000
CLI = os.path.join(orclib.REF_DIR, "lbzip2")


def reference_sha(data, level=9):
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(tmp, "lbz_fullsize_%d.raw" % os.getpid())
    with open(path, "wb") as f:
        f.write(data)
    try:
        h = hashlib.sha256()
        p = subprocess.Popen([CLI, "-%d" % level, "-c", path], stdout=subprocess.PIPE)
        for blk in iter(lambda: p.stdout.read(1 << 24), b""):
            h.update(blk)
        assert p.wait() == 0
        return h.hexdigest()
    finally:
        os.unlink(path)


def run_case(data):
    if not os.path.exists(CLI):
        pytest.skip("oracle/_ref/lbzip2 not present")
    eng = lbzip2_b200.Engine(device=0, level=9, max_chunks=560)
    try:
        got = eng.compress_stream(data)
    finally:
        eng.close()
    assert hashlib.sha256(got).hexdigest() == reference_sha(data)


def test_config4_random_1gb():
    run_case(synth.random_bytes(1000 * MB, seed=1))


def test_config5_runs_fib_1gb():
    run_case(synth.runs_and_fib(1000 * MB))


def test_config3_text_shard_1250mb():
    run_case(synth.text_streams(1250 * MB, first_offset=1000))
