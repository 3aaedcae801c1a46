"""Host logic of the batch-aware task graph (lbzip2_b200/host/compress_b200.c).

CPU part: the C file is linked with the reference's unmodified scheduler/CLI and an
oracle-backed stand-in for the batch entry points (tests/stub/lbz_stub.c) --
oracle/_ref/lbzip2_b200_hosttest -- so staging, dynamic batching, reordering and CRC
folding are checked against the reference CLI without a GPU.  GPU part: the real
binary (oracle/_ref/lbzip2_b200, same C file + libbz2b200.so) against the same goldens.
"""
import os
import subprocess

import pytest

import orclib
import synth

CPU_CLI = os.path.join(orclib.REF_DIR, "lbzip2")
HOSTTEST = os.path.join(orclib.REF_DIR, "lbzip2_b200_hosttest")
GPU_CLI = os.path.join(orclib.REF_DIR, "lbzip2_b200")


def _inputs():
    t = synth.text(2_000_000, offset=31)
    return {
        "empty": b"",
        "one_byte": b"x",
        "exact_chunk_l1": synth.text(100_000, offset=32),
        "chunk_plus_one_l1": synth.text(100_001, offset=33),
        "mixed": t + b"\0" * 700_000 + synth.random_bytes(300_000, seed=31),
        # 4-byte runs expand under RLE1: two blocks per chunk (aperiodic on purpose: the
        # primary index of exactly periodic blocks is the documented parity exception)
        "spill_blocks": synth.text(700_000, offset=35).replace(b" ", b"    ")[:1_000_000],
    }


def _run(cli, level, nthreads, data, env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([cli, "-%d" % level, "-n%d" % nthreads], input=data, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, timeout=600)
    assert r.returncode == 0 and r.stderr == b"", r.stderr[-500:]
    return r.stdout


def _reference(level, data):
    return subprocess.run([CPU_CLI, "-%d" % level], input=data, stdout=subprocess.PIPE, check=True).stdout


@pytest.mark.parametrize("batch,engines,nthreads", [(1, 1, 1), (2, 2, 3), (5, 3, 8), (32, 2, 16)])
def test_task_graph_host_logic_matches_reference_cli(batch, engines, nthreads):
    if not (os.path.exists(CPU_CLI) and os.path.exists(HOSTTEST)):
        pytest.skip("oracle/_ref binaries not present")
    env = {"LBZIP2_B200_BATCH": str(batch), "LBZIP2_B200_ENGINES": str(engines)}
    for name, data in _inputs().items():
        level = 1 if (len(data) < 1_000_000 or name == "spill_blocks") else 3
        assert _run(HOSTTEST, level, nthreads, data, env) == _reference(level, data), name


def test_task_graph_rejects_sequential_collect_mode():
    if not os.path.exists(HOSTTEST):
        pytest.skip("oracle/_ref binaries not present")
    r = subprocess.run([HOSTTEST, "-u", "-1"], input=b"abc", stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert r.returncode != 0 and b"-u is not supported" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("batch,engines,nthreads,gpus", [(4, 2, 4, 1), (32, 2, 16, 1), (16, 2, 8, 2)])
def test_batch_cli_on_the_gpu_matches_reference_cli(batch, engines, nthreads, gpus):
    import torch
    if not (os.path.exists(CPU_CLI) and os.path.exists(GPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    if gpus > torch.cuda.device_count():
        pytest.skip("needs %d GPUs (run under gpurun --gpus %d)" % (gpus, gpus))
    env = {"LBZIP2_B200_BATCH": str(batch), "LBZIP2_B200_ENGINES": str(engines), "LBZIP2_B200_GPUS": str(gpus)}
    ins = _inputs()
    ins["text_9"] = synth.text(5_000_000, offset=34)
    for name, data in ins.items():
        level = 9 if name == "text_9" else (1 if len(data) < 1_000_000 else 3)
        assert _run(GPU_CLI, level, nthreads, data, env) == _reference(level, data), name


@pytest.mark.gpu
def test_batch_cli_sequential_mode_forwards_to_the_reference_collector():
    """`lbzip2_b200 -u`: the reference's own sequential collector (src/compress.c:120-198, linked in as
    compression_ref) drives the library's per-block API; output = the CPU reference's with -u."""
    if not (os.path.exists(CPU_CLI) and os.path.exists(GPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    data = synth.text(1_300_000, offset=51) + b"\0" * 2_500_000 + synth.random_bytes(200_000, seed=52) + b"k" * 600_000
    for level, nthreads in ((9, 4), (2, 8)):
        got = subprocess.run([GPU_CLI, "-u", "-%d" % level, "-n%d" % nthreads], input=data, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, timeout=300)
        assert got.returncode == 0, got.stderr[-400:]
        want = subprocess.run([CPU_CLI, "-u", "-%d" % level], input=data, stdout=subprocess.PIPE, check=True).stdout
        assert got.stdout == want


# ---- expansion task graph (lbzip2_b200/host/expand_b200.c) ---------------------
# CPU: the same C file + the reference scheduler, the decoder entry points coming from the host
# emulation of the decoder's own source (tests/simt_emul, test infrastructure).  GPU: the real
# binary.  Expectations: the committed decode goldens (= the reference CLI's behaviour).
import hashlib
import json

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode")


def _decode_cases():
    return json.load(open(os.path.join(_GOLD, "manifest.json")))["cases"]


def _check_expand(cli, cases, env_extra, nthreads):
    env = dict(os.environ, **env_extra)
    for c in cases:
        z = open(os.path.join(_GOLD, c["file"]), "rb").read()
        r = subprocess.run([cli, "-d", "-c", "-n%d" % nthreads], input=z, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, env=env, timeout=300)
        if c["status"] == "OK":
            assert r.returncode == 0, (c["file"], c["name"], r.stderr[-200:])
            assert hashlib.sha256(r.stdout).hexdigest() == c["out_sha256"], (c["file"], c["name"])
        else:
            text = orclib.ERR_TEXT[orclib.ERR_NAMES.index(c["status"])]
            assert r.returncode != 0 and text.encode() in r.stderr, (c["file"], c["name"], r.stderr[-200:])


@pytest.mark.parametrize("blocks,nthreads", [(320, 2), (2, 5)])
def test_expansion_task_graph_host_logic(blocks, nthreads):
    if not os.path.exists(HOSTTEST):
        pytest.skip("oracle/_ref binaries not present")
    cases = _decode_cases()
    cases = cases[:: 3] + [c for c in cases if c["status"] == "OK"][:12]
    _check_expand(HOSTTEST, cases, {"LBZIP2_B200_DBLOCKS": str(blocks), "LBZIP2_B200_DWAVE_MB": "48"}, nthreads)


def test_expansion_task_graph_round_trip_with_reference_cli():
    if not (os.path.exists(HOSTTEST) and os.path.exists(CPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    data = synth.text(1_300_000, offset=41) + b"\0" * 300_000 + synth.random_bytes(200_000, seed=41)
    z = _reference(1, data) + _reference(3, data[:400_000])          # two streams, 20 blocks
    want = subprocess.run([CPU_CLI, "-d", "-c"], input=z, stdout=subprocess.PIPE, check=True).stdout
    for blocks in ("3", "64"):
        r = subprocess.run([HOSTTEST, "-d", "-c", "-n4"], input=z, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           env=dict(os.environ, LBZIP2_B200_DBLOCKS=blocks, LBZIP2_B200_DWAVE_MB="48"), timeout=600)
        assert r.returncode == 0 and r.stdout == want == data + data[:400_000], r.stderr[-200:]


@pytest.mark.gpu
def test_expansion_cli_on_the_gpu_matches_goldens_and_reference_cli():
    if not os.path.exists(GPU_CLI):
        pytest.skip("oracle/_ref binaries not present")
    cases = _decode_cases()
    _check_expand(GPU_CLI, cases[:: 12], {"LBZIP2_B200_DBLOCKS": "16"}, 4)
    _check_expand(GPU_CLI, [c for c in cases if c["num_blocks"] >= 2][:8], {"LBZIP2_B200_DBLOCKS": "2"}, 3)
    data = synth.text(12_000_000, offset=42)
    z = _reference(9, data)
    r = subprocess.run([GPU_CLI, "-d", "-c", "-n8"], input=z, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0 and r.stdout == data, r.stderr[-200:]


@pytest.mark.gpu
def test_unmodified_expand_c_drives_the_gpu_decoder():
    """oracle/_ref/lbzip2_gpu links the reference's main.c/process.c/expand.c/parse.c WITHOUT src/decode.c:
    decoder_init/retrieve/decode/emit/decoder_free (src/decode.h:72-81) come from libbz2b200.so.  The
    reference's scheduler feeds retrieve() one 256 KiB buffer at a time and takes the block end from the
    bit cursor; verdicts, messages and bytes must be the reference CLI's (the committed goldens)."""
    shim_cli = os.path.join(orclib.REF_DIR, "lbzip2_gpu")
    if not (os.path.exists(shim_cli) and os.path.exists(CPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    cases = _decode_cases()
    pick = cases[:: 9] + [c for c in cases if c["status"] == "OK" and c["num_blocks"] >= 2][:6]
    # which error of a damaged file is reported first depends on the scheduling of the reference's own
    # scanner / parser / retriever tasks (its CLI says "bad number of trees" with -n1 and "bad block header
    # magic" with -n2 for golden 03c2e5d6...): the goldens were made with one worker, so are these runs
    _check_expand(shim_cli, [c for c in pick if c["status"] != "OK"], {}, 1)
    _check_expand(shim_cli, [c for c in pick if c["status"] == "OK"], {}, 4)
    data = synth.text(6_000_000, offset=43) + synth.random_bytes(2_500_000, seed=43) + b"\0" * 4_000_000
    for level, nthreads in ((9, 8), (1, 16)):
        z = _reference(level, data)
        r = subprocess.run([shim_cli, "-d", "-c", "-n%d" % nthreads], input=z, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert r.returncode == 0 and r.stdout == data, r.stderr[-200:]


def test_expansion_task_graph_several_operands_and_thread_counts(tmp_path):
    """One process, several files (the decoder outlives an operand and is re-sized for a larger
    one, src/main.c:935), -n from 1 to 64, small and large waves."""
    if not (os.path.exists(HOSTTEST) and os.path.exists(CPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    files = {"a": synth.text(200_000, offset=5), "b": synth.text(2_200_000, offset=6), "c": b""}
    for name, data in files.items():
        (tmp_path / (name + ".bz2")).write_bytes(_reference(2, data))
    args = [str(tmp_path / (n + ".bz2")) for n in ("a", "b", "c")]
    r = subprocess.run([HOSTTEST, "-d", "-k", "-f"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0 and r.stderr == b"", r.stderr[-300:]
    for name, data in files.items():
        assert (tmp_path / name).read_bytes() == data
    z = (tmp_path / "b.bz2").read_bytes()
    for n, blocks in ((1, "1"), (3, "7"), (64, "320")):
        r = subprocess.run([HOSTTEST, "-d", "-c", "-n%d" % n], input=z, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           env=dict(os.environ, LBZIP2_B200_DBLOCKS=blocks, LBZIP2_B200_DWAVE_MB="48"), timeout=600)
        assert r.returncode == 0 and r.stdout == files["b"], (n, blocks, r.stderr[-200:])


def test_expansion_task_graph_streams_through_a_window_smaller_than_the_file():
    """6 MB of incompressible data at -1 (61 blocks of 100 KB, 6.03 MB compressed) through the minimum
    window of 4 MB (LBZIP2_B200_DWINDOW_MB=2 is raised to it): the decoder must drop the consumed front
    of its window and the stage task must wait for room, from a pipe (size unknown) and from a file."""
    if not (os.path.exists(HOSTTEST) and os.path.exists(CPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    data = synth.random_bytes(6_000_000, seed=77) + synth.text(300_000, offset=9)
    z = _reference(1, data)
    assert len(z) > 5_000_000
    env = dict(os.environ, LBZIP2_B200_DWINDOW_MB="2", LBZIP2_B200_DBLOCKS="16", LBZIP2_B200_DWAVE_MB="48", LBZIP2_B200_STATS="1")
    for n in (1, 4):
        r = subprocess.run([HOSTTEST, "-d", "-c", "-n%d" % n], input=z, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=900)
        assert r.returncode == 0 and r.stdout == data, r.stderr[-300:]
        assert b"through a window of 4194304" in r.stderr, r.stderr[-300:]
    # a damaged block in the last third: the bytes in front of it are still written, then the reference's message
    bad = bytearray(z)
    bad[len(z) * 2 // 3] ^= 0x10
    r = subprocess.run([HOSTTEST, "-d", "-c", "-n3"], input=bytes(bad), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=900)
    ref = subprocess.run([CPU_CLI, "-d", "-c", "-n1"], input=bytes(bad), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert r.returncode != 0 and ref.returncode != 0
    assert len(r.stdout) >= 3_000_000 and data.startswith(r.stdout)
    assert b"compressed data error" in r.stderr or b"block CRC mismatch" in r.stderr, r.stderr[-300:]


def test_expansion_task_graph_drops_the_input_behind_its_verdict():
    """Once the decoder has its verdict -- the end of the last stream with megabytes of garbage behind it,
    or a damaged block -- the rest of the input is read and dropped (a bounded staging buffer must not
    dam up the reader: found as a deadlock by review).  Code, bytes and message of the reference CLI."""
    if not (os.path.exists(HOSTTEST) and os.path.exists(CPU_CLI)):
        pytest.skip("oracle/_ref binaries not present")
    data = synth.text(300_000, offset=3)
    z = _reference(1, data)
    garbage = z + synth.random_bytes(12_000_000, seed=5)
    bad = bytearray(z + z)
    bad[len(z) // 2] ^= 0x20
    bad = bytes(bad) + synth.random_bytes(9_000_000, seed=6)
    env = dict(os.environ, LBZIP2_B200_DWINDOW_MB="2", LBZIP2_B200_DBLOCKS="8", LBZIP2_B200_DWAVE_MB="48")
    for inp in (garbage, bad):
        want = subprocess.run([CPU_CLI, "-d", "-c", "-n1"], input=inp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        for n in (1, 4):
            got = subprocess.run([HOSTTEST, "-d", "-c", "-n%d" % n], input=inp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
            assert got.returncode == want.returncode and got.stdout == want.stdout
            assert got.stderr.split(b": ", 1)[-1] == want.stderr.split(b": ", 1)[-1], (got.stderr, want.stderr)
