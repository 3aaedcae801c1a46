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
    gpus = min(gpus, torch.cuda.device_count())
    env = {"LBZIP2_B200_BATCH": str(batch), "LBZIP2_B200_ENGINES": str(engines), "LBZIP2_B200_GPUS": str(gpus)}
    ins = _inputs()
    ins["text_9"] = synth.text(5_000_000, offset=34)
    for name, data in ins.items():
        level = 9 if name == "text_9" else (1 if len(data) < 1_000_000 else 3)
        assert _run(GPU_CLI, level, nthreads, data, env) == _reference(level, data), name
