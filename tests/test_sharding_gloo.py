"""CPU, world_size 2 over gloo: the multi-GPU sharding/gather logic
(lbzip2_b200/sharding.py) reassembles the exact single-stream output.  The
per-rank compressor here is the oracle (test infrastructure); on GPUs bench.py
uses the CUDA engine with the same sharding code."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orclib
import synth

HERE = os.path.dirname(os.path.abspath(__file__))


class _Rec:
    def __init__(self, raw_offset, out_len, crc):
        self.raw_offset, self.out_len, self.crc = raw_offset, out_len, crc


def _worker(rank, world, path, level, data, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    from lbzip2_b200 import sharding
    dist.init_process_group("gloo", init_method="file://" + path, rank=rank, world_size=world)
    mbs = level * 100000
    mine = sharding.rank_input(data, mbs, world, rank)
    stream, infos = orclib.orc_stream(mine.tobytes(), level)
    payload = np.frombuffer(stream[4:-10], dtype=np.uint8)
    recs, off = [], 0
    for i in infos:
        recs.append(_Rec(off, i.out_len, i.block_crc))
        off += i.consumed
    table = sharding.block_table(recs, mbs)
    tables, payloads = sharding.gather_blocks(table, torch.from_numpy(payload.copy()), dist, "cpu")
    # the preallocated gatherer bench.py uses on GPUs (NCCL) must agree; run it twice (buffer reuse)
    gth = sharding.BlockGatherer(dist, "cpu", max_blocks=64, max_payload=2_000_000)
    for _ in range(2):
        t2, p2 = gth.gather(table, torch.from_numpy(payload.copy()))
    t3, p3 = gth.gather(table, torch.from_numpy(payload.copy()), tables_only=True)
    # the host sink of the N>1 end-to-end leg: every rank puts its own blocks at their stream
    # offsets in one shared mapping (on GPUs the copies are lbz_scatter_to_host from HBM)
    sink = sharding.SharedStream(dist, "cpu", path + ".stream", 2_000_000, max_blocks=64, register=False)
    shared = None
    for _ in range(2):
        every = sink.exchange(table)
        offs, total, cc = sharding.place_blocks(every, world)
        src = np.concatenate(([0], np.cumsum(table[:, 1])))
        for k in range(len(table)):
            sink.view[offs[rank][k]:offs[rank][k] + table[k, 1]] = payload[src[k]:src[k + 1]]
        end = sink.finish(level, total, cc)
        if rank == 0:
            shared = bytes(sink.view[:end])
    sink.close()
    if rank == 0:
        stream = sharding.assemble_stream(level, tables, payloads, world)
        assert shared == stream
        assert sharding.assemble_stream(level, *sharding.to_host(t2, p2), world) == stream
        assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(t2, t3)) and p3 == [None] * world
        q.put(stream)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_reassembles_stream():
    level = 1
    data = synth.text(530_000, offset=2) + b"\x07" * 120_000 + synth.random_bytes(75_000, seed=3)
    want, _ = orclib.orc_stream(data, level)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "rdzv")
        procs = [ctx.Process(target=_worker, args=(r, 2, path, level, data, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = q.get(timeout=120)
        for p in procs:
            p.join(60)
    assert got == want


def test_chunk_assignment_round_robin():
    from lbzip2_b200 import sharding
    assert sharding.rank_chunk_ids(950_000, 100_000, 4, 1) == [1, 5, 9]
    assert sharding.rank_chunk_ids(100_000, 100_000, 8, 3) == []
    a = np.arange(250_000, dtype=np.uint32).astype(np.uint8)
    parts = [sharding.rank_input(a, 100_000, 2, r) for r in range(2)]
    assert parts[0].size == 150_000 and parts[1].size == 100_000
