"""CPU: sharding the blocks of one .bz2 file over several decoders (lbzip2_b200/sharding.py
sharded_decompress).  The decoder is the host emulation of the product's own source (tests/simt_emul);
one process with world = 1, and world_size-2 gloo processes.  Expectations: the decode goldens."""
import hashlib
import json
import os
import sys

import pytest
import torch.multiprocessing as mp

import emulib
import orclib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "decode")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]


def _cases():
    pick = [c for c in MANIFEST if c["num_blocks"] >= 2 or c["name"] in ("planted magic", "magic in trailing garbage")]
    pick += [c for c in MANIFEST if c["status"] != "OK"][:: 6]
    return pick


def _check(dist, rank, world):
    sys.path.insert(0, ROOT)
    from lbzip2_b200 import sharding
    d = emulib.EmuDecoder(max_blocks=64, in_cap=1 << 20)
    for c in _cases():
        z = open(os.path.join(GOLD, c["file"]), "rb").read()
        st, out, info = sharding.sharded_decompress(dist, d, z, rank, world, emulib.DBlock)
        if st == 3:                       # not a bzip2 file: nothing to shard
            assert c["status"] == "ERR_MAGIC"
            continue
        assert orclib.ERR_NAMES[st] == c["status"], (c["file"], c["name"], st)
        if rank == 0:
            assert len(out) == c["out_len"] and hashlib.sha256(out).hexdigest() == c["out_sha256"], (c["file"], c["name"])
            assert info.num_blocks == c["num_blocks"]
    d.close()


def test_sharded_decompress_single_process():
    _check(None, 0, 1)


def _worker(rank, world, port):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        _check(dist, rank, world)
    finally:
        dist.destroy_process_group()


def test_sharded_decompress_gloo_world2():
    mp.spawn(_worker, args=(2, 29611), nprocs=2, join=True)


def _check_parts_only(dist, rank, world):
    """The shape bench.py times at N > 1: the decoded bytes stay on every rank (in a caller-provided
    array), only (offset, length, CRC) per block travel; the parts of all ranks tile the output."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from lbzip2_b200 import sharding
    d = emulib.EmuDecoder(max_blocks=64, in_cap=1 << 20)
    mine = np.empty(48 << 20, dtype=np.uint8)
    for c in [c for c in MANIFEST if c["status"] == "OK" and c["num_blocks"] >= 2][:10]:
        z = np.frombuffer(open(os.path.join(GOLD, c["file"]), "rb").read(), dtype=np.uint8)      # an array, like the shared mapping
        keep = {}
        st, out, info = sharding.sharded_decompress(dist, d, z, rank, world, emulib.DBlock, gather_payload=False, keep=keep, out=mine)
        assert st == 0 and out is None and info.num_blocks == c["num_blocks"]
        parts = [(g, ln, bytes(keep["payload"][lo:lo + ln])) for g, ln, lo, _ in keep["parts"]]
        everyone = [None] * world
        if world > 1:
            dist.all_gather_object(everyone, parts)
        else:
            everyone = [parts]
        whole = bytearray(c["out_len"])
        covered = 0
        for prt in everyone:
            for g, ln, data in prt:
                whole[g:g + ln] = data
                covered += ln
        assert covered == c["out_len"] and hashlib.sha256(bytes(whole)).hexdigest() == c["out_sha256"], c["file"]
    d.close()


def test_sharded_decompress_parts_stay_on_the_ranks_single_process():
    _check_parts_only(None, 0, 1)


def _worker_parts(rank, world, port):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        _check_parts_only(dist, rank, world)
    finally:
        dist.destroy_process_group()


def test_sharded_decompress_parts_stay_on_the_ranks_gloo_world2():
    mp.spawn(_worker_parts, args=(2, 29613), nprocs=2, join=True)
