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
