"""GPU: the sharding building blocks of the decompressor (lbz_decoder_decode_at, lbz_walk_table,
lbz_decoder_emit_at) through lbzip2_b200.sharding.sharded_decompress with world = 1, and with the
candidates split by hand over two decoders on the same GPU (what two ranks would do), against the
decode goldens.  The world_size-2 process version runs on CPU (tests/test_unbz_shards.py)."""
import hashlib
import json
import os

import pytest

import orclib
import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "decode")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]


def test_sharded_decompress_world1_matches_goldens():
    import lbzip2_b200
    from lbzip2_b200 import api, sharding
    d = lbzip2_b200.Decoder(device=0, max_blocks=64, in_cap=1 << 20)
    pick = [c for c in MANIFEST if c["num_blocks"] >= 2 or c["name"] in ("planted magic", "magic in trailing garbage")]
    pick += [c for c in MANIFEST if c["status"] != "OK"][:: 6]
    for c in pick:
        z = open(os.path.join(GOLD, c["file"]), "rb").read()
        st, out, info = sharding.sharded_decompress(None, d, z, 0, 1, api.DBlock)
        if st == 3:
            assert c["status"] == "ERR_MAGIC"
            continue
        assert orclib.ERR_NAMES[st] == c["status"], (c["file"], c["name"], st)
        assert len(out) == c["out_len"] and hashlib.sha256(out).hexdigest() == c["out_sha256"], (c["file"], c["name"])
    d.close()


def test_two_decoders_share_one_file():
    import lbzip2_b200
    from lbzip2_b200 import api
    raw = synth.text(9_000_000, seed=123)
    eng = lbzip2_b200.Engine(device=0, level=9, max_chunks=12)
    z = eng.compress_stream(raw)
    eng.close()
    decs = [lbzip2_b200.Decoder(device=0, max_blocks=16, in_cap=len(z) + 64, out_cap=64 << 20) for _ in range(2)]
    hits = decs[0].scan(z)
    shares = [hits[r::2] for r in range(2)]
    tables = [decs[r].decode_at(z, shares[r]) for r in range(2)]
    merged = sorted(((b, r, i) for r in range(2) for i, b in enumerate(tables[r])), key=lambda t: t[0].pos)
    st, chain, chain_crc, info = decs[0].walk_table(z, [m[0] for m in merged])
    assert st == 0 and len(chain) == len(hits) == info.num_blocks
    goff, o = {}, 0
    for k in chain:
        goff[merged[k][0].pos] = o
        o += merged[k][0].out_len
    assert o == len(raw)
    out = bytearray(len(raw))
    for r in range(2):
        local, cur = [], 0
        for b in tables[r]:
            local.append(cur)
            cur += b.out_len
        payload, crcs = decs[r].emit_at(local, cur)
        for i, b in enumerate(tables[r]):
            out[goff[b.pos]: goff[b.pos] + b.out_len] = payload[local[i]: local[i] + b.out_len]
            assert crcs[i] == chain_crc[[merged[k][0].pos for k in chain].index(b.pos)]
    assert hashlib.sha256(bytes(out)).digest() == hashlib.sha256(raw).digest()
    for d in decs:
        d.close()
