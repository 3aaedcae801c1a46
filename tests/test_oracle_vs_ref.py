"""Dev-container only: stage-by-stage comparison of the oracle restatement with
the compiled, unmodified reference (oracle/_ref/libref_stages.so)."""
import numpy as np
import pytest

import golden_util
import orclib
import synth

pytestmark = pytest.mark.ref
if not orclib.have_ref():
    pytest.skip("oracle/_ref not built (needs /root/reference)", allow_module_level=True)


def _compare(raw, cap):
    o = orclib.orc_block_stages(raw, cap)
    r = orclib.ref_block_stages(raw, cap)
    assert o["consumed"] == r["consumed"] and o["nblock"] == r["nblock"] and o["crc"] == r["crc"]
    assert np.array_equal(o["block"], r["block"]) and np.array_equal(o["used"] != 0, r["used"] != 0)
    assert np.array_equal(o["bwt"], r["bwt"])
    if o["tie_count"] == 1:
        assert o["bwt_idx"] == r["bwt_idx"]
    assert o["nmtf"] == r["nmtf"] and np.array_equal(o["mtfv"], r["mtfv"])
    cd = o["coding"]
    assert cd.num_trees == r["num_trees"] and cd.num_selectors == r["num_selectors"]
    assert cd.tree_pad == r["tree_pad"] and cd.out_len == r["out_len"]
    for t in range(cd.num_trees):
        old = r["new2old"][t]
        assert bytes(cd.length[t])[: o["alpha_size"]] == bytes(r["length_old"][old][: o["alpha_size"]])
    assert bytes(cd.selector_mtf)[: cd.num_selectors] == bytes(r["selector_mtf"])
    if o["tie_count"] == 1:
        assert np.array_equal(o["bits"], r["bits"])


def test_stages_on_fixture_sample():
    man = golden_util.manifest()
    names = sorted(man)[::5]
    for n in names:
        _compare(golden_util.load_input(n), 900000)
        _compare(golden_util.load_input(n), 100000)


def test_stages_on_synthetic():
    _compare(synth.text(900000), 900000)
    _compare(synth.random_bytes(120000), 100000)
    _compare(synth.fib(100000), 100000)
    _compare(b"a" * 900000, 900000)
    _compare(bytes(range(256)) * 100 + b"zz" * 300, 100000)


def test_stages_with_arbitrary_block_sizes():
    """encoder_init takes any max_block_size in 1..900000 (src/encode.c:121-122), not only
    level * 100000: pins the oracle for the sizes tests/test_gpu_parity.py drives the GPU with."""
    for cap, raw in [(1, b"a"), (7, b"aaaaaaa"), (7, b"aaaabcd"), (1000, synth.text(1000, offset=31)),
                     (4099, b"r" * 4099), (123457, synth.text(123457, offset=32)),
                     (899_999, synth.random_bytes(899_999, seed=33)), (12345, b"\x00" * 9000 + synth.fib(3345))]:
        _compare(raw, cap)
