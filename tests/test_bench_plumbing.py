"""CPU: bench.py's decompression leg runs end to end (JSON keys, verification, the 4x-blocks run)
with the decoder replaced by the host emulation and CUDA events by a wall clock.  Guards the
plumbing of the bench line, not any number."""
import bz2
import ctypes as C
import os
import sys
import time
import types

import numpy as np

import emulib
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decompress_leg_plumbing(monkeypatch):
    class Ev:
        def __init__(self, enable_timing=True):
            self.t = 0.0

        def record(self):
            self.t = time.perf_counter()

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    fake_torch = types.ModuleType("torch")
    fake_torch.cuda = types.SimpleNamespace(Event=Ev, synchronize=lambda: None)
    monkeypatch.setitem(sys.modules, "torch", fake_torch)
    monkeypatch.syspath_prepend(ROOT)
    import bench

    class FakeDecoder:
        def __init__(self, device=0, max_blocks=8, in_cap=0, out_cap=0):
            self.e = emulib.EmuDecoder(max_blocks=max_blocks, in_cap=in_cap, out_cap=out_cap)
            self.launches, self.last_ms, self.device_bytes = 0, 1.0, 1

        def load(self, ptr, n):
            pass

        def decompress_ptr(self, in_ptr, n, out_ptr, out_cap, flags=0):
            z = bytes((C.c_uint8 * n).from_address(in_ptr))
            st, out, info = self.e.decompress(z, cap=out_cap)
            self.out = out
            self.launches += 26
            if out_ptr:
                C.memmove(out_ptr, out, len(out))
            return st, len(out), info

        def array(self, which, slot, nbytes):
            return np.frombuffer(self.out[slot: slot + nbytes], np.uint8)

        def stage_ms(self):
            return dict(zip(("upload", "scan", "retrieve", "successors", "walks", "expand", "tail"), [0.1] * 7))

        def close(self):
            self.e.close()

    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    L = types.SimpleNamespace(lbz_host_alloc=lambda n: libc.malloc(n), lbz_host_free=lambda p: libc.free(p))
    data = synth.text(250_000)
    stream = bz2.compress(data, 1)
    recs = [types.SimpleNamespace(nblock=100_000, nmtf=40_000) for _ in range(3)]
    a = types.SimpleNamespace(steps=2, warmup=1, no_cpu_baseline=True, workload="text", level=9, size_mb=100)
    res = bench.decompress_leg(a, types.SimpleNamespace(Decoder=FakeDecoder), L, 0, stream, data, recs)
    assert res["verified"] == {"device_output_sha256_equals_input": True, "host_output_equals_input": True}
    assert res["blocks"] == 3 and res["waves"] == 1 and res["gpu_launches"] == 26
    assert res["more_blocks_in_flight"]["blocks"] == 12 and res["more_blocks_in_flight"]["last_copy_sha256_equals_input"]
    for key in ("value", "e2e", "roofline", "path_roofline", "stage_ms"):
        assert key in res


def test_traffic_comes_from_the_committed_captures(monkeypatch):
    """`roofline.traffic` is looked up in profiles/ncu_traffic.json by kernel and element count: the default
    workload (100 MB text = 100.14 M rotations) must find the newest capture of the plain pass, whose DRAM
    bytes are the algorithmic 16 B per element within a few per cent; an unknown size gives null."""
    monkeypatch.syspath_prepend(ROOT)
    import bench
    t = bench.ncu_traffic("k_text_pass2", 100136441)
    assert t is not None and abs(t - 16 * 100136441) < 0.05 * 16 * 100136441
    assert bench.ncu_traffic("k_text_pass2", 55_000_000) is None
    assert bench.ncu_traffic("k_ub_chain2", 223) is not None
    assert bench.ncu_traffic("no_such_kernel", 100136441) is None


def test_both_arms_name_the_workload_alike(monkeypatch):
    monkeypatch.syspath_prepend(ROOT)
    import bench
    import types as _t
    for wl, mb in (("text", 100), ("text", 1250), ("random", 1000), ("runs_fib", 1000)):
        a = _t.SimpleNamespace(workload=wl, size_mb=mb, level=9)
        name = bench.workload_name(a)
        assert str(mb) in name and "-9" in name
