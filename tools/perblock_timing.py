"""Latency of the reference-shaped per-block API (one thread), per call."""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import lbzip2_b200, synth
L = lbzip2_b200.load_library()
raw = synth.text(5_400_000, offset=3)
mbs = 900000
pos = 0
k = 0
while pos < len(raw):
    st = C.create_string_buffer(L.encoder_alloc_size(mbs))
    t0 = time.perf_counter(); L.encoder_init(st, mbs, 8)
    chunk = raw[pos:pos + mbs]; cbuf = C.create_string_buffer(chunk, len(chunk)); left = C.c_size_t(len(chunk))
    t1 = time.perf_counter(); L.collect(st, cbuf, C.byref(left))
    t2 = time.perf_counter(); crc = C.c_uint32(0); size = L.encode(st, C.byref(crc))
    t3 = time.perf_counter(); out = C.create_string_buffer((size + 3) // 4 * 4); L.transmit(st, out)
    t4 = time.perf_counter()
    print("block %d: init %.2f ms collect %.2f ms encode %.2f ms transmit %.2f ms (consumed %d)" % (k, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, len(chunk)-left.value))
    pos += len(chunk) - left.value; k += 1
