"""Throughput of the reference-shaped per-block API from N host threads."""
import ctypes as C, sys, time, os, threading
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import lbzip2_b200, synth
L = lbzip2_b200.load_library()
raw = synth.text(900_000 * 8, offset=3)
mbs = 900000
chunks = [raw[i:i + mbs] for i in range(0, len(raw), mbs)]

def one(chunk):
    st = C.create_string_buffer(L.encoder_alloc_size(mbs))
    L.encoder_init(st, mbs, 8)
    cbuf = C.create_string_buffer(chunk, len(chunk)); left = C.c_size_t(len(chunk))
    L.collect(st, cbuf, C.byref(left))
    crc = C.c_uint32(0); size = L.encode(st, C.byref(crc))
    out = C.create_string_buffer((size + 3) // 4 * 4); L.transmit(st, out)

t0 = time.perf_counter(); e = lbzip2_b200.Engine(device=0, level=9, max_chunks=1); t1 = time.perf_counter()
e2 = lbzip2_b200.Engine(device=0, level=9, max_chunks=1); t2 = time.perf_counter()
print("engine create: first %.1f ms, second %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
one(chunks[0])
for nt in (1, 4, 16, 32):
    reps = 4
    def work(k):
        for r in range(reps):
            one(chunks[(k + r) % len(chunks)])
    # warm the pool with nt contexts
    ths = [threading.Thread(target=one, args=(chunks[k % len(chunks)],)) for k in range(nt)]
    tw = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; tw = time.perf_counter() - tw
    ths = [threading.Thread(target=work, args=(k,)) for k in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
    print("threads %2d: warm-up (context creation) %.0f ms; %d blocks in %.1f ms = %.1f ms/block wall, %.0f MB/s" % (nt, tw * 1e3, nt * reps, dt * 1e3, dt * 1e3 / (nt * reps), nt * reps * 0.9 / dt))
