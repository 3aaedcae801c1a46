"""ctypes bindings for tests/simt_emul/libunbz_emul.so -- TEST INFRASTRUCTURE ONLY.

The library is the decompressor's device + orchestration source compiled for the host with every
kernel launch run as a sequential loop (see tests/simt_emul/unbz_emul.cpp).  It exists so that the
decode logic can be checked against the oracle where there is no GPU; the product never loads it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "simt_emul")
CSRC = os.path.join(ROOT, "lbzip2_b200", "csrc")
u8p = C.POINTER(C.c_uint8)


class DStreamInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "status", "num_blocks", "num_streams", "bad_block", "garbage", "candidates",
        "false_candidates", "waves")] + [("end_bit", C.c_uint64)]


class DBlock(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("pos", "end_bit", "out_len", "out_off")] + [
        (n, C.c_uint32) for n in ("status", "rand", "bwt_idx", "block_size", "alpha_size", "num_trees",
                                  "num_selectors", "period", "rl_state", "crc_acc", "crc", "ntok", "nsym", "ngrp")] + [
        ("sym_bit", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(EMU_DIR, "libunbz_emul.so")
        deps = [os.path.join(EMU_DIR, "unbz_emul.cpp"), os.path.join(CSRC, "unbz_kernels.cuh"),
                os.path.join(CSRC, "unbz_engine.inc"), os.path.join(ROOT, "include", "lbzip2_b200.h")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + CSRC, "-o", so, deps[0]])
        L = C.CDLL(so)
        L.emu_decoder_create.restype = C.c_void_p
        L.emu_decoder_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
        L.emu_decoder_destroy.argtypes = [C.c_void_p]
        L.emu_decompress.restype = C.c_int
        L.emu_decompress.argtypes = [C.c_void_p, u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t),
                                     C.POINTER(DStreamInfo), C.c_uint]
        L.emu_scan_blocks.restype = C.c_long
        L.emu_scan_blocks.argtypes = [C.c_void_p, u8p, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t]
        L.emu_decoder_read.restype = C.c_int
        L.emu_decoder_read.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_size_t]
        L.emu_last_wave_blocks.restype = C.c_uint32
        L.emu_last_wave_blocks.argtypes = [C.c_void_p]
        L.emu_launches.restype = C.c_uint64
        L.emu_launches.argtypes = [C.c_void_p]
        L.emu_mtf_front.restype = C.c_uint32
        L.emu_mtf_front.argtypes = [C.POINTER(C.c_uint32), C.c_uint32]
        L.emu_gf_shift.restype = C.c_uint32
        L.emu_gf_shift.argtypes = [C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32)]
        L.emu_set_reverse.argtypes = [C.c_int]
        L.emu_decode_at.restype = C.c_int
        L.emu_decode_at.argtypes = [C.c_void_p, u8p, C.c_size_t, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(DBlock), C.c_uint]
        L.emu_emit_at.restype = C.c_int
        L.emu_emit_at.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_size_t,
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
        L.emu_walk_table.restype = C.c_int
        L.emu_walk_table.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(DBlock), C.c_size_t, C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_size_t), C.POINTER(DStreamInfo)]
        L.emu_decoder_open_stream.restype = C.c_int
        L.emu_decoder_open_stream.argtypes = [C.c_void_p, C.c_uint]
        L.emu_decoder_feed.restype = C.c_int
        L.emu_decoder_feed.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
        L.emu_decoder_next.restype = C.c_int
        L.emu_decoder_next.argtypes = [C.c_void_p, u8p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(DStreamInfo)]
        _lib = L
    return _lib


class EmuDecoder:
    """Same surface as lbzip2_b200.Decoder, over the emulation."""

    def __init__(self, max_blocks=8, in_cap=1 << 22, out_cap=0):
        self.L = lib()
        self.h = self.L.emu_decoder_create(max_blocks, in_cap, out_cap)
        assert self.h

    def close(self):
        if self.h:
            self.L.emu_decoder_destroy(self.h)
            self.h = None

    def decompress(self, z, cap=None):
        a = np.frombuffer(bytes(z), dtype=np.uint8) if len(z) else np.zeros(1, np.uint8)
        if cap is None:
            cap = max(1 << 20, 64 * len(z))
        out = np.empty(max(cap, 1), np.uint8)
        n = C.c_size_t(0)
        info = DStreamInfo()
        st = self.L.emu_decompress(self.h, a.ctypes.data_as(u8p), len(z), out.ctypes.data_as(u8p), cap,
                                   C.byref(n), C.byref(info), 0)
        return st, out[: n.value].tobytes(), info

    def decompress_pieces(self, pieces, wave_cap, trace=None, greedy=True):
        """Streaming session (open_stream / feed / next) over the parts of one file; same driver logic
        as lbzip2_b200.Decoder.decompress_pieces.  trace (a list) receives the status of every next();
        greedy=False hands over one piece per request (the decoder starves often)."""
        assert self.L.emu_decoder_open_stream(self.h, 0) == 0
        buf = np.empty(max(wave_cap, 1), np.uint8)
        out, info = [], DStreamInfo()
        it = iter(pieces)
        state = {"cur": np.zeros(0, np.uint8), "off": 0, "eof": False}

        def feed():
            progressed = False
            while not state["eof"]:
                cur, off = state["cur"], state["off"]
                if off >= cur.size:
                    nxt = next(it, None)
                    if nxt is None:
                        took = C.c_size_t(0)
                        assert self.L.emu_decoder_feed(self.h, buf.ctypes.data_as(u8p), 0, 1, C.byref(took)) == 0
                        state["eof"] = True
                        return True
                    state["cur"], state["off"] = np.frombuffer(bytes(nxt), dtype=np.uint8), 0
                    continue
                took = C.c_size_t(0)
                part = cur[off:]
                assert self.L.emu_decoder_feed(self.h, part.ctypes.data_as(u8p), part.size, 0, C.byref(took)) == 0
                state["off"] = off + took.value
                progressed = progressed or took.value > 0
                if state["off"] < cur.size or not greedy:
                    return progressed
            return progressed

        feed()
        while True:
            n = C.c_size_t(0)
            st = self.L.emu_decoder_next(self.h, buf.ctypes.data_as(u8p), wave_cap, C.byref(n), C.byref(info))
            assert st >= 0, "emu_decoder_next failed"
            if trace is not None:
                trace.append(st)
            if n.value:
                out.append(buf[: n.value].tobytes())
            if st == 1:
                feed()
                continue
            if st == 101:
                assert feed(), "the decoder wants input but its window is full"
                continue
            return st, b"".join(out), info

    def scan(self, z):
        a = np.frombuffer(bytes(z), dtype=np.uint8) if len(z) else np.zeros(1, np.uint8)
        cap = len(z) // 6 + 64
        pos = (C.c_uint64 * cap)()
        k = self.L.emu_scan_blocks(self.h, a.ctypes.data_as(u8p), len(z), pos, cap)
        assert k >= 0
        return list(pos[:k])

    def block(self, slot):
        b = DBlock()
        assert self.L.emu_decoder_read(self.h, 0, slot, C.byref(b), C.sizeof(b)) == 0
        return b

    def array(self, which, slot, nbytes):
        a = np.empty(max(nbytes, 1), np.uint8)
        assert self.L.emu_decoder_read(self.h, which, slot, a.ctypes.data_as(C.c_void_p), nbytes) == 0
        return a[:nbytes]

    # the sharding building blocks, same surface as lbzip2_b200.Decoder
    def decode_at(self, z, positions):
        a = np.frombuffer(bytes(z), dtype=np.uint8) if len(z) else np.zeros(1, np.uint8)
        k = len(positions)
        pos = (C.c_uint64 * max(k, 1))(*positions)
        table = (DBlock * max(k, 1))()
        assert self.L.emu_decode_at(self.h, a.ctypes.data_as(u8p), len(z), pos, k, table, 0) == 0
        out = []
        for i in range(k):
            c = DBlock()
            C.memmove(C.byref(c), C.byref(table[i]), C.sizeof(DBlock))
            out.append(c)
        return out

    def emit_at(self, out_offs, cap, out=None):
        k = len(out_offs)
        offs = (C.c_uint64 * max(k, 1))(*out_offs)
        crc = (C.c_uint32 * max(k, 1))()
        assert out is None or out.size >= cap
        buf = out if out is not None else np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        assert self.L.emu_emit_at(self.h, offs, k, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), crc) == 0
        return (buf[: n.value] if out is not None else buf[: n.value].tobytes()), list(crc[:k])

    def walk_table(self, z, table):
        a = np.frombuffer(bytes(z), dtype=np.uint8) if len(z) else np.zeros(1, np.uint8)
        k = len(table)
        arr = (DBlock * max(k, 1))(*table)
        chain = (C.c_uint32 * max(k, 1))()
        ccrc = (C.c_uint32 * max(k, 1))()
        n = C.c_size_t(0)
        info = DStreamInfo()
        st = self.L.emu_walk_table(a.ctypes.data_as(C.c_void_p), len(z), arr, k, chain, ccrc, C.byref(n), C.byref(info))
        return st, list(chain[: n.value]), list(ccrc[: n.value]), info

    @property
    def last_wave_blocks(self):
        return self.L.emu_last_wave_blocks(self.h)
