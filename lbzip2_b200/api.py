"""ctypes mirror of include/lbzip2_b200.h (host-side convenience only)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LbzError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libbz2b200.so")


class BlockRec(C.Structure):
    _fields_ = [("raw_offset", C.c_uint64)] + [
        (n, C.c_uint32) for n in ("raw_len", "nblock", "crc", "bwt_idx", "tie_count", "nmtf",
                                  "num_trees", "num_selectors", "out_len", "reserved")]


class BlockMeta(C.Structure):
    """Mirror of struct LbzBlockMeta (csrc/lbz_common.cuh)."""
    _fields_ = [(n, C.c_uint32) for n in (
        "n", "raw_len", "crc", "bwt_idx", "tie_count", "nmtf", "alpha_size", "num_trees",
        "num_selectors", "tree_pad", "out_len", "unsorted", "depth", "tree_cost")] + [
        ("used", C.c_uint32 * 8), ("pad_", C.c_uint32 * 2)] + [
        (n, C.c_uint32) for n in ("us", "ul", "lbase", "us_next", "ul_next", "lbase_next")]


class Coding(C.Structure):
    """Mirror of struct LbzCoding (csrc/lbz_common.cuh)."""
    _fields_ = [("length", (C.c_uint8 * 260) * 6), ("code", (C.c_uint32 * 260) * 6),
                ("selector", C.c_uint8 * 18008), ("selector_mtf", C.c_uint8 * 18008)]


class DStreamInfo(C.Structure):
    """Mirror of lbz_dstream_info."""
    _fields_ = [(n, C.c_uint32) for n in (
        "status", "num_blocks", "num_streams", "bad_block", "garbage", "candidates",
        "false_candidates", "waves")] + [("end_bit", C.c_uint64)]


class DBlock(C.Structure):
    """Mirror of lbz_dblock."""
    _fields_ = [(n, C.c_uint64) for n in ("pos", "end_bit", "out_len", "out_off")] + [
        (n, C.c_uint32) for n in ("status", "rand", "bwt_idx", "block_size", "alpha_size", "num_trees",
                                  "num_selectors", "period", "rl_state", "crc_acc", "crc", "ntok", "nsym", "ngrp")] + [
        ("sym_bit", C.c_uint64)]


# the reference's `enum error` (src/common.h:54-76)
STATUS_NAMES = ["OK", "MORE", "FINISH", "ERR_MAGIC", "ERR_HEADER", "ERR_BITMAP", "ERR_TREES",
                "ERR_GROUPS", "ERR_SELECTOR", "ERR_DELTA", "ERR_PREFIX", "ERR_INCOMPLT",
                "ERR_EMPTY", "ERR_UNTERM", "ERR_RUNLEN", "ERR_BLKCRC", "ERR_STRMCRC",
                "ERR_OVERFLOW", "ERR_BWTIDX", "ERR_EOF"]
ERR_OUTCAP = 100
D_RESIDENT_INPUT, D_DEVICE_OUTPUT = 1, 2
DA_BLOCK, DA_BWT, DA_TEXT, DA_OUT = range(4)

ST_RLE1, ST_BWT, ST_MTF, ST_HUFFMAN, ST_PACK = range(5)
AR_TEXT, AR_BWT, AR_MTFV, AR_FREQ, AR_CODING, AR_OUT, AR_META, AR_SA = range(8)

EXPORTS = [
    # reference-shaped API (src/encode.h:29-36)
    "encoder_alloc_size", "encoder_init", "collect", "encode", "transmit", "generate_prefix_code", "divbwt",
    "crc_table", "lbz_set_fatal_handler",
    # reference-shaped decoder API (src/decode.h:72-81)
    "decoder_init", "decoder_free", "retrieve", "decode", "emit",
    # batch API
    "lbz_engine_create", "lbz_engine_destroy", "lbz_bound", "lbz_compress_chunks",
    "lbz_compress_chunks_device", "lbz_compress_chunks_h2d", "lbz_scatter_to_host", "lbz_compress_stream", "lbz_host_alloc", "lbz_host_free",
    "lbz_engine_launches", "lbz_engine_last_rounds", "lbz_engine_device_bytes", "lbz_version",
    "lbz_engine_last_ms", "lbz_engine_stage_ms", "lbz_engine_k0_stats",
    # stage hooks
    "lbz_dbg_load", "lbz_dbg_run", "lbz_dbg_read", "lbz_dbg_write", "lbz_dbg_num_slots",
    "lbz_dbg_set_chunks",
    # batch decompression (section 4 of the header)
    "lbz_decoder_create", "lbz_decoder_destroy", "lbz_decompress_stream", "lbz_decompress_ex",
    "lbz_decoder_load", "lbz_scan_blocks", "lbz_decoder_read", "lbz_decoder_last_wave_blocks",
    "lbz_decoder_launches", "lbz_decoder_device_bytes", "lbz_decoder_last_ms", "lbz_decoder_stage_ms",
    "lbz_strerror", "lbz_decoder_open", "lbz_decoder_next",
    "lbz_decoder_decode_at", "lbz_decoder_emit_at", "lbz_walk_table", "lbz_decoder_open_stream", "lbz_decoder_feed",
]


def load_library():
    """dlopen the product library.  No fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise LbzError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % path)
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, u8p, szp = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
    L.lbz_engine_create.restype = vp
    L.lbz_engine_create.argtypes = [C.c_int, C.c_int, C.c_int]
    L.lbz_engine_destroy.restype = None
    L.lbz_engine_destroy.argtypes = [vp]
    L.lbz_bound.restype = C.c_size_t
    L.lbz_bound.argtypes = [C.c_size_t]
    L.lbz_compress_chunks.restype = C.c_int
    L.lbz_compress_chunks.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp, C.POINTER(BlockRec), C.c_size_t, szp]
    L.lbz_compress_chunks_device.restype = C.c_int
    L.lbz_compress_chunks_device.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp, C.POINTER(BlockRec), C.c_size_t, szp]
    L.lbz_compress_chunks_h2d.restype = C.c_int
    L.lbz_compress_chunks_h2d.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp, C.POINTER(BlockRec), C.c_size_t, szp]
    L.lbz_scatter_to_host.restype = C.c_int
    L.lbz_scatter_to_host.argtypes = [vp, vp, C.POINTER(C.c_uint64), vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t]
    L.lbz_compress_stream.restype = C.c_int
    L.lbz_compress_stream.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp]
    L.lbz_host_alloc.restype = vp
    L.lbz_host_alloc.argtypes = [C.c_size_t]
    L.lbz_host_free.restype = None
    L.lbz_host_free.argtypes = [vp]
    L.lbz_engine_launches.restype = C.c_uint64
    L.lbz_engine_launches.argtypes = [vp]
    L.lbz_engine_last_rounds.restype = C.c_uint32
    L.lbz_engine_last_rounds.argtypes = [vp]
    L.lbz_engine_device_bytes.restype = C.c_size_t
    L.lbz_engine_device_bytes.argtypes = [vp]
    L.lbz_version.restype = C.c_char_p
    L.lbz_engine_last_ms.restype = C.c_double
    L.lbz_engine_last_ms.argtypes = [vp]
    L.lbz_engine_stage_ms.restype = None
    L.lbz_engine_stage_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.lbz_engine_k0_stats.restype = None
    L.lbz_engine_k0_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.lbz_dbg_load.restype = C.c_int
    L.lbz_dbg_load.argtypes = [vp, vp, C.c_size_t]
    L.lbz_dbg_run.restype = C.c_int
    L.lbz_dbg_run.argtypes = [vp, C.c_int]
    L.lbz_dbg_read.restype = C.c_int
    L.lbz_dbg_read.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_size_t]
    L.lbz_dbg_write.restype = C.c_int
    L.lbz_dbg_write.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_size_t]
    L.lbz_dbg_num_slots.restype = C.c_uint32
    L.lbz_dbg_num_slots.argtypes = [vp]
    L.lbz_dbg_set_chunks.restype = C.c_int
    L.lbz_dbg_set_chunks.argtypes = [vp, C.c_uint32]
    # reference-shaped API
    L.encoder_alloc_size.restype = C.c_size_t
    L.encoder_alloc_size.argtypes = [C.c_ulong]
    L.encoder_init.restype = None
    L.encoder_init.argtypes = [vp, C.c_ulong, C.c_uint]
    L.collect.restype = C.c_int
    L.collect.argtypes = [vp, vp, szp]
    L.encode.restype = C.c_size_t
    L.encode.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.transmit.restype = vp
    L.transmit.argtypes = [vp, vp]
    L.divbwt.restype = C.c_int32
    L.divbwt.argtypes = [vp, vp, vp, C.c_int32]
    # batch decompression
    L.lbz_decoder_create.restype = vp
    L.lbz_decoder_create.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t]
    L.lbz_decoder_destroy.restype = None
    L.lbz_decoder_destroy.argtypes = [vp]
    L.lbz_decompress_stream.restype = C.c_int
    L.lbz_decompress_stream.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp, C.POINTER(DStreamInfo)]
    L.lbz_decompress_ex.restype = C.c_int
    L.lbz_decompress_ex.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp, C.POINTER(DStreamInfo), C.c_uint]
    L.lbz_decoder_load.restype = C.c_int
    L.lbz_decoder_load.argtypes = [vp, vp, C.c_size_t]
    L.lbz_scan_blocks.restype = C.c_long
    L.lbz_scan_blocks.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t]
    L.lbz_decoder_read.restype = C.c_int
    L.lbz_decoder_read.argtypes = [vp, C.c_int, C.c_uint64, vp, C.c_size_t]
    L.lbz_decoder_last_wave_blocks.restype = C.c_uint32
    L.lbz_decoder_last_wave_blocks.argtypes = [vp]
    L.lbz_decoder_launches.restype = C.c_uint64
    L.lbz_decoder_launches.argtypes = [vp]
    L.lbz_decoder_device_bytes.restype = C.c_size_t
    L.lbz_decoder_device_bytes.argtypes = [vp]
    L.lbz_decoder_last_ms.restype = C.c_double
    L.lbz_decoder_last_ms.argtypes = [vp]
    L.lbz_decoder_stage_ms.restype = None
    L.lbz_decoder_stage_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.lbz_decoder_open.restype = C.c_int
    L.lbz_decoder_open.argtypes = [vp, vp, C.c_size_t, C.c_uint]
    L.lbz_decoder_open_stream.restype = C.c_int
    L.lbz_decoder_open_stream.argtypes = [vp, C.c_uint]
    L.lbz_decoder_feed.restype = C.c_int
    L.lbz_decoder_feed.argtypes = [vp, vp, C.c_size_t, C.c_int, szp]
    L.lbz_decoder_next.restype = C.c_int
    L.lbz_decoder_next.argtypes = [vp, vp, C.c_size_t, szp, C.POINTER(DStreamInfo)]
    L.lbz_decoder_decode_at.restype = C.c_int
    L.lbz_decoder_decode_at.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(DBlock), C.c_uint]
    L.lbz_decoder_emit_at.restype = C.c_int
    L.lbz_decoder_emit_at.argtypes = [vp, C.POINTER(C.c_uint64), C.c_uint32, vp, C.c_size_t, szp, C.POINTER(C.c_uint32)]
    L.lbz_walk_table.restype = C.c_int
    L.lbz_walk_table.argtypes = [vp, C.c_size_t, C.POINTER(DBlock), C.c_size_t, C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint32), szp, C.POINTER(DStreamInfo)]
    L.lbz_strerror.restype = C.c_char_p
    L.lbz_strerror.argtypes = [C.c_int]
    _LIB = L
    return L


class PinnedArray:
    """A uint8 numpy array over page-locked host memory (lbz_host_alloc): copies to and from the
    device run at the full link rate and asynchronously.  `.a` is the array; close() frees it."""

    def __init__(self, nbytes, L=None):
        self.L = L or load_library()
        self.n = max(int(nbytes), 1)
        self.ptr = self.L.lbz_host_alloc(self.n)
        if not self.ptr:
            raise LbzError("lbz_host_alloc(%d) failed" % self.n)
        self.a = np.ctypeslib.as_array((C.c_uint8 * self.n).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.a = None
            self.L.lbz_host_free(self.ptr)
            self.ptr = None


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class Engine:
    """One GPU context for one bzip2 level (mirror of lbz_engine)."""

    def __init__(self, device=0, level=9, max_chunks=64):
        self.L = load_library()
        self.level = level
        self.mbs = level * 100000
        self.max_chunks = max_chunks
        self.h = self.L.lbz_engine_create(device, level, max_chunks)
        if not self.h:
            raise LbzError("lbz_engine_create failed (no usable GPU? this package has no CPU path)")

    def close(self):
        if self.h:
            self.L.lbz_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batch API ----------------------------------------------------------
    def compress_stream(self, data):
        a = _as_u8(data)
        cap = self.L.lbz_bound(a.size) + 64
        out = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t(0)
        src = a if a.size else np.zeros(1, np.uint8)
        rc = self.L.lbz_compress_stream(self.h, src.ctypes.data, a.size, out.ctypes.data, cap, C.byref(ln))
        if rc:
            raise LbzError("lbz_compress_stream failed (%d)" % rc)
        return out[: ln.value].tobytes()

    def compress_chunks(self, data):
        a = _as_u8(data)
        cap = self.L.lbz_bound(a.size) + 64
        out = np.empty(cap, dtype=np.uint8)
        maxrec = 2 * (a.size // self.mbs + 2)
        recs = (BlockRec * maxrec)()
        ln, nr = C.c_size_t(0), C.c_size_t(0)
        src = a if a.size else np.zeros(1, np.uint8)
        rc = self.L.lbz_compress_chunks(self.h, src.ctypes.data, a.size, out.ctypes.data, cap, C.byref(ln),
                                        recs, maxrec, C.byref(nr))
        if rc:
            raise LbzError("lbz_compress_chunks failed (%d)" % rc)
        return out[: ln.value].tobytes(), [recs[i] for i in range(nr.value)]

    def scatter_to_host(self, d_src, src_off, h_dst, dst_off, lens):
        """Blocks from device memory to their places in a host buffer (see lbz_scatter_to_host)."""
        n = len(lens)
        a = np.ascontiguousarray(src_off, dtype=np.uint64)
        b = np.ascontiguousarray(dst_off, dtype=np.uint64)
        c = np.ascontiguousarray(lens, dtype=np.uint64)
        u64p = C.POINTER(C.c_uint64)
        rc = self.L.lbz_scatter_to_host(self.h, d_src, a.ctypes.data_as(u64p), h_dst, b.ctypes.data_as(u64p),
                                        c.ctypes.data_as(u64p), n)
        if rc:
            raise LbzError("lbz_scatter_to_host failed (%d)" % rc)

    def compress_chunks_ptr(self, in_ptr, n, out_ptr, out_cap, device=False, max_recs=0, h2d=False):
        """Raw-pointer form (host or device memory; h2d: host in, device out); returns (out_len, recs)."""
        maxrec = max_recs or 2 * (n // self.mbs + 2)
        recs = (BlockRec * maxrec)()
        ln, nr = C.c_size_t(0), C.c_size_t(0)
        fn = self.L.lbz_compress_chunks_h2d if h2d else (self.L.lbz_compress_chunks_device if device else self.L.lbz_compress_chunks)
        rc = fn(self.h, in_ptr, n, out_ptr, out_cap, C.byref(ln), recs, maxrec, C.byref(nr))
        if rc:
            raise LbzError("compress failed (%d)" % rc)
        return ln.value, [recs[i] for i in range(nr.value)]

    @property
    def launches(self):
        return self.L.lbz_engine_launches(self.h)

    @property
    def last_ms(self):
        return self.L.lbz_engine_last_ms(self.h)

    def stage_ms(self):
        a = (C.c_double * 7)()
        self.L.lbz_engine_stage_ms(self.h, a)
        return dict(zip(("rle1", "sort_initial", "sort_refine", "bwt_gather", "mtf", "huffman", "pack"), list(a)))

    def k0_stats(self):
        ms, nl, ne = C.c_double(0), C.c_uint32(0), C.c_uint64(0)
        self.L.lbz_engine_k0_stats(self.h, C.byref(ms), C.byref(nl), C.byref(ne))
        return ms.value, nl.value, ne.value

    @property
    def last_rounds(self):
        return self.L.lbz_engine_last_rounds(self.h)

    @property
    def device_bytes(self):
        return self.L.lbz_engine_device_bytes(self.h)

    # ---- stage hooks (tests) ------------------------------------------------
    def dbg_load(self, data):
        a = _as_u8(data)
        src = a if a.size else np.zeros(1, np.uint8)
        if self.L.lbz_dbg_load(self.h, src.ctypes.data, a.size):
            raise LbzError("dbg_load failed")

    def dbg_set_chunks(self, k):
        if self.L.lbz_dbg_set_chunks(self.h, k):
            raise LbzError("dbg_set_chunks failed")

    def dbg_run(self, stage):
        if self.L.lbz_dbg_run(self.h, stage):
            raise LbzError("stage %d failed" % stage)

    def dbg_read(self, array, slot, dtype, count):
        out = np.empty(count, dtype=dtype)
        if count and self.L.lbz_dbg_read(self.h, array, slot, out.ctypes.data, out.nbytes):
            raise LbzError("dbg_read failed")
        return out

    def dbg_read_struct(self, array, slot, typ):
        v = typ()
        if self.L.lbz_dbg_read(self.h, array, slot, C.addressof(v), C.sizeof(v)):
            raise LbzError("dbg_read failed")
        return v

    def dbg_write(self, array, slot, arr):
        a = np.ascontiguousarray(arr)
        if a.nbytes and self.L.lbz_dbg_write(self.h, array, slot, a.ctypes.data, a.nbytes):
            raise LbzError("dbg_write failed")

    def dbg_write_struct(self, array, slot, v):
        if self.L.lbz_dbg_write(self.h, array, slot, C.addressof(v), C.sizeof(v)):
            raise LbzError("dbg_write failed")

    def meta(self, slot):
        return self.dbg_read_struct(AR_META, slot, BlockMeta)


NOEMIT = 0xFFFFFFFFFFFFFFFF


def dblock_copy(b):
    c = DBlock()
    C.memmove(C.byref(c), C.byref(b), C.sizeof(DBlock))
    return c


def walk_table(fn, z, table):
    """Framing walk over decoded candidates sorted by position (lbz_walk_table):
    (status, chain indices, expected CRCs, DStreamInfo)."""
    a = _as_u8(z)
    k = len(table)
    arr = (DBlock * max(k, 1))(*table)
    chain = (C.c_uint32 * max(k, 1))()
    ccrc = (C.c_uint32 * max(k, 1))()
    n = C.c_size_t(0)
    info = DStreamInfo()
    src = a if a.size else np.zeros(1, np.uint8)
    st = fn(src.ctypes.data, a.size, arr, k, chain, ccrc, C.byref(n), C.byref(info))
    return st, list(chain[: n.value]), list(ccrc[: n.value]), info


class Decoder:
    """One GPU context for batch decompression (mirror of lbz_decoder)."""

    def __init__(self, device=0, max_blocks=64, in_cap=1 << 24, out_cap=0):
        self.L = load_library()
        self.max_blocks = max_blocks
        self.h = self.L.lbz_decoder_create(device, max_blocks, in_cap, out_cap)
        if not self.h:
            raise LbzError("lbz_decoder_create failed (no usable GPU? this package has no CPU path)")

    def close(self):
        if self.h:
            self.L.lbz_decoder_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decompress(self, z, cap=None):
        """(status, output bytes, DStreamInfo).  status: 0, the reference's error kind, or 100.
        Without `cap` the output grows as needed: the stream is decoded wave by wave
        (lbz_decoder_open / lbz_decoder_next) into a wave buffer that holds at least the largest
        possible block (a block of 900 000 bytes of RLE1 output can expand to 46.6 MB), so inputs
        that expand far beyond any fixed ratio (long runs) decode like everything else."""
        a = _as_u8(z)
        src = a if a.size else np.zeros(1, np.uint8)
        if cap is None:
            wave_cap = max(64 << 20, 4 * a.size)
            st = self.L.lbz_decoder_open(self.h, src.ctypes.data, a.size, 0)
            if st < 0:
                raise LbzError("lbz_decoder_open failed (%d)" % st)
            info = DStreamInfo()
            if st != 0:
                info.status = st
                return st, b"", info
            buf = np.empty(wave_cap, dtype=np.uint8)
            parts = []
            while True:
                n = C.c_size_t(0)
                st = self.L.lbz_decoder_next(self.h, buf.ctypes.data, wave_cap, C.byref(n), C.byref(info))
                if st < 0:
                    raise LbzError("lbz_decoder_next failed (%d)" % st)
                parts.append(buf[: n.value].tobytes())
                if st != 1:                      # LBZ_MORE: another wave follows
                    return st, b"".join(parts), info
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        info = DStreamInfo()
        st = self.L.lbz_decompress_stream(self.h, src.ctypes.data, a.size, out.ctypes.data, cap, C.byref(n),
                                          C.byref(info))
        if st < 0:
            raise LbzError("lbz_decompress_stream failed (%d)" % st)
        return st, out[: n.value].tobytes(), info

    def decompress_waves(self, z, wave_cap):
        """Generator over (status, bytes) of lbz_decoder_open / lbz_decoder_next."""
        a = _as_u8(z)
        src = a if a.size else np.zeros(1, np.uint8)
        st = self.L.lbz_decoder_open(self.h, src.ctypes.data, a.size, 0)
        if st < 0:
            raise LbzError("lbz_decoder_open failed")
        if st != 0:
            yield st, b""
            return
        buf = np.empty(max(wave_cap, 1), dtype=np.uint8)
        while True:
            n = C.c_size_t(0)
            info = DStreamInfo()
            st = self.L.lbz_decoder_next(self.h, buf.ctypes.data, wave_cap, C.byref(n), C.byref(info))
            if st < 0:
                raise LbzError("lbz_decoder_next failed")
            yield st, buf[: n.value].tobytes()
            if st != 1:
                return

    def decompress_pieces(self, pieces, wave_cap):
        """Streaming session (lbz_decoder_open_stream / feed / next): `pieces` is an iterable of
        bytes-like parts of one file, of any sizes; only a window of in_cap compressed bytes is
        resident.  Returns (status, output bytes, info)."""
        if self.L.lbz_decoder_open_stream(self.h, 0) != 0:
            raise LbzError("lbz_decoder_open_stream failed")
        buf = np.empty(max(wave_cap, 1), dtype=np.uint8)
        out = []
        info = DStreamInfo()
        it = iter(pieces)
        cur, off, eof = np.zeros(0, np.uint8), 0, False

        def feed():
            nonlocal cur, off, eof
            progressed = False
            while not eof:
                if off >= cur.size:
                    nxt = next(it, None)
                    if nxt is None:
                        if self.L.lbz_decoder_feed(self.h, buf.ctypes.data, 0, 1, None) != 0:
                            raise LbzError("lbz_decoder_feed failed")
                        eof = True
                        return True
                    cur, off = _as_u8(nxt), 0
                    continue
                took = C.c_size_t(0)
                if self.L.lbz_decoder_feed(self.h, cur[off:].ctypes.data, cur.size - off, 0, C.byref(took)) != 0:
                    raise LbzError("lbz_decoder_feed failed")
                off += took.value
                progressed = progressed or took.value > 0
                if off < cur.size:          # window full
                    return progressed
            return progressed

        feed()
        while True:
            n = C.c_size_t(0)
            st = self.L.lbz_decoder_next(self.h, buf.ctypes.data, wave_cap, C.byref(n), C.byref(info))
            if st < 0:
                raise LbzError("lbz_decoder_next failed")
            if n.value:
                out.append(buf[: n.value].tobytes())
            if st == 1:                     # LBZ_MORE: blocks remain in the window; top it up meanwhile
                feed()
                continue
            if st == 101:                   # LBZ_NEED_INPUT
                if not feed():
                    raise LbzError("the decoder wants input but its window is full")
                continue
            return st, b"".join(out), info

    # ---- sharding building blocks (see lbzip2_b200/sharding.py sharded_decompress) ----
    keeps_scanned_input = True        # scan() leaves the stream on the device: decode_at(..., resident=True)

    def decode_at(self, z, positions, resident=False):
        """Decode the candidate blocks whose magics start at the given bit positions; list of DBlock.
        resident=True: `z` is the input the decoder already holds (scan() / load() of the same bytes),
        no second upload."""
        a = _as_u8(z)
        k = len(positions)
        pos = (C.c_uint64 * max(k, 1))(*positions)
        table = (DBlock * max(k, 1))()
        src = a if a.size else np.zeros(1, np.uint8)
        if self.L.lbz_decoder_decode_at(self.h, src.ctypes.data, a.size, pos, k, table, 1 if resident else 0):
            raise LbzError("lbz_decoder_decode_at failed")
        return [dblock_copy(table[i]) for i in range(k)]

    def emit_at(self, out_offs, cap, out=None):
        """Write the blocks of the last decode_at at the given offsets (NOEMIT = skip): (bytes, crcs).
        `out` (a uint8 array of >= cap bytes, e.g. pinned_u8()): the decoded bytes land there and a
        view of it is returned instead of a copy."""
        k = len(out_offs)
        offs = (C.c_uint64 * max(k, 1))(*out_offs)
        crc = (C.c_uint32 * max(k, 1))()
        if out is not None and out.size < cap:
            raise LbzError("emit_at: the output array holds %d bytes, %d needed" % (out.size, cap))
        buf = out if out is not None else np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        if self.L.lbz_decoder_emit_at(self.h, offs, k, buf.ctypes.data, cap, C.byref(n), crc):
            raise LbzError("lbz_decoder_emit_at failed")
        return (buf[: n.value] if out is not None else buf[: n.value].tobytes()), list(crc[:k])

    def walk_table(self, z, table):
        return walk_table(self.L.lbz_walk_table, z, table)

    def decompress_ptr(self, in_ptr, n, out_ptr, out_cap, flags=0):
        """Raw-pointer form; returns (status, out_len, info)."""
        ln = C.c_size_t(0)
        info = DStreamInfo()
        st = self.L.lbz_decompress_ex(self.h, in_ptr, n, out_ptr, out_cap, C.byref(ln), C.byref(info), flags)
        if st < 0:
            raise LbzError("lbz_decompress_ex failed (%d)" % st)
        return st, ln.value, info

    def load(self, in_ptr, n):
        if self.L.lbz_decoder_load(self.h, in_ptr, n):
            raise LbzError("lbz_decoder_load failed")

    def scan(self, z):
        a = _as_u8(z)
        cap = a.size // 6 + 64
        pos = np.empty(cap, dtype=np.uint64)           # untouched pages cost nothing: only the hits are written
        src = a if a.size else np.zeros(1, np.uint8)
        k = self.L.lbz_scan_blocks(self.h, src.ctypes.data, a.size, pos.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
        if k < 0:
            raise LbzError("lbz_scan_blocks failed")
        return pos[:k].tolist()

    def block(self, slot):
        b = DBlock()
        if self.L.lbz_decoder_read(self.h, DA_BLOCK, slot, C.byref(b), C.sizeof(b)):
            raise LbzError("lbz_decoder_read failed")
        return b

    def array(self, which, slot, nbytes):
        a = np.empty(max(nbytes, 1), np.uint8)
        if self.L.lbz_decoder_read(self.h, which, slot, a.ctypes.data, nbytes):
            raise LbzError("lbz_decoder_read failed")
        return a[:nbytes]

    @property
    def last_wave_blocks(self):
        return self.L.lbz_decoder_last_wave_blocks(self.h)

    @property
    def launches(self):
        return self.L.lbz_decoder_launches(self.h)

    @property
    def last_ms(self):
        return self.L.lbz_decoder_last_ms(self.h)

    def stage_ms(self):
        v = (C.c_double * 7)()
        self.L.lbz_decoder_stage_ms(self.h, v)
        return dict(zip(("upload", "scan", "retrieve", "successors", "walks", "expand", "tail"), list(v)))

    @property
    def device_bytes(self):
        return self.L.lbz_decoder_device_bytes(self.h)
