"""ctypes bindings for the TEST-ONLY libraries under oracle/.

* ``oracle/liboracle.so``        -- our CPU restatement (always available; built on demand)
* ``oracle/_ref/libref_stages.so`` -- harness around the unmodified reference
  (only where ``oracle/_ref`` was built, i.e. the dev container, or where the
  prebuilt files travelled with the snapshot).

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


class OrcBlockInfo(C.Structure):
    _fields_ = [("consumed", C.c_uint64)] + [
        (n, C.c_uint32) for n in (
            "nblock", "block_crc", "bwt_idx", "tie_count", "nmtf", "alpha_size",
            "num_trees", "num_selectors", "tree_pad", "out_len")]


class OrcCoding(C.Structure):
    _fields_ = [
        ("num_trees", C.c_uint32), ("num_groups", C.c_uint32),
        ("num_selectors", C.c_uint32), ("tree_pad", C.c_uint32),
        ("out_len", C.c_uint32),
        ("length", (C.c_uint8 * 259) * 6),
        ("code", (C.c_uint32 * 259) * 6),
        ("selector", C.c_uint8 * 18002),
        ("selector_mtf", C.c_uint8 * 18010),
    ]


class RefDump(C.Structure):
    _fields_ = [("consumed", C.c_uint64)] + [
        (n, C.c_uint32) for n in (
            "full", "nblock", "block_crc", "bwt_idx", "nmtf", "alpha_size",
            "num_selectors", "num_trees", "tree_pad", "out_len")] + [
        ("used", C.c_uint8 * 256),
        ("tmap_new2old", C.c_uint8 * 6),
        ("length", (C.c_uint8 * 259) * 6),
        ("code", (C.c_uint32 * 259) * 6),
    ]


class OrcDBlock(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "status", "rand", "bwt_idx", "block_size", "alpha_size", "num_trees",
        "num_selectors", "pad")] + [("end_bit", C.c_uint64)]


class OrcDStream(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "status", "num_blocks", "num_streams", "bad_block", "garbage", "pad")] + [
        ("end_bit", C.c_uint64)]


# the reference's `enum error` (common.h:54-76) and err2str() texts (expand.c:70-88)
ERR_NAMES = ["OK", "MORE", "FINISH", "ERR_MAGIC", "ERR_HEADER", "ERR_BITMAP", "ERR_TREES",
             "ERR_GROUPS", "ERR_SELECTOR", "ERR_DELTA", "ERR_PREFIX", "ERR_INCOMPLT",
             "ERR_EMPTY", "ERR_UNTERM", "ERR_RUNLEN", "ERR_BLKCRC", "ERR_STRMCRC",
             "ERR_OVERFLOW", "ERR_BWTIDX", "ERR_EOF"]
ERR_TEXT = {3: "not a valid bzip2 file", 4: "bad block header magic", 5: "empty source alphabet",
            6: "bad number of trees", 7: "no coding groups", 8: "invalid selector",
            9: "invalid delta code", 10: "invalid prefix code", 11: "incomplete prefix code",
            12: "empty block", 13: "unterminated block", 14: "missing run length",
            15: "block CRC mismatch", 16: "stream CRC mismatch", 17: "block overflow",
            18: "primary index too large", 19: "unexpected end of file"}

_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "bz_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"],
                                  stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.orc_crc_update.restype = C.c_uint32
        L.orc_crc_update.argtypes = [C.c_uint32, u8p, C.c_size_t]
        L.orc_rle1.restype = C.c_int
        L.orc_rle1.argtypes = [u8p, C.c_size_t, C.c_uint32, u8p, u32p,
                               C.POINTER(C.c_size_t), u8p, u32p]
        L.orc_bwt.restype = C.c_uint32
        L.orc_bwt.argtypes = [u8p, C.c_uint32, u8p, u32p]
        L.orc_mtf.restype = C.c_uint32
        L.orc_mtf.argtypes = [u8p, C.c_uint32, u8p, u16p, u32p, u32p]
        L.orc_prefix_code.restype = None
        L.orc_prefix_code.argtypes = [u16p, C.c_uint32, C.c_uint32, u32p, u8p,
                                      C.c_uint, C.POINTER(OrcCoding)]
        L.orc_pack.restype = C.c_size_t
        L.orc_pack.argtypes = [C.POINTER(OrcCoding), u16p, C.c_uint32, C.c_uint32,
                               u8p, C.c_uint32, C.c_uint32, u8p]
        L.orc_encode_block.restype = C.c_size_t
        L.orc_encode_block.argtypes = [u8p, C.c_size_t, C.c_uint32, u8p,
                                       C.POINTER(OrcBlockInfo)]
        L.orc_stream_bound.restype = C.c_size_t
        L.orc_stream_bound.argtypes = [C.c_size_t]
        L.orc_compress_stream.restype = C.c_size_t
        L.orc_compress_stream.argtypes = [u8p, C.c_size_t, C.c_int, u8p,
                                          C.POINTER(OrcBlockInfo), C.c_size_t,
                                          C.POINTER(C.c_size_t)]
        L.orc_d_retrieve.restype = C.c_int
        L.orc_d_retrieve.argtypes = [u8p, C.c_size_t, C.c_uint64, u8p, C.POINTER(OrcDBlock)]
        L.orc_d_ibwt.restype = None
        L.orc_d_ibwt.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_int, u8p]
        L.orc_d_unrle.restype = C.c_int
        L.orc_d_unrle.argtypes = [u8p, C.c_uint32, u8p, C.c_size_t, C.POINTER(C.c_size_t), u32p]
        L.orc_decompress_stream.restype = C.c_int
        L.orc_decompress_stream.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.POINTER(OrcDStream)]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_stages.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_stages.so"))
        L.ref_encode_block.restype = C.c_int
        L.ref_encode_block.argtypes = [u8p, C.c_uint64, C.c_uint32, C.POINTER(RefDump),
                                       u8p, u8p, u16p, u8p, u8p, u8p]
        L.ref_divbwt.restype = C.c_int32
        L.ref_divbwt.argtypes = [u8p, C.c_int32, u8p]
        L.ref_crc_table_entry.restype = C.c_uint32
        L.ref_crc_table_entry.argtypes = [C.c_uint]
        _ref = L
    return _ref


def as_u8(data):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


# ----------------------------------------------------------------- oracle ---

def orc_stream(data, level=9):
    """Whole .bz2 stream + per-block infos from the oracle restatement."""
    L = oracle()
    a = as_u8(data)
    n = a.size
    out = np.empty(L.orc_stream_bound(n), dtype=np.uint8)
    maxb = 2 * (n // (level * 100000) + 2)
    infos = (OrcBlockInfo * maxb)()
    nb = C.c_size_t(0)
    src = a if n else np.zeros(1, np.uint8)
    ln = L.orc_compress_stream(_ptr(src, u8p), n, level, _ptr(out, u8p), infos, maxb, C.byref(nb))
    return out[:ln].tobytes(), [infos[i] for i in range(nb.value)]


def orc_block_stages(data, cap):
    """All intermediate results of ONE block from the oracle (dict of numpy arrays)."""
    L = oracle()
    a = as_u8(data)
    block = np.zeros(cap + 8, np.uint8)
    used = np.zeros(256, np.uint8)
    nblock = C.c_uint32(0)
    consumed = C.c_size_t(0)
    crc = C.c_uint32(0)
    full = L.orc_rle1(_ptr(a, u8p), a.size, cap, _ptr(block, u8p), C.byref(nblock),
                      C.byref(consumed), _ptr(used, u8p), C.byref(crc))
    nb = nblock.value
    res = dict(full=full, consumed=consumed.value, nblock=nb, crc=crc.value,
               block=block[:nb].copy(), used=used.copy())
    if nb == 0:
        return res
    bwt = np.zeros(nb, np.uint8)
    tie = C.c_uint32(0)
    idx = L.orc_bwt(_ptr(block, u8p), nb, _ptr(bwt, u8p), C.byref(tie))
    mtfv = np.zeros(nb + 64, np.uint16)
    freq = np.zeros(259, np.uint32)
    asz = C.c_uint32(0)
    nm = L.orc_mtf(_ptr(bwt, u8p), nb, _ptr(used, u8p), _ptr(mtfv, u16p), _ptr(freq, u32p), C.byref(asz))
    cd = OrcCoding()
    L.orc_prefix_code(_ptr(mtfv, u16p), nm, asz.value, _ptr(freq, u32p), _ptr(used, u8p), 8, C.byref(cd))
    out = np.zeros(cd.out_len + 8, np.uint8)
    ln = L.orc_pack(C.byref(cd), _ptr(mtfv, u16p), nm, asz.value, _ptr(used, u8p), crc.value, idx, _ptr(out, u8p))
    res.update(bwt=bwt, bwt_idx=idx, tie_count=tie.value, mtfv=mtfv[:nm].copy(), nmtf=nm,
               freq=freq, alpha_size=asz.value, coding=cd, bits=out[:ln].copy(), out_len=cd.out_len)
    return res


def orc_decompress(z, cap=None):
    """Whole-file decompression by the oracle: (status, output bytes, OrcDStream)."""
    L = oracle()
    a = as_u8(z)
    if cap is None:
        cap = max(1 << 20, 64 * a.size)
    out = np.empty(cap, np.uint8)
    n = C.c_size_t(0)
    si = OrcDStream()
    src = a if a.size else np.zeros(1, np.uint8)
    st = L.orc_decompress_stream(_ptr(src, u8p), a.size, _ptr(out, u8p), cap, C.byref(n), C.byref(si))
    return st, out[: n.value].tobytes(), si


def orc_retrieve(z, bitpos):
    """One block's payload at absolute bit `bitpos`: (OrcDBlock, bwt bytes)."""
    L = oracle()
    a = as_u8(z)
    bwt = np.zeros(900000, np.uint8)
    bi = OrcDBlock()
    L.orc_d_retrieve(_ptr(a, u8p), a.size, bitpos, _ptr(bwt, u8p), C.byref(bi))
    return bi, bwt[: bi.block_size].copy()


def orc_ibwt(bwt, idx, rand=0):
    L = oracle()
    a = as_u8(bwt)
    out = np.zeros(max(a.size, 1), np.uint8)
    L.orc_d_ibwt(_ptr(a, u8p), a.size, idx, rand, _ptr(out, u8p))
    return out[: a.size]


def orc_unrle(src, cap=None):
    L = oracle()
    a = as_u8(src)
    if cap is None:
        cap = 64 * a.size + 1024
    out = np.zeros(cap, np.uint8)
    n = C.c_size_t(0)
    crc = C.c_uint32(0)
    s = a if a.size else np.zeros(1, np.uint8)
    st = L.orc_d_unrle(_ptr(s, u8p), a.size, _ptr(out, u8p), cap, C.byref(n), C.byref(crc))
    return st, out[: n.value].copy(), crc.value


def ref_cli_decompress(z):
    """(exit status, stdout, stderr) of the compiled reference `lbzip2 -d -n1`."""
    p = subprocess.run([os.path.join(REF_DIR, "lbzip2"), "-d", "-c", "-n1"], input=bytes(z),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr.decode("latin1")


# -------------------------------------------------------------- reference ---

def ref_block_stages(data, cap):
    """All intermediate results of ONE block from the compiled reference."""
    L = ref()
    a = as_u8(data)
    d = RefDump()
    block = np.zeros(cap + 8, np.uint8)
    bwt = np.zeros(cap + 8, np.uint8)
    mtfv = np.zeros(cap + 64, np.uint16)
    sel = np.zeros(18002, np.uint8)
    selmtf = np.zeros(18010, np.uint8)
    bits = np.zeros(cap + cap // 4 + 4096, np.uint8)
    src = a if a.size else np.zeros(1, np.uint8)
    rc = L.ref_encode_block(_ptr(src, u8p), a.size, cap, C.byref(d), _ptr(block, u8p), _ptr(bwt, u8p),
                            _ptr(mtfv, u16p), _ptr(sel, u8p), _ptr(selmtf, u8p), _ptr(bits, u8p))
    if rc != 0:
        return dict(nblock=0, consumed=0)
    length = np.ctypeslib.as_array(d.length).copy()
    code = np.ctypeslib.as_array(d.code).copy()
    n2o = list(d.tmap_new2old)[: d.num_trees]
    ng = (d.nmtf + 49) // 50
    return dict(full=d.full, consumed=d.consumed, nblock=d.nblock, crc=d.block_crc,
                block=block[: d.nblock].copy(), used=np.array(list(d.used), np.uint8),
                bwt=bwt[: d.nblock].copy(), bwt_idx=d.bwt_idx, mtfv=mtfv[: d.nmtf].copy(),
                nmtf=d.nmtf, alpha_size=d.alpha_size, num_trees=d.num_trees,
                num_selectors=d.num_selectors, tree_pad=d.tree_pad, out_len=d.out_len,
                new2old=n2o, length_old=length, code_old=code,
                selector_old=sel[:ng].copy(), selector_mtf=selmtf[: d.num_selectors].copy(),
                bits=bits[: d.out_len].copy())


def ref_divbwt(block):
    L = ref()
    a = as_u8(block)
    out = np.zeros(a.size, np.uint8)
    idx = L.ref_divbwt(_ptr(a, u8p), a.size, _ptr(out, u8p))
    return out, idx


def ref_cli(data, level=9, threads=None):
    """Output of the unmodified reference CLI (oracle/_ref/lbzip2)."""
    args = [os.path.join(REF_DIR, "lbzip2"), "-%d" % level]
    if threads:
        args.append("-n%d" % threads)
    return subprocess.run(args, input=bytes(data), stdout=subprocess.PIPE, check=True).stdout
