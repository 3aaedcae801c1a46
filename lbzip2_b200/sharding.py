"""Multi-GPU sharding of the block-compression path (host logic only).

Blocks are independent: raw chunk i (level*100000 bytes,
reference src/process.c:631) goes to rank i mod world; no exchange happens
during compute; afterwards the byte-aligned block bitstreams are gathered to
rank 0 in stream order (reference order key: (major = chunk, minor = block in
chunk), src/compress.c:85-86,99-100) and rank 0 adds the stream header,
trailer and combined CRC (src/compress.c:290-321, src/encode.h:38).

Works with any torch.distributed backend: NCCL over NVLink on GPUs, gloo on
CPU tensors (used by the world_size-2 CPU test)."""
import numpy as np


def rank_chunk_ids(n_bytes, mbs, world, rank):
    nchunks = (n_bytes + mbs - 1) // mbs
    return list(range(rank, nchunks, world))


def rank_input(data, mbs, world, rank):
    """Concatenation of this rank's chunks (a view-free copy; the short last
    chunk of the stream, if this rank owns it, is last here as well)."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    ids = rank_chunk_ids(a.size, mbs, world, rank)
    if not ids:
        return np.zeros(0, np.uint8)
    return np.concatenate([a[i * mbs:(i + 1) * mbs] for i in ids])


def block_table(recs, mbs):
    """(local chunk index, out_len, crc) per block, in local stream order."""
    t = np.zeros((len(recs), 3), dtype=np.int64)
    for k, r in enumerate(recs):
        t[k] = (r.raw_offset // mbs, r.out_len, r.crc)
    return t


def assemble_stream(level, tables, payloads, world):
    """Rank-0 side: interleave the per-rank block payloads into stream order.
    tables[r]: int64 [nblocks_r, 3]; payloads[r]: uint8 array of rank r's blocks
    concatenated in its local order."""
    offs = []
    for r in range(world):
        o = np.concatenate(([0], np.cumsum(tables[r][:, 1]))) if len(tables[r]) else np.zeros(1, np.int64)
        offs.append(o)
    cursor = [0] * world
    parts = [b"BZh" + bytes([ord("0") + level])]
    cc = 0
    nchunks_total = sum((int(t[:, 0].max()) + 1) if len(t) else 0 for t in tables)
    for i in range(nchunks_total):
        r = i % world
        local = i // world
        t = tables[r]
        while cursor[r] < len(t) and t[cursor[r], 0] == local:
            k = cursor[r]
            parts.append(payloads[r][offs[r][k]:offs[r][k + 1]].tobytes())
            crc = int(t[k, 2]) & 0xFFFFFFFF
            cc = (((cc << 1) & 0xFFFFFFFF) ^ (cc >> 31) ^ crc ^ 0xFFFFFFFF) & 0xFFFFFFFF
            cursor[r] += 1
    parts.append(bytes([0x17, 0x72, 0x45, 0x38, 0x50, 0x90]) + cc.to_bytes(4, "big"))
    return b"".join(parts)


def to_host(tables, payloads):
    """Device tensors of gather_blocks(..., to_host=False) -> numpy arrays."""
    return [t.cpu().numpy() for t in tables], [p.cpu().numpy() for p in payloads]


def gather_blocks(table, payload, dist, device, to_host=True):
    """Gather (table, payload) of every rank to rank 0 with torch.distributed.
    `payload` is a uint8 torch tensor on `device` (stays on the GPU for NCCL).
    Returns (tables, payloads) on rank 0 -- numpy arrays, or tensors that stay on
    `device` when to_host is False -- else (None, None)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = torch.tensor([table.shape[0], payload.numel()], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().tolist() for s in all_sizes]
    max_rows = max(s[0] for s in all_sizes)
    max_bytes = max(s[1] for s in all_sizes)
    tpad = torch.zeros((max(max_rows, 1), 3), dtype=torch.int64, device=device)
    if table.shape[0]:
        tpad[: table.shape[0]] = torch.from_numpy(table).to(device)
    ppad = torch.zeros(max(max_bytes, 1), dtype=torch.uint8, device=device)
    ppad[: payload.numel()] = payload
    if rank == 0:
        tl = [torch.zeros_like(tpad) for _ in range(world)]
        pl = [torch.zeros_like(ppad) for _ in range(world)]
        dist.gather(tpad, tl, dst=0)
        dist.gather(ppad, pl, dst=0)          # NCCL over NVLink when tensors are on GPUs
        tables = [tl[r][: all_sizes[r][0]] for r in range(world)]
        payloads = [pl[r][: all_sizes[r][1]] for r in range(world)]
        if to_host:
            return [t.cpu().numpy() for t in tables], [p.cpu().numpy() for p in payloads]
        return tables, payloads
    dist.gather(tpad, None, dst=0)
    dist.gather(ppad, None, dst=0)
    return None, None


class BlockGatherer:
    """Reusable gather of (block table, payload) to rank 0 with preallocated buffers.

    Step 1: one fixed-shape gather of a small header+table tensor (int64
    [1 + max_blocks, 3]; row 0 = (nblocks, payload bytes, 0)) -- the only
    host-visible synchronisation (rank 0 reads it to learn the sizes).
    Step 2: every rank sends exactly its payload bytes to rank 0 with one
    point-to-point transfer (NCCL send/recv over NVLink; batched so that all
    transfers are in flight together); rank 0 receives each into a fixed region
    of one preallocated buffer.  Nothing is padded and nothing is allocated per
    call.  With `tables_only` the payload step is skipped (host-sink mode: every
    rank keeps its blocks in its own host memory and writes them at the offsets
    the table gives, the way a multi-process writer would pwrite them).
    """

    def __init__(self, dist, device, max_blocks, max_payload):
        import torch
        self.dist, self.device = dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.max_blocks, self.max_payload = int(max_blocks), int(max_payload)
        self.tbuf = torch.zeros((1 + self.max_blocks, 3), dtype=torch.int64, device=device)
        self.tstage = torch.zeros((1 + self.max_blocks, 3), dtype=torch.int64).pin_memory() if str(device) != "cpu" \
            else torch.zeros((1 + self.max_blocks, 3), dtype=torch.int64)
        if self.rank == 0:
            self.tall = [torch.zeros_like(self.tbuf) for _ in range(self.world)]
            self.pall = torch.empty((self.world, self.max_payload), dtype=torch.uint8, device=device)
        else:
            self.tall, self.pall = None, None

    def gather(self, table, payload, tables_only=False):
        """table: int64 numpy [nblocks, 3]; payload: uint8 tensor on `device`.
        Rank 0 returns (tables, payloads) as tensors on `device` (payloads are views of
        the preallocated buffer; None entries with tables_only); other ranks (None, None)."""
        import torch
        dist = self.dist
        nb, nbytes = int(table.shape[0]), int(payload.numel())
        if nb > self.max_blocks or nbytes > self.max_payload:
            raise ValueError("BlockGatherer capacity exceeded")
        self.tstage[0, 0], self.tstage[0, 1], self.tstage[0, 2] = nb, nbytes, 0
        if nb:
            self.tstage[1:1 + nb] = torch.from_numpy(table)
        self.tbuf.copy_(self.tstage, non_blocking=True)
        dist.gather(self.tbuf, self.tall, dst=0)
        if self.rank != 0:
            if not tables_only and nbytes:
                # batched on both sides: rank 0 posts batched receives, and batched and
                # unbatched point-to-point calls do not share a communicator in torch's NCCL group
                for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload, 0)]):
                    w.wait()
            return None, None
        heads = torch.stack([t[0] for t in self.tall]).cpu().tolist()      # the one host sync
        tables = [self.tall[r][1:1 + heads[r][0]] for r in range(self.world)]
        if tables_only:
            return tables, [None] * self.world
        payloads = [None] * self.world
        ops = []
        for r in range(self.world):
            n = heads[r][1]
            payloads[r] = self.pall[r, :n]
            if r == 0:
                payloads[0].copy_(payload, non_blocking=True)
            elif n:
                ops.append(dist.P2POp(dist.irecv, payloads[r], r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return tables, payloads


def place_blocks(tables, world, header=4):
    """Stream offsets of every rank's blocks in the one output stream.

    tables[r]: int64 [nblocks_r, 3] rows (local chunk, out_len, crc) in rank r's local order.
    Stream order is the reference's (src/compress.c:85-86,238-252): major = chunk of the job
    (local chunk * world + rank), minor = block within the chunk.  Returns (offs, total, cc):
    offs[r] = int64 array of byte offsets of rank r's blocks (the 4-byte stream header comes
    first), total = stream length without the 10-byte trailer, cc = combined CRC (src/encode.h:38)."""
    keys, lens, crcs, owner = [], [], [], []
    for r in range(world):
        t = np.asarray(tables[r], dtype=np.int64).reshape(-1, 3)
        if not len(t):
            continue
        chunk = t[:, 0] * world + r
        # minor: index of the block inside its chunk (blocks of a chunk are consecutive rows)
        first = np.concatenate(([True], chunk[1:] != chunk[:-1]))
        start = np.maximum.accumulate(np.where(first, np.arange(len(t)), 0))
        minor = np.arange(len(t)) - start
        keys.append(chunk * 4 + minor)
        lens.append(t[:, 1])
        crcs.append(t[:, 2])
        owner.append(np.full(len(t), r, dtype=np.int64))
    if not keys:
        return [np.zeros(0, np.int64) for _ in range(world)], header, 0
    keys, lens, crcs, owner = (np.concatenate(x) for x in (keys, lens, crcs, owner))
    order = np.argsort(keys, kind="stable")
    off_sorted = header + np.concatenate(([0], np.cumsum(lens[order])[:-1]))
    off = np.empty_like(off_sorted)
    off[order] = off_sorted
    offs = [off[owner == r] for r in range(world)]
    # combined CRC (cc' = rotl(cc, 1) ^ ~crc, src/encode.h:38) without a Python loop: the fold is linear
    # over GF(2), block k of n contributes ~crc_k rotated left by (n - 1 - k) mod 32
    x = (~crcs[order]) & 0xFFFFFFFF
    r = (len(x) - 1 - np.arange(len(x))) % 32
    rot = ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF
    cc = int(np.bitwise_xor.reduce(rot)) if len(x) else 0
    return offs, int(header + lens.sum()), cc


class SharedStream:
    """The ONE stream-ordered .bz2 of a sharded job in host memory: a /dev/shm file that every rank
    of the node maps, registered with CUDA where possible so that device->host copies land in it
    directly.  Per step: exchange() the block tables (one small all_gather), place_blocks(), every
    rank scatters its blocks to their offsets (lbz_scatter_to_host), rank 0 adds header and trailer
    (src/compress.c:290-321); after the closing barrier the file holds the complete stream."""

    def __init__(self, dist, device, path, capacity, max_blocks, register=True):
        import mmap
        import os
        import torch
        self.dist, self.device = dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.path, self.capacity, self.max_blocks = path, int(capacity), int(max_blocks)
        if self.rank == 0:
            with open(path, "wb") as f:
                f.truncate(self.capacity)
        dist.barrier()
        self.fd = os.open(path, os.O_RDWR)
        self.mm = mmap.mmap(self.fd, self.capacity)
        import ctypes as C
        self.ptr = C.addressof(C.c_char.from_buffer(self.mm))
        self.view = np.frombuffer(self.mm, dtype=np.uint8)
        self.registered = False
        if register and str(device) != "cpu":
            try:
                rc = torch.cuda.cudart().cudaHostRegister(self.ptr, self.capacity, 0)
                self.registered = int(rc) == 0
            except Exception:
                self.registered = False
            if not self.registered:
                # a refused registration (locked-memory limits) leaves a CUDA error behind that the next
                # torch call would report: clear it; the copies then go through staged (pageable) D2H
                try:
                    torch.cuda.cudart().cudaGetLastError()
                except Exception:
                    pass
        on_gpu = str(device) != "cpu"
        self.tbuf = torch.zeros((1 + self.max_blocks, 3), dtype=torch.int64, device=device)
        self.tstage = torch.zeros((1 + self.max_blocks, 3), dtype=torch.int64)
        self.tall_host = torch.zeros((self.world, 1 + self.max_blocks, 3), dtype=torch.int64)
        if on_gpu:
            self.tstage = self.tstage.pin_memory()
            self.tall_host = self.tall_host.pin_memory()
        self.tall = torch.zeros((self.world, 1 + self.max_blocks, 3), dtype=torch.int64, device=device)

    def exchange(self, table):
        """All ranks learn all block tables (fixed-shape all_gather; row 0 = number of blocks)."""
        import torch
        nb = int(table.shape[0])
        if nb > self.max_blocks:
            raise ValueError("SharedStream capacity exceeded")
        self.tstage[0, 0] = nb
        if nb:
            self.tstage[1:1 + nb] = torch.from_numpy(np.ascontiguousarray(table, dtype=np.int64))
        self.tbuf.copy_(self.tstage, non_blocking=True)
        self.dist.all_gather_into_tensor(self.tall, self.tbuf) if hasattr(self.dist, "all_gather_into_tensor") and str(self.device) != "cpu" \
            else self.dist.all_gather(list(self.tall.unbind(0)), self.tbuf)
        self.tall_host.copy_(self.tall)
        a = self.tall_host.numpy()
        return [a[r, 1:1 + int(a[r, 0, 0])] for r in range(self.world)]

    def finish(self, level, total, cc):
        """Rank 0: stream header and trailer around the blocks; everyone: closing barrier."""
        if self.rank == 0:
            if total + 10 > self.capacity:
                raise ValueError("SharedStream too small")
            self.view[0:4] = np.frombuffer(b"BZh" + bytes([ord("0") + level]), dtype=np.uint8)
            self.view[total:total + 10] = np.frombuffer(bytes([0x17, 0x72, 0x45, 0x38, 0x50, 0x90]) + cc.to_bytes(4, "big"), dtype=np.uint8)
        self.dist.barrier()
        return total + 10

    def close(self):
        import os
        import torch
        if self.registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(self.ptr)
            except Exception:
                pass
        self.view = None
        try:
            self.mm.close()
        except BufferError:
            pass
        os.close(self.fd)
        self.dist.barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


# ----------------------------------------------------------------------------------------------
# Decompression: the blocks of ONE .bz2 file over several decoders / GPUs.
#
# Once the scanner (lbz_scan_blocks, reference src/parse.c:281-342) has found the block magics,
# blocks are independent: candidate i goes to rank i mod world.  Every rank decodes its share
# speculatively (lbz_decoder_decode_at), the small per-block tables are all-gathered, every rank
# runs the same framing walk on the merged table (lbz_walk_table; src/parse.c:147-263 and the
# per-block checks of src/expand.c:725-736) and so knows which of its candidates are real blocks
# and where their bytes go, writes them (lbz_decoder_emit_at) and sends (offset, bytes, CRC) to
# rank 0, which checks the CRCs in stream order.  No collective on the data path besides that
# gather.  Works over any torch.distributed backend (gloo in the CPU test).

def share_of_candidates(hits, n_bytes, rank, world):
    """This rank's candidates.  Work per candidate grows with the compressed bytes up to the next
    candidate, and streams written by lbzip2 alternate large blocks with tiny spill blocks, so a
    plain round robin would give every other rank all the large ones: candidates are dealt out
    largest span first to the least loaded rank (the same deal on every rank), kept in stream order."""
    if world == 1:
        return list(hits)
    ends = list(hits[1:]) + [8 * n_bytes]
    fixed = 400_000                                   # per-candidate cost (tables, launches) in "bits": also evens out the counts
    order = sorted(range(len(hits)), key=lambda i: (-(ends[i] - hits[i]), i))
    load = [0] * world
    owner = [0] * len(hits)
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += ends[i] - hits[i] + fixed
    return [h for i, h in enumerate(hits) if owner[i] == rank]


def _row(b):
    return (int(b.pos), int(b.end_bit), int(b.out_len), int(b.status), int(b.block_size), int(b.rl_state),
            int(b.bwt_idx), int(b.rand))


def _decode_share(dec, z, mine):
    # decoders that keep the scanned input on the device skip the second upload
    if getattr(dec, "keeps_scanned_input", False):
        return dec.decode_at(z, mine, resident=True)
    return dec.decode_at(z, mine)


def sharded_decompress(dist, dec, z, rank, world, dblock_type, noemit=0xFFFFFFFFFFFFFFFF, gather_payload=True, keep=None,
                       out=None):
    """Decompress the file `z` (bytes, present on every rank) with `world` decoders.
    `dec` offers scan / decode_at / emit_at / walk_table (lbzip2_b200.Decoder).
    Returns (status, output bytes, info) on rank 0 and (the same status, None, info) elsewhere.
    Status and surviving output follow lbz_decompress_stream.
    gather_payload=False: the decoded bytes stay in every rank's host memory (a multi-process writer
    would pwrite them at their offsets); only (offset, length, CRC) per block travel to rank 0, which
    still checks the CRCs in stream order; rank 0 then returns None for the output.  `keep` (a dict)
    receives this rank's (parts, payload) for checks by the caller.  `out`: a uint8 array (page-locked:
    api.PinnedArray) that receives this rank's decoded bytes instead of a fresh buffer per call.
    The stream is uploaded once: the decode of the share works on the bytes the scanner left on
    the device."""
    hits = dec.scan(z)
    mine = share_of_candidates(hits, len(z), rank, world)
    cap_blocks = getattr(dec, "max_blocks", None)
    if cap_blocks is not None and len(mine) > cap_blocks:
        # every candidate of the share stays resident between decode_at and emit_at (the framing walk
        # over ALL ranks' tables decides which candidates are blocks): size the decoder for the share
        raise ValueError("decoder holds %d blocks, this rank's share of the file has %d candidates: create the "
                         "Decoder with max_blocks >= ceil(candidates / world)" % (cap_blocks, len(mine)))
    rows = [_row(b) for b in _decode_share(dec, z, mine)]
    every = [None] * world
    if world > 1:
        dist.all_gather_object(every, rows)
    else:
        every = [rows]
    merged = sorted((r + (src,) for src in range(world) for r in every[src]), key=lambda r: r[0])
    table = []
    for r in merged:
        b = dblock_type()
        b.pos, b.end_bit, b.out_len, b.status, b.block_size, b.rl_state, b.bwt_idx, b.rand = r[:8]
        table.append(b)
    status, chain, chain_crc, info = dec.walk_table(z, table)
    if status == 1:                       # LBZ_MORE cannot happen: every block magic is a scanner hit
        raise RuntimeError("a block of the stream was not among the scanner's candidates")
    # global output offsets of the confirmed blocks, and this rank's share of them
    goff, o = {}, 0
    for k in chain:
        goff[merged[k][0]] = o
        o += merged[k][2]
    local_off, cursor = [], 0
    for b in rows:
        if b[0] in goff:
            local_off.append(cursor)
            cursor += b[2]
        else:
            local_off.append(noemit)
    # a failure on one rank (e.g. its output buffer is too small) must not leave the others waiting
    # in a collective: it travels with the gathered parts and is raised everywhere afterwards
    failure = None
    try:
        payload, crcs = dec.emit_at(local_off, max(cursor, 1)) if out is None else dec.emit_at(local_off, max(cursor, 1), out=out)
    except Exception as ex:
        failure, payload, crcs = "rank %d: %s" % (rank, ex), b"", [0] * len(rows)
    parts = [(goff[b[0]], b[2], local_off[i], crcs[i]) for i, b in enumerate(rows) if b[0] in goff]
    if keep is not None:
        keep["parts"], keep["payload"] = parts, payload
    gathered = [None] * world
    if world > 1:
        dist.gather_object((parts, payload if gather_payload else None, failure), gathered if rank == 0 else None, dst=0)
    else:
        gathered = [(parts, payload if gather_payload else None, failure)]
    if rank != 0:
        final = [None]
        dist.broadcast_object_list(final, src=0)      # a CRC error found by rank 0 outranks the walk's verdict
        if final[0][2]:
            raise RuntimeError(final[0][2])
        info.status, info.num_blocks = final[0][:2]
        return final[0][0], None, info
    failures = [g[2] for g in gathered if g[2]]
    if failures:
        if world > 1:
            dist.broadcast_object_list([(status, 0, "; ".join(failures))], src=0)
        raise RuntimeError("; ".join(failures))
    # rank 0: CRCs in stream order; the first bad block ends the output (src/expand.c:731-736)
    got = {}
    for prt, pay, _ in gathered:
        for g, ln, lo, crc in prt:
            got[g] = (pay[lo:lo + ln] if pay is not None else None, crc)
    out, nblocks = [], 0
    for k, want in zip(chain, chain_crc):
        data, crc = got[goff[merged[k][0]]]
        if crc != want:
            status = 15                   # LBZ_ERR_BLKCRC
            break
        out.append(data)
        nblocks += 1
    info.num_blocks = nblocks
    info.status = status
    if world > 1:
        dist.broadcast_object_list([(status, nblocks, None)], src=0)
    return status, (b"".join(out) if gather_payload else None), info
