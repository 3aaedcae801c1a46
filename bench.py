#!/usr/bin/env python
"""bench.py -- input MB/s of bzip2 block compression at -9 on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (RLE1+CRC -> BWT -> MTF/RLE2 -> multi-table
Huffman -> bit-pack) over one batch: 100 MB of enwik8-shaped synthetic text per
GPU at -9 (BASELINE.json configs[1]; for N>1 every rank gets its own 100 MB
stream of the same generator family = weak scaling of configs[2], and the
byte-aligned block bitstreams are gathered to rank 0 in stream order over NCCL).

  value  : whole-job input MB/s with the input already resident in HBM
           (lbz_compress_chunks_device), timed with CUDA events, max over ranks
  e2e    : same metric through the public host API (lbz_compress_chunks) from
           PINNED HOST memory: H2D of the input and D2H of the .bz2 bytes are
           inside the timed region
  roofline / cpu_baseline / clocks : see DESIGN.md "Measurement"

--impl reference times the reference's own CPU implementation (the unmodified
lbzip2 CLI built into oracle/_ref by oracle/Makefile; the oracle port if that
binary is absent) on the same workload with all host threads.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

MB = 1_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size-mb", type=int, default=100, help="raw MB per GPU per step")
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--workload", default="text", choices=["text", "random", "runs_fib"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-decompress", action="store_true", help="skip the decompression leg")
    ap.add_argument("--batch-chunks", type=int, default=560,
                    help="engine capacity in chunks; larger inputs run as consecutive batches (560 chunks = 37 GB of HBM at -9)")
    ap.add_argument("--ref-sample-mb", type=int, default=0,
                    help="reference arm / cpu_baseline: MB of the workload per step (0 = choose: >= 1000 MB when time allows)")
    return ap.parse_args()


def workload_name(a):
    """One string for both arms (the driver compares config.workload of the two lines)."""
    shape = {"text": "enwik8-shaped synthetic text", "random": "uniform random bytes (PCG64)",
             "runs_fib": "single-byte run + Fibonacci word (tests/fib.c)"}[a.workload]
    return "%d MB %s per GPU at -%d" % (a.size_mb, shape, a.level)


def make_input(workload, nbytes, rank, procs=None):
    """Synthetic input of rank `rank`.  Text up to 100 MB is one stream of the generator (offset =
    rank); larger text inputs are consecutive 100 MB streams of the same family (offsets
    1000 * (rank + 1) + j), generated in parallel worker processes -- call before CUDA init."""
    import synth
    if workload == "text":
        if nbytes <= 100 * MB:
            return synth.text(nbytes, offset=rank)
        return synth.text_streams(nbytes, first_offset=1000 * (rank + 1), procs=procs)
    if workload == "random":
        return synth.random_bytes(nbytes, seed=1 + rank)
    return synth.runs_and_fib(nbytes)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def ncu_traffic(kernel, elements):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the
    newest committed `ncu --set full` capture of a launch over the same number of elements
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep); null when no
    capture matches -- the figure is never a constant in this file."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p) or not elements:
        return None
    best = None
    for ent in json.load(open(p)):
        if ent.get("kernel") == kernel and abs(ent.get("elements", 0) - elements) <= 0.005 * elements:
            if best is None or ent.get("order", 0) > best.get("order", 0):
                best = ent
    return int(best["dram_bytes"]) if best else None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------- reference arm ---
def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "lbzip2")
    return p if os.path.exists(p) else None


def time_reference(data, level, steps, warmup, threads):
    """Time the reference CPU implementation on `data` (bytes).  Returns (ms list, kind, sha256)."""
    binp = ref_binary()
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(tmp, "lbz_bench_%d.raw" % os.getpid())
    with open(path, "wb") as f:
        f.write(data)
    times, sha = [], None
    try:
        if binp:
            kind = "reference"
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                if i == 0:
                    out = subprocess.run([binp, "-%d" % level, "-n%d" % threads, "-c", path], stdout=subprocess.PIPE, check=True).stdout
                    sha = hashlib.sha256(out).hexdigest()
                else:
                    with open(os.devnull, "wb") as dn:
                        subprocess.run([binp, "-%d" % level, "-n%d" % threads, "-c", path], stdout=dn, check=True)
                dt = (time.perf_counter() - t0) * 1e3
                if i >= warmup:
                    times.append(dt)
        else:
            import orclib
            kind = "port"
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                out, _ = orclib.orc_stream(data, level)
                dt = (time.perf_counter() - t0) * 1e3
                sha = hashlib.sha256(out).hexdigest()
                if i >= warmup:
                    times.append(dt)
    finally:
        os.unlink(path)
    return times, kind, sha


def reference_sample(a, own_data=None):
    """The bytes the reference CLI is timed on: a bounded sample of the workload, at least 1 GB
    when the CLI is available (a 100 MB batch is only 112 work units: with 16-32 host threads
    the tail of the run idles and process start-up weighs in), 10 MB for the scalar oracle port."""
    if not ref_binary():
        d = own_data if own_data is not None else make_input(a.workload, min(a.size_mb, 10) * MB, 0)
        return d[: 10 * MB]
    mb = a.ref_sample_mb or max(a.size_mb, 1000)
    if own_data is not None and len(own_data) >= mb * MB:
        return own_data[: mb * MB]
    return make_input(a.workload, mb * MB, 0)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    binp = ref_binary()
    threads = cores if binp else 1
    sample = reference_sample(a)
    # keep the whole run within a few minutes: one calibration pass, then shrink the sample if needed
    t_cal, _, _ = time_reference(sample, a.level, 1, 0, threads)
    budget_s = 170.0
    est = t_cal[0] / 1e3 * (a.steps + a.warmup)
    if est > budget_s and len(sample) > 100 * MB:
        keep = max(100, int(len(sample) / MB * budget_s / est) // 100 * 100)
        sample = sample[: keep * MB]
    times, kind, _ = time_reference(sample, a.level, a.steps, a.warmup, threads)
    ms = float(np.mean(times))
    val = len(sample) / MB / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "input MB/s at -9 (bit-exact .bz2)", "value": round(val, 2), "unit": "MB/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": round(val, 2), "unit": "MB/s", "cores": threads, "kind": kind,
                         "sample": "%d MB of the workload per step, lbzip2 -%d -n%d from /dev/shm to /dev/null" % (len(sample) // MB, a.level, threads)
                         if binp else "10 MB of the workload, scalar oracle port"},
        "e2e": {"value": round(val, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ decompress leg ---
def decompress_leg(a, lbzip2_b200, L, local, stream, data, recs):
    """Output MB/s of lbz_decompress_stream on the .bz2 the compress leg produced (same 100 MB).
    value: compressed bytes already in HBM, decoded bytes left in HBM; e2e: pinned host in/out."""
    import torch
    from lbzip2_b200 import api
    nz, nbytes = len(stream), len(data)
    steps, warm = min(a.steps, 5), min(a.warmup, 3)
    dec = lbzip2_b200.Decoder(device=local, max_blocks=len(recs) + 8, in_cap=nz + 64, out_cap=nbytes + (1 << 20))
    h_z = L.lbz_host_alloc(nz)
    h_o = L.lbz_host_alloc(nbytes + 64)
    C.memmove(h_z, stream, nz)
    hbm_peak, peak_kind = peaks()

    def run(flags, out_ptr):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        st, n, info = dec.decompress_ptr(h_z, nz, out_ptr, nbytes + 64, flags)
        ev1.record()
        ev1.synchronize()
        if st != 0 or n != nbytes:
            raise RuntimeError("decoder returned status %d, %d bytes" % (st, n))
        return ev0.elapsed_time(ev1), info

    def timed(flags, out_ptr):
        for _ in range(warm):
            run(flags, out_ptr)
        torch.cuda.synchronize()
        l0 = dec.launches
        ms, stage, info = [], {}, None
        for _ in range(steps):
            t, info = run(flags, out_ptr)
            ms.append(t)
            for k, v in dec.stage_ms().items():
                stage[k] = stage.get(k, 0.0) + v / steps
        torch.cuda.synchronize()
        return sum(ms) / steps, stage, info, (dec.launches - l0) // steps

    dec.load(h_z, nz)
    ms_dev, stage, info, launches = timed(api.D_RESIDENT_INPUT | api.D_DEVICE_OUTPUT, None)
    dev_sha = hashlib.sha256(dec.array(api.DA_OUT, 0, nbytes).tobytes()).hexdigest()
    ms_host, _, _, _ = timed(0, h_o)
    host_out = bytes((C.c_uint8 * nbytes).from_address(h_o))
    want = hashlib.sha256(data).hexdigest()
    # stage-interface bytes of the decode path: z + n' (prefix decode: bits in, last column out)
    # + 5n' (successor table) + 5n' (walks: nodes in, text out) + n' + n (run expansion)
    nprime = sum(r.nblock for r in recs)
    path_bytes = nz + 12 * nprime + nbytes
    res = {
        "metric": "output MB/s of batch decompression of the same stream", "unit": "MB/s",
        "value": round(nbytes / MB / (ms_dev / 1e3), 2), "ms_per_step": round(ms_dev, 3),
        "e2e": {"value": round(nbytes / MB / (ms_host / 1e3), 2), "unit": "MB/s", "ms_per_step": round(ms_host, 3),
                "h2d_bytes_per_step": nz, "d2h_bytes_per_step": nbytes, "api": "lbz_decompress_stream (pinned host in/out)"},
        "steps": steps, "warmup": warm, "gpu_launches": int(launches),
        "blocks": int(info.num_blocks), "candidates": int(info.candidates), "waves": int(info.waves),
        "stage_ms": {k: round(v, 3) for k, v in stage.items()},
        "stage_names": "upload | scan (block magics) | retrieve (header, tables, length walk) | successors (symbols, runs, inverse MTF, counting sort) | walks (inverse BWT) | expand (run expansion, CRC) | tail",
        "path_roofline": {"model": "z + 12n' + n per block", "bytes_per_step": int(path_bytes),
                          "achieved": round(path_bytes / (ms_dev / 1e3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                          "frac": round(path_bytes / (ms_dev / 1e3) / 1e9 / hbm_peak, 5), "peak_kind": peak_kind},
        "verified": {"device_output_sha256_equals_input": dev_sha == want,
                     "host_output_equals_input": host_out == data},
        "device_bytes": int(dec.device_bytes),
    }
    # dominant kernel of this leg: the per-block walk over the code lengths (k_ub_chain2, timed with
    # the header and table kernels as stage "retrieve", CUDA events on the decoder's stream).  It reads
    # the compressed bits once and writes 9 bytes per 50-code group; it is latency-bound by
    # construction (one walking warp per block), the fraction is reported for completeness.
    retrieve_ms = float(stage.get("retrieve", 0.0))
    walk_bytes = nz + 9 * ((sum(r.nmtf for r in recs) + 49 * len(recs)) // 50)
    walk_gbs = walk_bytes / (retrieve_ms / 1e3) / 1e9 if retrieve_ms > 0 else 0.0
    res["roofline"] = {"bound": "hbm", "kernel": "k_ub_chain2 (+ k_ub_header, k_ub_tree_*): stage 'retrieve'",
                       "achieved": round(walk_gbs, 2), "peak": hbm_peak, "unit": "GB/s",
                       "frac": round(walk_gbs / hbm_peak, 6), "bytes_per_launch": int(walk_bytes),
                       "avg_launch_ms": round(retrieve_ms, 3),
                       "traffic": ncu_traffic("k_ub_chain2", len(recs)),
                       "note": "latency-bound: one walking warp per block, a dependent chain of shift -> shared-memory byte -> "
                               "shift per table step (profiles/r02_ncu_chain2.txt); traffic = dram read+write of that capture"}
    dec.close()
    dec = None
    # The walk over the code lengths is serial per block, so decode time at 223 blocks is latency;
    # the same stream four times over (a concatenated .bz2, 4x the blocks in one wave) shows what
    # the kernels do with more blocks in flight.
    try:
        if nbytes > 200 * MB:
            raise RuntimeError("skipped for inputs > 200 MB (the stream already holds > 400 blocks)")
        reps = 4
        z4 = stream * reps
        dec4 = lbzip2_b200.Decoder(device=local, max_blocks=reps * len(recs) + 8, in_cap=len(z4) + 64,
                                   out_cap=reps * nbytes + (1 << 20))
        h_z4 = L.lbz_host_alloc(len(z4))
        C.memmove(h_z4, z4, len(z4))
        dec4.load(h_z4, len(z4))
        ms4 = []
        for i in range(warm + steps):
            st, n4, info4 = dec4.decompress_ptr(h_z4, len(z4), None, reps * nbytes + 64,
                                                api.D_RESIDENT_INPUT | api.D_DEVICE_OUTPUT)
            if st != 0 or n4 != reps * nbytes:
                raise RuntimeError("decoder returned status %d, %d bytes" % (st, n4))
            if i >= warm:
                ms4.append(dec4.last_ms)
        part = hashlib.sha256(dec4.array(api.DA_OUT, (reps - 1) * nbytes, nbytes).tobytes()).hexdigest()
        res["more_blocks_in_flight"] = {
            "workload": "the same stream %d times over as one concatenated file" % reps, "blocks": int(info4.num_blocks),
            "waves": int(info4.waves), "value": round(reps * nbytes / MB / (sum(ms4) / len(ms4) / 1e3), 2), "unit": "MB/s",
            "ms_per_step": round(sum(ms4) / len(ms4), 3), "timed": "CUDA events on the decoder's stream, HBM-resident",
            "last_copy_sha256_equals_input": part == want}
        L.lbz_host_free(h_z4)
        dec4.close()
    except Exception as ex:
        res["more_blocks_in_flight"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    binp = ref_binary()
    if binp and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        path = os.path.join(tmp, "lbz_bench_%d.bz2" % os.getpid())
        with open(path, "wb") as f:
            f.write(stream)
        try:
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                with open(os.devnull, "wb") as dn:
                    subprocess.run([binp, "-d", "-n%d" % cores, "-c", path], stdout=dn, check=True)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        finally:
            os.unlink(path)
        res["cpu_baseline"] = {"value": round(nbytes / MB / best, 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                               "sample": "the whole stream, lbzip2 -d -n%d, /dev/shm -> /dev/null, best of 3" % cores}
    L.lbz_host_free(h_z)
    L.lbz_host_free(h_o)
    return res


def decompress_leg_multi(a, lbzip2_b200, dist, dev, local, rank, world, sink, end, nblocks_total, total_plain):
    """N > 1: the blocks of the ONE .bz2 the host leg assembled (shared stream) over the N decoders:
    candidate i -> rank i mod N, tables all-gathered, same framing walk everywhere, every rank writes
    its blocks (sharding.sharded_decompress).  Timed end to end per rank (wall clock between
    barriers, compressed bytes from host memory, decoded bytes back in each rank's host memory),
    max over ranks.  The decoded bytes are not concatenated: (offset, length, CRC) per block go to
    rank 0, which checks the CRC chain -- what a multi-process writer needs to pwrite them."""
    import torch
    from lbzip2_b200 import api, sharding
    z = sink.view[:end]                      # the shared mapping itself (CUDA-registered where that worked): no copy
    share = min(nblocks_total + 16, 2 * ((nblocks_total + world - 1) // world) + 64)
    dec = lbzip2_b200.Decoder(device=local, max_blocks=share, in_cap=len(z) + 64,
                              out_cap=2 * (total_plain // world) + 8 * 900000 + (1 << 20))
    out_cap = 2 * (total_plain // world) + 8 * 900000 + (1 << 20)
    pinned = api.PinnedArray(out_cap, lbzip2_b200.api.load_library())
    times, keep, st = [], {}, None
    for i in range(1 + 3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        st, _, info = sharding.sharded_decompress(dist, dec, z, rank, world, api.DBlock, gather_payload=False, keep=keep, out=pinned.a)
        torch.cuda.synchronize(); dist.barrier()
        dt = (time.perf_counter() - t0) * 1e3
        if i:
            times.append(dt)
    t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # every rank: sha256 of each decoded part; rank 0 compares with the reference CLI's output of the stream
    mine = [(g, ln, hashlib.sha256(bytes(keep["payload"][lo:lo + ln])).hexdigest()) for g, ln, lo, _ in keep["parts"]]
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    dec.close()
    keep.clear()
    pinned.close()
    if rank != 0:
        return None
    res = {"metric": "output MB/s of sharded batch decompression of the assembled stream", "unit": "MB/s",
           "value": round(total_plain / MB / (ms / 1e3), 2), "ms_per_step": round(ms, 3), "steps": 3, "warmup": 1,
           "status": int(st), "blocks": int(info.num_blocks), "n_gpus": world,
           "timed": "wall clock between barriers, max over ranks; compressed stream in (shared, CUDA-registered) host memory, "
                    "uploaded once per rank, decoded bytes in each rank's page-locked host memory, block table + CRCs gathered to rank 0 (CRC chain checked there)"}
    binp = ref_binary()
    if binp and total_plain <= 2000 * MB:
        plain = subprocess.run([binp, "-d", "-c", sink.path], stdout=subprocess.PIPE, check=True).stdout
        ok = len(plain) == total_plain
        covered = 0
        for parts in allp:
            for g, ln, h in parts:
                ok = ok and hashlib.sha256(plain[g:g + ln]).hexdigest() == h
                covered += ln
        res["verified"] = {"every_decoded_block_equals_reference_cli_output": bool(ok and covered == total_plain)}
    return res


# ------------------------------------------------------------------------ our arm ---
def run_ours(a):
    import torch
    import torch.distributed as dist
    import lbzip2_b200
    from lbzip2_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the product has no CPU path (use --impl reference for the CPU arm)")
    # the input is generated before CUDA is initialised (large text inputs fork worker processes)
    data = make_input(a.workload, a.size_mb * MB, rank, procs=max(1, (os.cpu_count() or 1) // world))
    # reference CPU path beside the GPU numbers (rank 0 at N=1 only): timed first, while the GPU is
    # still idle, on a bounded sample of the workload (>= 1 GB: all host threads stay busy)
    cpu_base = None
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        have_cli = ref_binary() is not None
        sample = reference_sample(a, own_data=data)
        times, kind, _ = time_reference(sample, a.level, 3, 1, cores if have_cli else 1)
        cpu_base = {"value": round(len(sample) / MB / (min(times) / 1e3), 2), "unit": "MB/s",
                    "cores": cores if have_cli else 1, "kind": kind,
                    "sample": ("%d MB of the workload, lbzip2 -%d -n%d, /dev/shm -> /dev/null, best of 3" % (len(sample) // MB, a.level, cores))
                    if have_cli else "10 MB of the batch, scalar oracle port"}
        del sample
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    level, mbs = a.level, a.level * 100000
    nbytes = a.size_mb * MB
    nchunks = (nbytes + mbs - 1) // mbs
    eng_chunks = min(nchunks, max(1, a.batch_chunks))      # larger inputs: consecutive batches inside one call
    eng = lbzip2_b200.Engine(device=local, level=level, max_chunks=eng_chunks)
    L = eng.L
    cap = L.lbz_bound(nbytes) + 64

    # pinned host buffers for the e2e leg
    h_in = L.lbz_host_alloc(nbytes)
    h_out = L.lbz_host_alloc(cap)
    C.memmove(h_in, data, nbytes)
    # HBM-resident buffers for the kernel-throughput leg
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gth = sharding.BlockGatherer(dist, dev, max_blocks=2 * nchunks + 2, max_payload=cap) if world > 1 else None
    # N > 1, host leg: ONE stream-ordered .bz2 in host memory shared by the ranks (a /dev/shm file
    # mapped by every rank, registered with CUDA); completed inside the timed region of every step
    sink = None                                   # created after the device leg, when the stream size is known

    def gather(recs, payload_dev):
        """Device leg: NCCL gather of the block bitstreams into rank 0's HBM in one
        point-to-point transfer per rank (stream order is then a table lookup)."""
        if world == 1:
            return None
        return gth.gather(sharding.block_table(recs, mbs), payload_dev, tables_only=False)

    def step_device():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        n_out, recs = eng.compress_chunks_ptr(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, device=True)
        g = gather(recs, d_out[:n_out])
        ev1.record()
        ev1.synchronize()
        return ev0.elapsed_time(ev1), eng.last_ms, n_out, recs, g

    def step_host():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if world == 1:
            n_out, recs = eng.compress_chunks_ptr(h_in, nbytes, h_out, cap, device=False)
            g = None
        else:
            # host in -> blocks in HBM; block tables of all ranks; every rank copies its blocks to their
            # offsets in the shared stream; rank 0 adds header and trailer; closing barrier
            n_out, recs = eng.compress_chunks_ptr(h_in, nbytes, d_out.data_ptr(), cap, h2d=True)
            table = sharding.block_table(recs, mbs)
            every = sink.exchange(table)
            offs, total, cc = sharding.place_blocks(every, world)
            src = np.concatenate(([0], np.cumsum(table[:, 1])))[:-1]
            eng.scatter_to_host(d_out.data_ptr(), src, sink.ptr, offs[rank], table[:, 1])
            g = sink.finish(level, total, cc)
        ev1.record()
        ev1.synchronize()
        return ev0.elapsed_time(ev1), eng.last_ms, n_out, recs, g

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        sampler = ClockSampler(local)
        barrier()
        l0 = eng.launches
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        ms, eng_ms, stage, k0 = [], [], {}, [0.0, 0, 0]
        last = None
        for _ in range(a.steps):
            last = fn()
            ms.append(last[0])
            eng_ms.append(last[1])
            for k, v in eng.stage_ms().items():
                stage[k] = stage.get(k, 0.0) + v
            s = eng.k0_stats()
            k0[0] += s[0]; k0[1] += s[1]; k0[2] = s[2]
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop() if rank == 0 else None
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return dict(total_ms=float(tot.item()), wall_ms=wall, launches=eng.launches - l0, stage=stage, k0=k0,
                    clocks=clocks, last=last, eng_ms=sum(eng_ms))

    rd = timed(step_device)
    if world > 1:
        # the output is deterministic: the device leg tells how long the assembled stream is, so the
        # shared mapping (registered with CUDA by every rank) is no larger than it has to be
        tot_out = torch.tensor([rd["last"][2]], dtype=torch.int64, device=dev)
        dist.all_reduce(tot_out)
        sink = sharding.SharedStream(dist, dev, "/dev/shm/lbz_bench_stream_%s.bz2" % os.environ.get("MASTER_PORT", "0"),
                                     int(tot_out.item()) + (1 << 20), max_blocks=2 * nchunks + 2)
    rh = timed(step_host)
    sort_rounds = eng.last_rounds
    dev_bytes = eng.device_bytes

    # The dominant kernel is timed ALONE for the roofline: the production engine runs
    # two lanes whose kernels overlap, which stretches every individual launch, so a
    # second, single-lane engine repeats the HBM-resident step and its CUDA-event
    # launch times are used (same kernels, same data).
    k0_single = None
    if world == 1:
        prev = os.environ.get("LBZ_LANES")
        os.environ["LBZ_LANES"] = "1"
        eng.close()                                       # one engine's working set at a time (37 GB at 560 chunks)
        eng1 = lbzip2_b200.Engine(device=local, level=level, max_chunks=eng_chunks)
        if prev is None:
            del os.environ["LBZ_LANES"]
        else:
            os.environ["LBZ_LANES"] = prev
        acc = [0.0, 0, 0]
        reps1 = (a.warmup + a.steps) if nbytes <= 200 * MB else 2
        for i in range(reps1):
            eng1.compress_chunks_ptr(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, device=True)
            if i >= min(a.warmup, reps1 - 1):
                s1 = eng1.k0_stats()
                acc[0] += s1[0]; acc[1] += s1[1]; acc[2] = s1[2]
        k0_single = acc
        eng1.close()

    # ---- correctness of what was timed (outside the timed region) -------------------
    n_out, recs = rd["last"][2], rd["last"][3]
    verified = {}
    ref_sha_own = None
    stream = None
    big = nbytes > 200 * MB
    binp = ref_binary()
    if not a.no_verify:
        # every rank: the blocks it produced, framed as a stream of their own, against the reference
        # CLI on the same bytes (bit-exact), or a bz2 round trip when the CLI is absent
        import bz2
        body = d_out[:n_out].cpu().numpy().tobytes()
        cc = 0
        for r in recs:
            cc = (((cc << 1) & 0xFFFFFFFF) ^ (cc >> 31) ^ r.crc ^ 0xFFFFFFFF) & 0xFFFFFFFF
        stream = b"BZh" + bytes([48 + level]) + body + bytes([0x17, 0x72, 0x45, 0x38, 0x50, 0x90]) + cc.to_bytes(4, "big")
        own = {"sha256": hashlib.sha256(stream).hexdigest()}
        if world == 1:
            host_body = bytes((C.c_uint8 * rh["last"][2]).from_address(h_out))
            own["host_equals_device_path"] = host_body == body
            del host_body
        else:
            own["host_equals_device_path"] = None       # N > 1: rank 0 compares the shared stream as a whole below
        if binp:
            tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            path = os.path.join(tmp, "lbz_bench_own_%d.raw" % os.getpid())
            with open(path, "wb") as f:
                f.write(data)
            try:
                cores = max(1, (os.cpu_count() or 1) // world)
                t0 = time.perf_counter()
                ref_out = subprocess.run([binp, "-%d" % level, "-n%d" % cores, "-c", path], stdout=subprocess.PIPE, check=True).stdout
                own["ref_cli_s"] = round(time.perf_counter() - t0, 3)
            finally:
                os.unlink(path)
            ref_sha_own = hashlib.sha256(ref_out).hexdigest()
            own["bit_exact_vs_reference_cli"] = ref_sha_own == own["sha256"]
            del ref_out
        if not big or not binp:
            own["roundtrip"] = bz2.decompress(stream) == data
        own["periodic_blocks"] = sum(1 for r in recs if r.tie_count > 1)
        if world == 1:
            verified = own
        else:
            allv = [None] * world
            dist.all_gather_object(allv, own)
            if rank == 0:
                verified["bit_exact_vs_reference_cli_every_rank"] = all(v.get("bit_exact_vs_reference_cli", False) for v in allv) if binp else None
                verified["periodic_blocks"] = sum(v["periodic_blocks"] for v in allv)
                verified["rank_sha256"] = [v["sha256"][:16] for v in allv]
    all_sha = [hashlib.sha256(data).hexdigest()]
    if world > 1:
        all_sha = [None] * world
        dist.all_gather_object(all_sha, hashlib.sha256(data).hexdigest())
    if rank == 0 and not a.no_verify and world > 1:
        import bz2
        tables, payloads = sharding.to_host(*rd["last"][4])
        gstream = sharding.assemble_stream(level, tables, payloads, world)
        # the host leg's product: the shared stream every rank wrote its blocks into during the last step
        verified["host_stream_equals_gathered_stream"] = bytes(sink.view[: rh["last"][4]]) == gstream
        if not big:
            # chunk i of the job is chunk i // world of rank i % world: de-interleave the
            # decoded stream and compare every rank's part with the sha256 of its input
            plain = bz2.decompress(gstream)
            ok = len(plain) == world * nbytes
            lens = [min(mbs, nbytes - (i // world) * mbs) for i in range(world * nchunks)]
            offs = [0]
            for ln in lens:
                offs.append(offs[-1] + ln)
            for r in range(world):
                part = b"".join(plain[offs[i]:offs[i + 1]] for i in range(r, world * nchunks, world))
                ok = ok and hashlib.sha256(part).hexdigest() == all_sha[r]
            verified["gathered_stream_roundtrip"] = ok
        elif binp:
            # large inputs: the gathered stream passes the reference CLI's integrity test (every block
            # CRC and the combined CRC, which depends on the block order), and its head decodes to
            # the first chunks in round-robin order (rank 0's first chunk, rank 1's, ...)
            tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            path = os.path.join(tmp, "lbz_bench_gathered_%d.bz2" % os.getpid())
            with open(path, "wb") as f:
                f.write(gstream)
            try:
                verified["gathered_stream_passes_lbzip2_t"] = subprocess.run([binp, "-t", path]).returncode == 0
                p = subprocess.Popen([binp, "-d", "-c", path], stdout=subprocess.PIPE)
                head = p.stdout.read(mbs)
                p.stdout.close()
                p.kill()
                p.wait()
                verified["gathered_stream_head_is_rank0_chunk0"] = head == data[:mbs]
            finally:
                os.unlink(path)
        verified["gathered_stream_bytes"] = len(gstream)
        del gstream

    # ---- decompression leg: the stream just produced, back through the
    # batch decompressor; reported under "decompress", the headline stays the compressor ----
    decomp = None
    if world == 1 and rank == 0 and not a.no_verify and not a.no_decompress:
        try:
            decomp = decompress_leg(a, lbzip2_b200, L, local, stream, data, recs)
        except Exception as ex:  # the compressor's line must survive a decoder problem
            decomp = {"error": "%s: %s" % (type(ex).__name__, ex)}

    decomp_multi = None
    if world > 1 and not a.no_verify and not a.no_decompress:
        try:
            nblk_total = torch.tensor([len(recs)], dtype=torch.int64, device=dev)
            dist.all_reduce(nblk_total)
            decomp_multi = decompress_leg_multi(a, lbzip2_b200, dist, dev, local, rank, world, sink, rh["last"][4],
                                                int(nblk_total.item()), world * nbytes)
        except Exception as ex:
            decomp_multi = {"error": "%s: %s" % (type(ex).__name__, ex)}
    if rank != 0:
        if world > 1:
            sink.close()
            dist.destroy_process_group()
        return

    ms_step = rd["total_ms"] / a.steps
    value = world * nbytes / MB / (ms_step / 1e3)
    ms_e2e = rh["total_ms"] / a.steps
    e2e = world * nbytes / MB / (ms_e2e / 1e3)
    hbm_peak, peak_kind = peaks()

    # dominant kernel: one LSD pass of the initial rotation sort (k_text_pass2):
    # per element it reads an 8-byte (key, index) pair and writes it to its sorted place
    k0_ms, k0_n, k0_elems = k0_single if k0_single else rd["k0"]
    k0_avg_ms = k0_ms / max(k0_n, 1)
    k0_bytes = 16.0 * k0_elems
    k0_gbs = k0_bytes / (k0_avg_ms / 1e3) / 1e9 if k0_avg_ms > 0 else 0.0
    # whole path: stage-interface model of SURVEY.md 8d: n + 13 n' + 20 nm + z per block
    path_bytes = sum(r.raw_len + 13 * r.nblock + 20 * r.nmtf + r.out_len for r in recs)
    path_gbs = path_bytes / (rd["eng_ms"] / a.steps / 1e3) / 1e9

    stage = {k: round(v / a.steps, 3) for k, v in rd["stage"].items()}
    line = {
        "metric": "input MB/s at -9 (bit-exact .bz2)", "value": round(value, 2), "unit": "MB/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a),
                   "chunks_per_gpu": nchunks, "blocks": len(recs), "batches_per_step": (nchunks + eng_chunks - 1) // eng_chunks,
                   "l2": "inputs (%d MB) and working set (%.1f GB) larger than L2" % (a.size_mb, dev_bytes / 1e9),
                   "generator": "tests/synth.py (text: seed 0x5EED, stream offset = rank; > 100 MB: consecutive 100 MB streams)",
                   "gather": ("device leg: NCCL gather of block bitstreams + block table into rank 0's HBM; host leg (e2e): "
                              "one stream-ordered .bz2 assembled inside every timed step in a host buffer shared by the "
                              "ranks (/dev/shm mapping%s): tables all-gathered, every rank's blocks copied D2H to their "
                              "stream offsets, header + trailer by rank 0, closing barrier"
                              % (", CUDA-registered" if sink.registered else ", not registered: staged copies")) if world > 1 else "none (single GPU)"},
        "e2e": {"value": round(e2e, 2), "unit": "MB/s", "ms_per_step": round(ms_e2e, 3),
                "h2d_bytes_per_step": world * nbytes, "d2h_bytes_per_step": int(rh["last"][4] - 14) if world > 1 else int(rh["last"][2]),
                "api": "lbz_compress_chunks (pinned host in/out)"},
        "gpu_launches": int(rd["launches"]),
        "clocks": rd["clocks"],
        "roofline": {"bound": "hbm", "kernel": "k_text_pass2 (one LSD pass of the initial rotation sort, 8 per batch)",
                     "achieved": round(k0_gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(k0_gbs / hbm_peak, 4),
                     "traffic": ncu_traffic("k_text_pass2", k0_elems), "peak_kind": peak_kind,
                     "bytes_per_launch": int(k0_bytes), "avg_launch_ms": round(k0_avg_ms, 4),
                     "timed": "single-lane engine, kernel alone on the GPU" if k0_single else "inside the two-lane step"},
        "path_roofline": {"model": "n + 13n' + 20nm + z per block (SURVEY.md 8d)", "bytes_per_step": int(path_bytes),
                          "achieved": round(path_gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(path_gbs / hbm_peak, 5)},
        "stage_ms_summed_over_lanes": stage, "sort_rounds": sort_rounds,
        "wall_ms_per_step": round(rd["wall_ms"] / a.steps, 3), "verified": verified,
        "compressed_ratio": round(nbytes / max(n_out, 1), 3),
    }
    if decomp is not None:
        line["decompress"] = decomp
    if decomp_multi is not None:
        line["decompress"] = decomp_multi
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))
    L.lbz_host_free(h_in)
    L.lbz_host_free(h_out)
    if world > 1:
        sink.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
