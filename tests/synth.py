"""Deterministic synthetic inputs of the shapes BASELINE.json names (no network:
enwik8 itself is unavailable).  Used by tests, smoke() and bench.py."""
import numpy as np


def _vocab(rng, V):
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    lp = 1.0 / np.arange(1, 27) ** 0.9
    lp /= lp.sum()
    # frequent words are short (Zipf's law of abbreviation): length grows with log rank
    lens = np.clip(1 + (np.log2(np.arange(V) + 2) * 0.30).astype(np.int64) + rng.geometric(1 / 1.7, V) - 1, 1, 12)
    words = []
    flat = letters[rng.choice(26, size=int(lens.sum()), p=lp)]
    o = 0
    for L in lens:
        words.append(flat[o:o + L].tobytes())
        o += L
    # a sprinkle of markup-like and numeric tokens at fixed ranks
    special = [b"[[", b"]]", b"{{", b"}}", b"&quot;", b"<ref>", b"</ref>", b"<title>", b"</title>",
               b"<page>", b"</page>", b"<id>", b"</id>", b"==", b"''", b"|", b"*", b"#", b"http://"]
    for k, s in enumerate(special):
        words[40 + 13 * k] = s
    for k in range(200):
        words[1000 + 37 * k] = str(int(rng.integers(0, 2100))).encode()
    return words


FOLLOW_P = 0.80
PHRASE_RATE = 0.10


def text(n, seed=0x5EED, offset=0):
    """`n` bytes of enwik8-shaped text: Zipf(1.0) vocabulary of 1e5 pseudo-words
    with first-order (bigram) dependence, sentence/paragraph punctuation, XML-ish
    markup tokens and occasional indents.  `offset` selects a different stream
    of the same generator family (for sharded multi-GPU inputs)."""
    rng = np.random.default_rng([seed, 0])
    V = 100_000
    words = _vocab(rng, V)
    p = 1.0 / np.arange(1, V + 1) ** 1.0
    p /= p.sum()
    cdf = np.cumsum(p)
    succ = np.searchsorted(cdf, rng.random((V, 4))).astype(np.int64)  # 4 preferred successors per word (Zipf-drawn)
    # recurring multi-word phrases (boilerplate): 20000 phrases of 3..12 words, Zipf-used
    NP = 3_000
    plen = rng.integers(3, 13, NP)
    pwords = np.searchsorted(cdf, rng.random(int(plen.sum())))
    poff = np.concatenate(([0], np.cumsum(plen)))
    pp = 1.0 / np.arange(1, NP + 1) ** 1.0
    pcdf = np.cumsum(pp / pp.sum())
    seps = [b" ", b" ", b" ", b" ", b" ", b" ", b", ", b". ", b".\n", b".\n\n", b"\n    ", b"; ", b" (", b") ", b": ", b"\n"]
    sp = np.array([30, 30, 30, 30, 30, 30, 10, 6, 3, 1.2, 0.8, 0.7, 1, 1, 1, 1.5], dtype=float)
    sp /= sp.sum()
    # flat byte tables: the token stream is assembled with numpy gathers (same bytes as
    # joining the Python strings, ~20x faster: 1 GB inputs for BASELINE configs 3-5)
    flat = np.frombuffer(b"".join(words) + b"".join(seps), dtype=np.uint8)
    wl = np.array([len(x) for x in words], dtype=np.int64)
    sl = np.array([len(x) for x in seps], dtype=np.int64)
    woff = np.concatenate(([0], np.cumsum(wl)))[:-1]
    soff = int(wl.sum()) + np.concatenate(([0], np.cumsum(sl)))[:-1]
    out = []
    got = 0
    srng = np.random.default_rng([seed, 1 + offset])
    while got < n:
        m = 400_000
        w = np.searchsorted(cdf, srng.random(m))
        follow = srng.random(m) < FOLLOW_P
        k = srng.integers(0, 4, m)
        idx = np.nonzero(follow[1:])[0] + 1
        for _ in range(3):                                           # resolve bigram chains up to depth 3
            w[idx] = succ[w[idx - 1], k[idx]]
        nph = int(m * PHRASE_RATE)
        at = srng.integers(0, m - 16, nph)
        which = np.searchsorted(pcdf, srng.random(nph))
        # phrase q overwrites w[at : at + plen[q]]; later phrases win where they overlap
        # (numpy assigns repeated indices in order)
        pl = plen[which]
        ramp = np.arange(int(pl.sum())) - np.repeat(np.concatenate(([0], np.cumsum(pl)))[:-1], pl)
        w[np.repeat(at, pl) + ramp] = pwords[np.repeat(poff[which], pl) + ramp]
        s = srng.choice(len(seps), size=m, p=sp)
        tl = np.empty(2 * m, dtype=np.int64)
        ts = np.empty(2 * m, dtype=np.int64)
        tl[0::2] = wl[w]; tl[1::2] = sl[s]
        ts[0::2] = woff[w]; ts[1::2] = soff[s]
        total = int(tl.sum())
        oo = np.concatenate(([0], np.cumsum(tl)))[:-1]
        blob = flat[np.repeat(ts - oo, tl) + np.arange(total)]
        out.append(blob)
        got += total
    return np.concatenate(out)[:n].tobytes()


def _text_job(args):
    n, off = args
    return text(n, offset=off)


def text_streams(n, first_offset=0, stream_bytes=100_000_000, procs=None):
    """`n` bytes of the text shape as consecutive streams of `stream_bytes` of the same generator
    family (stream j has offset first_offset + j) -- SURVEY.md 8d config 3: "the same generator
    (different stream offsets, same seed family)".  The streams are generated in parallel worker
    processes (fork: call this before CUDA is initialised)."""
    import multiprocessing as mp
    import os
    jobs = []
    left, j = n, 0
    while left > 0:
        jobs.append((min(stream_bytes, left), first_offset + j))
        left -= jobs[-1][0]
        j += 1
    procs = procs or max(1, min(len(jobs), (os.cpu_count() or 1)))
    if procs == 1 or len(jobs) == 1:
        return b"".join(_text_job(jb) for jb in jobs)
    with mp.get_context("fork").Pool(procs) as pool:
        return b"".join(pool.map(_text_job, jobs))


def random_bytes(n, seed=1):
    """/dev/urandom-shaped: iid uniform bytes, PCG64(seed)."""
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 256, n, dtype=np.uint8).tobytes()


def fib(n):
    """Fibonacci word over {a,b}, same sequence as reference tests/fib.c:28-33."""
    s = np.array([ord("a")], dtype=np.uint8)
    a, b = ord("a"), ord("b")
    while s.size < n:
        isa = s == a
        lens = np.where(isa, 2, 1)
        offs = np.concatenate(([0], np.cumsum(lens)))
        t = np.full(offs[-1], a, dtype=np.uint8)
        t[offs[:-1][isa]] = b                     # f(a) = "ba", f(b) = "a"
        s = np.concatenate((np.array([a], np.uint8), t))
    return s[:n].tobytes()


def runs_and_fib(n):
    """Config 5 of BASELINE.json: half single-byte run, half Fibonacci word."""
    h = n // 2
    return b"a" * h + fib(n - h)
