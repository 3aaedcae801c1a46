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
        for a_, q_ in zip(at.tolist(), which.tolist()):
            w[a_:a_ + plen[q_]] = pwords[poff[q_]:poff[q_ + 1]]
        s = srng.choice(len(seps), size=m, p=sp)
        parts = [None] * (2 * m)
        parts[0::2] = [words[i] for i in w]
        parts[1::2] = [seps[i] for i in s]
        blob = b"".join(parts)
        out.append(blob)
        got += len(blob)
    return b"".join(out)[:n]


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
