"""Seeded synthetic corpora shaped like BASELINE.json's configs.  Pure numpy /
stdlib so they regenerate identically here and on the GPU box."""
import random

import numpy as np

_SYL = ["a", "an", "ar", "as", "at", "be", "ca", "co", "da", "de", "di", "do", "el", "en", "er", "es", "fa",
        "fi", "ge", "ha", "he", "hi", "in", "is", "it", "ka", "la", "le", "li", "lo", "ma", "me", "mi", "mo",
        "na", "ne", "no", "nt", "of", "on", "or", "ou", "pa", "pe", "po", "ra", "re", "ri", "ro", "sa", "se",
        "si", "so", "st", "ta", "te", "th", "ti", "to", "tr", "un", "ur", "ve", "vi", "wa", "we", "wi", "yo"]


def _vocab(rng, n):
    words = set()
    while len(words) < n:
        k = 1 + min(int(rng.expovariate(0.7)), 4)
        words.add("".join(rng.choice(_SYL) for _ in range(k)))
    return sorted(words)


def synth_text(n, seed=20250922):
    """enwik8-shaped text: Zipf(1.1) word choice over a 30k vocabulary, punctuation, newlines,
    occasional markup -- C2 of BASELINE.json."""
    rng = random.Random(seed)
    vocab = _vocab(rng, 30000)
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    ranks = np.arange(1, len(vocab) + 1, dtype=np.float64)
    p = 1.0 / ranks ** 1.1
    p /= p.sum()
    out = []
    size = 0
    while size < n:
        k = 20000
        idx = rs.choice(len(vocab), size=k, p=p)
        punct = rs.random_sample(k)
        parts = []
        for i in range(k):
            w = vocab[idx[i]]
            u = punct[i]
            if u < 0.06:
                w = w.capitalize()
            parts.append(w)
            if u > 0.985:
                parts.append(".\n\n" if u > 0.995 else ".\n")
            elif u > 0.93:
                parts.append(". ")
            elif u > 0.86:
                parts.append(", ")
            elif u > 0.85:
                parts.append(" [[%s]] " % vocab[idx[(i * 7) % k]])
            elif u > 0.845:
                parts.append(" %d " % int(u * 1e6))
            else:
                parts.append(" ")
        s = "".join(parts).encode("ascii")
        out.append(s)
        size += len(s)
    return b"".join(out)[:n]


def synth_web(n, seed=20250923):
    """HTML / minified JS / JSON mix in 16-256 KiB documents -- C3 of BASELINE.json."""
    rng = random.Random(seed)
    tags = ["div", "span", "a", "p", "li", "ul", "td", "tr", "table", "h1", "h2", "img", "section", "nav"]
    idents = ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz_$") for _ in range(rng.randint(1, 9))) for _ in range(5000)]
    keys = ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 10))) for _ in range(500)]
    words = _vocab(rng, 8000)
    out = []
    size = 0
    while size < n:
        kind = rng.random()
        target = rng.randint(16 << 10, 256 << 10)
        parts = []
        got = 0
        if kind < 0.4:
            while got < target:
                t = rng.choice(tags)
                s = '<%s class="%s %s" id="%s%d">%s</%s>\n' % (
                    t, rng.choice(idents), rng.choice(idents), rng.choice(keys), rng.randint(0, 999),
                    " ".join(rng.choice(words) for _ in range(rng.randint(1, 14))), t)
                parts.append(s); got += len(s)
        elif kind < 0.7:
            ops = ["=", "+", "-", "*", "===", "&&", "||", "<", ">", "?", ":", "."]
            while got < target:
                s = "function %s(%s,%s){var %s=%s%s%s;return %s(%s)%s%d}" % (
                    rng.choice(idents), rng.choice(idents), rng.choice(idents), rng.choice(idents),
                    rng.choice(idents), rng.choice(ops), rng.choice(idents), rng.choice(idents),
                    rng.choice(idents), rng.choice(ops), rng.randint(0, 65535))
                parts.append(s); got += len(s)
        else:
            while got < target:
                s = '{"%s":%d,"%s":"%s","%s":[%d,%d,%.3f],"%s":%s},' % (
                    rng.choice(keys), rng.randint(0, 10 ** 6), rng.choice(keys), rng.choice(words),
                    rng.choice(keys), rng.randint(0, 255), rng.randint(0, 255), rng.random() * 100,
                    rng.choice(keys), rng.choice(["true", "false", "null"]))
                parts.append(s); got += len(s)
        b = "".join(parts).encode("ascii")
        out.append(b); size += len(b)
    return b"".join(out)[:n]


def synth_binary(n, seed=20250924):
    """Silesia-shaped binary mix: text, mutated repeats, int32 random walks, opcode soup,
    uniform noise -- C4 of BASELINE.json."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    rng = random.Random(seed)
    out = []
    size = 0
    tile = rs.randint(0, 256, size=300000, dtype=np.uint8)
    tile[::3] = (np.arange(100000) % 251).astype(np.uint8)
    pats = [bytes(rs.randint(0, 256, size=rng.randint(2, 9), dtype=np.uint8)) for _ in range(64)]
    while size < n:
        u = rng.random()
        chunk = rng.randint(200000, 1500000)
        if u < 0.25:
            b = synth_text(chunk, seed=rng.randint(0, 1 << 30))
        elif u < 0.5:
            reps = chunk // len(tile) + 1
            a = np.tile(tile, reps)[:chunk].copy()
            k = max(1, chunk // 100)
            a[rs.randint(0, chunk, size=k)] = rs.randint(0, 256, size=k, dtype=np.uint8)
            b = a.tobytes()
        elif u < 0.75:
            w = np.cumsum(rs.normal(0, 50, size=chunk // 4).astype(np.int64)).astype(np.int32)
            b = w.tobytes()
        elif u < 0.9:
            b = b"".join(rng.choice(pats) for _ in range(chunk // 5))
        else:
            b = rs.randint(0, 256, size=chunk, dtype=np.uint8).tobytes()
        out.append(b); size += len(b)
    return b"".join(out)[:n]
