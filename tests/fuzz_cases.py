"""Structured random inputs (runs, self-copies at all distances, small alphabets, dictionary words) that
exercise the corners seeded corpora miss: copies running to / over input-block ends, ExtendLastCommand over
many chunks, heavy overlap of speculative walker ranges.  Deterministic: case (seed, index) regenerates."""
import random

WORDS = [b"the ", b"and ", b"http://", b"<div class=", b"function", b"return ", b"</a>", b" of the "]


def _gen(rnd):
    n = rnd.choice([rnd.randint(0, 64), rnd.randint(0, 3000), rnd.randint(0, 70000), rnd.randint(60000, 300000)])
    kind = rnd.randint(0, 5)
    out = bytearray()
    alpha = rnd.choice([2, 4, 16, 64, 256])
    while len(out) < n:
        r = rnd.random()
        if kind == 0 or r < 0.3:
            out += bytes(rnd.randrange(alpha) for _ in range(rnd.randint(1, 50)))
        elif r < 0.6 and len(out) > 4:
            d = rnd.randint(1, min(len(out), rnd.choice([4, 64, 5000, 300000])))
            length = rnd.randint(2, rnd.choice([8, 40, 400, 5000]))
            for _ in range(length):
                out.append(out[-d])
        elif r < 0.8:
            out += bytes([rnd.randrange(alpha)]) * rnd.randint(1, rnd.choice([10, 300, 20000]))
        else:
            out += rnd.choice(WORDS)
    return bytes(out[:n])


def cases(seed, count):
    """Yields (index, data, quality, lgwin)."""
    rnd = random.Random(seed)
    for i in range(count):
        d = _gen(rnd)
        q = rnd.choice([1, 1, 5, 5, 6, 9])
        w = rnd.choice([10, 14, 18, 22]) if q == 1 else rnd.choice([17, 18, 20, 22, 24])
        yield i, d, q, w


# (seed, index): inputs that exposed real bugs in the speculative parse (stitch-bit ownership when a copy runs
# to the block end; stale overlapping walker ranges behind ExtendLastCommand)
REGRESSIONS = [(1, 304), (4, 382), (7, 249), (7, 13), (6, 104), (5, 252), (5, 297),
               (35, 22)]   # (35, 22): a view that reaches through the chunk's own unstored positions into a range flipped in the same launch


def dict_cases(seed, count, tables_path):
    """Texts made of words of the static dictionary (whole, tail-cut, upper-cased) between noise: exercises the
    dictionary search, its cutoff transforms and the lookup gate (hash.h:140-202).  Yields (index, data, q, lgwin)."""
    import struct
    blob = open(tables_path, "rb").read()           # layout: oracle/gen_tables.c
    size_bits = list(blob[8:40]); offsets = struct.unpack("<32I", blob[40:168]); dic = blob[168:168 + 122784]
    rnd = random.Random(seed)

    def word():
        n = rnd.randint(4, 24)
        i = rnd.randrange(1 << size_bits[n])
        return dic[offsets[n] + n * i: offsets[n] + n * i + n]

    for idx in range(count):
        n = rnd.choice([rnd.randint(100, 5000), rnd.randint(5000, 100000)])
        out = bytearray()
        p_word = rnd.choice([0.9, 0.5, 0.1, 0.02])
        sep = rnd.choice([b" ", b"", b"\n", None])
        while len(out) < n:
            r = rnd.random()
            if r < p_word:
                w = word()
                t = rnd.random()
                if t < 0.2:
                    w = w[:max(1, len(w) - rnd.randint(1, 9))]
                elif t < 0.3:
                    w = w.upper()
                out += w
                out += sep if sep is not None else bytes([rnd.randrange(256)])
            elif r < p_word + 0.05 and len(out) > 10:
                d = rnd.randint(1, len(out))
                for _ in range(rnd.randint(3, 60)):
                    out.append(out[-d])
            else:
                out += bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 30)))
        yield idx, bytes(out[:n]), rnd.choice([5, 5, 6, 7, 9]), rnd.choice([17, 18, 22, 24])
