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
REGRESSIONS = [(1, 304), (4, 382), (7, 249), (7, 13), (6, 104), (5, 252), (5, 297)]
