"""Seeded inputs shared by the golden-vector generator and the tests."""
import numpy as np

from corpus import synth_binary, synth_text, synth_web


def make_case(c):
    k, n, seed = c["kind"], c["n"], c["seed"]
    if k == "text":
        return synth_text(n, seed)
    if k == "web":
        return synth_web(n, seed)
    if k == "binary":
        return synth_binary(n, seed)
    if k == "zeros":
        return bytes(n)
    if k == "random":
        return np.random.RandomState(seed).randint(0, 256, n, dtype=np.uint8).tobytes()
    if k == "periodic":   # long matches crossing input blocks: exercises ExtendLastCommand
        base = synth_text(c.get("period", 70001), seed)
        return (base * (n // len(base) + 1))[:n]
    if k == "heavy":      # one 5-byte key inserted > 65536 times: the uint16 bucket counter wraps
        rs = np.random.RandomState(seed)
        a = np.zeros((n // 8, 8), np.uint8)
        a[:, :5] = np.frombuffer(b"abcde", np.uint8)
        a[:, 5:] = rs.randint(0, 256, (n // 8, 3), dtype=np.uint8)
        return a.tobytes()
    if k == "mixed":
        parts = [synth_text(n // 4, seed), np.random.RandomState(seed).randint(0, 256, n // 4, dtype=np.uint8).tobytes(),
                 synth_web(n // 4, seed + 1), bytes(n // 8), synth_binary(n - 3 * (n // 4) - n // 8, seed + 2)]
        return b"".join(parts)
    raise ValueError(k)


CASES = [
    dict(kind="text", n=1, seed=1, q=5, lgwin=22),
    dict(kind="text", n=2, seed=1, q=5, lgwin=22),
    dict(kind="text", n=7, seed=1, q=9, lgwin=24),
    dict(kind="text", n=64, seed=2, q=5, lgwin=22),
    dict(kind="text", n=1000, seed=3, q=6, lgwin=20),
    dict(kind="text", n=65536, seed=4, q=5, lgwin=22),       # BASELINE config C1 shape
    dict(kind="text", n=65537, seed=4, q=7, lgwin=18),
    dict(kind="text", n=300000, seed=5, q=5, lgwin=22),
    dict(kind="text", n=300000, seed=5, q=9, lgwin=24),
    dict(kind="text", n=300000, seed=5, q=8, lgwin=17),
    dict(kind="text", n=1500000, seed=6, q=5, lgwin=22),     # >= 1 MiB: H68 (5-byte hash)
    dict(kind="text", n=1500000, seed=6, q=9, lgwin=24),     # H6, 256-deep buckets
    dict(kind="text", n=1048576, seed=7, q=6, lgwin=19),
    dict(kind="web", n=1200000, seed=8, q=5, lgwin=22),
    dict(kind="web", n=400000, seed=8, q=7, lgwin=22),
    dict(kind="binary", n=1500000, seed=9, q=5, lgwin=22),
    dict(kind="binary", n=1500000, seed=9, q=9, lgwin=24),
    dict(kind="zeros", n=400000, seed=0, q=5, lgwin=22),
    dict(kind="zeros", n=1300000, seed=0, q=9, lgwin=24),
    dict(kind="random", n=300000, seed=10, q=5, lgwin=22),   # uncompressed metablocks
    dict(kind="random", n=1200000, seed=10, q=6, lgwin=22),
    dict(kind="periodic", n=1300000, seed=11, q=5, lgwin=22, period=70001),
    dict(kind="periodic", n=700000, seed=11, q=9, lgwin=18, period=300007),
    dict(kind="heavy", n=1280000, seed=12, q=5, lgwin=22),
    dict(kind="heavy", n=1280000, seed=12, q=9, lgwin=24),
    dict(kind="mixed", n=2500000, seed=13, q=5, lgwin=22),
    dict(kind="mixed", n=2500000, seed=13, q=8, lgwin=24),
    dict(kind="text", n=9500000, seed=14, q=5, lgwin=22),    # > 8 MiB ring buffer of the reference: wrap rules
    dict(kind="text", n=5000000, seed=15, q=5, lgwin=17),    # small window: ring wraps many times
    # quality 1: the two-pass fragment coder
    dict(kind="text", n=1, seed=1, q=1, lgwin=22),
    dict(kind="text", n=15, seed=1, q=1, lgwin=22),           # below the 16-byte input margin: literals only
    dict(kind="text", n=16, seed=1, q=1, lgwin=22),
    dict(kind="text", n=17, seed=1, q=1, lgwin=22),
    dict(kind="text", n=1000, seed=3, q=1, lgwin=16),         # 2^10-entry table, min_match 4
    dict(kind="web", n=65536, seed=16, q=1, lgwin=22),        # BASELINE config C5 shape: 2^16 table, min_match 6
    dict(kind="text", n=300000, seed=5, q=1, lgwin=22),       # three 128 KiB blocks sharing one table
    dict(kind="text", n=300000, seed=5, q=1, lgwin=10),       # 1 KiB fragments
    dict(kind="web", n=1200000, seed=8, q=1, lgwin=18),       # 256 KiB fragments
    dict(kind="binary", n=1500000, seed=9, q=1, lgwin=22),
    dict(kind="zeros", n=400000, seed=0, q=1, lgwin=22),
    dict(kind="random", n=300000, seed=10, q=1, lgwin=22),    # raw meta-blocks (ShouldCompress)
    dict(kind="random", n=300000, seed=10, q=1, lgwin=10),    # larger than MaxCompressedSize: raw stream
    dict(kind="mixed", n=2500000, seed=13, q=1, lgwin=22),
    dict(kind="text", n=5000000, seed=15, q=1, lgwin=20),     # five fragments
]


# Qualities 2..4: restated in the oracle as groundwork, not built on the GPU yet.
# Pinned in tests/golden/golden_oracle_only.json; only tests/test_oracle.py reads them.
CASES_ORACLE_ONLY = [
    dict(kind="text", n=1, seed=1, q=2, lgwin=22),
    dict(kind="text", n=9, seed=1, q=3, lgwin=22),
    dict(kind="text", n=1000, seed=3, q=4, lgwin=16),
    dict(kind="web", n=65536, seed=16, q=2, lgwin=22),        # <= 128 commands per meta-block: static command code
    dict(kind="web", n=65536, seed=16, q=3, lgwin=22),
    dict(kind="web", n=65536, seed=16, q=4, lgwin=22),
    dict(kind="text", n=300000, seed=5, q=2, lgwin=18),
    dict(kind="text", n=300000, seed=5, q=3, lgwin=10),
    dict(kind="text", n=300000, seed=5, q=4, lgwin=22),       # H4
    dict(kind="text", n=1500000, seed=6, q=4, lgwin=22),      # H54 (>= 1 MiB)
    dict(kind="binary", n=1500000, seed=9, q=2, lgwin=22),
    dict(kind="binary", n=1500000, seed=9, q=3, lgwin=24),
    dict(kind="zeros", n=400000, seed=0, q=2, lgwin=22),
    dict(kind="random", n=300000, seed=10, q=3, lgwin=22),
    dict(kind="mixed", n=2500000, seed=13, q=4, lgwin=20),
]
