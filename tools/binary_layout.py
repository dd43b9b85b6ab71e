"""Region layout of tests/corpus.py synth_binary(n, seed): kind and byte range of every piece (replays the generator's
random streams without building the bytes)."""
import random, sys
import numpy as np
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "tests"))
def layout(n, seed=20250924):
    rs = np.random.RandomState(seed & 0x7FFFFFFF); rng = random.Random(seed)
    rs.randint(0, 256, size=300000, dtype=np.uint8)
    pats = [bytes(rs.randint(0, 256, size=rng.randint(2, 9), dtype=np.uint8)) for _ in range(64)]
    size, out = 0, []
    while size < n:
        u = rng.random(); chunk = rng.randint(200000, 1500000)
        if u < 0.25: kind = "text"; rng.randint(0, 1 << 30); ln = chunk
        elif u < 0.5:
            kind = "tile"; k = max(1, chunk // 100); rs.randint(0, chunk, size=k); rs.randint(0, 256, size=k, dtype=np.uint8); ln = chunk
        elif u < 0.75: kind = "walk"; rs.normal(0, 50, size=chunk // 4); ln = (chunk // 4) * 4
        elif u < 0.9: kind = "soup"; ln = sum(len(rng.choice(pats)) for _ in range(chunk // 5))
        else: kind = "random"; rs.randint(0, 256, size=chunk, dtype=np.uint8); ln = chunk
        out.append((kind, size, size + ln)); size += ln
    return out
if __name__ == "__main__":
    for kind, a, b in layout(int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000):
        print("%-6s %10d %10d  MiB %6.1f - %6.1f" % (kind, a, b, a / 2**20, b / 2**20))
