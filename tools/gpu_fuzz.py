"""Structured random inputs (tests/fuzz_cases.py) through the C ABI on the GPU against the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Oracle
from fuzz_cases import cases
ora = Oracle()
# usage: gpu_fuzz.py <seed> <count> [low]   -- "low": the same inputs at quality 2..4 with windows 10..24 instead
seed, count = int(sys.argv[1]), int(sys.argv[2])
low = len(sys.argv) > 3 and sys.argv[3] == "low"
bad, t0, nbytes = 0, time.time(), 0
for i, d, q, w in cases(seed, count):
    if low:
        if not d:
            continue
        q, w = 2 + i % 3, 10 + (i * 7) % 15
    got = brotli_b200.compress_oneshot(d, q, w)
    nbytes += len(d)
    if got != ora.compress(d, q, w):
        bad += 1
        print("MISMATCH seed %d case %d: %d bytes q%d w%d" % (seed, i, len(d), q, w), flush=True)
print("gpu fuzz seed %d: %d cases, %d bytes, %d mismatches, %.1fs" % (seed, count, nbytes, bad, time.time() - t0), flush=True)
