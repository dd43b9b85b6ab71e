"""Structured random inputs (tests/fuzz_cases.py) through the C ABI on the GPU against the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Oracle
from fuzz_cases import cases
ora = Oracle()
seed, count = int(sys.argv[1]), int(sys.argv[2])
bad, t0, nbytes = 0, time.time(), 0
for i, d, q, w in cases(seed, count):
    got = brotli_b200.compress_oneshot(d, q, w)
    nbytes += len(d)
    if got != ora.compress(d, q, w):
        bad += 1
        print("MISMATCH seed %d case %d: %d bytes q%d w%d" % (seed, i, len(d), q, w), flush=True)
print("gpu fuzz seed %d: %d cases, %d bytes, %d mismatches, %.1fs" % (seed, count, nbytes, bad, time.time() - t0), flush=True)
