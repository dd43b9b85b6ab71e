"""One compression of an N-byte synthetic text stream (for ncu)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from corpus import synth_text
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = synth_text(n, seed=3)
for _ in range(reps):
    out = brotli_b200.compress_oneshot(d, 5, 22)
print(len(d), len(out), brotli_b200.last_stats())
