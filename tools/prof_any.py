"""One compression of a synthetic stream (for ncu): prof_any.py <text|web|binary> <bytes> <quality> <lgwin> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from corpus import synth_binary, synth_text, synth_web
kind, n, q, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
d = {"binary": synth_binary, "text": synth_text, "web": synth_web}[kind](n)
for _ in range(reps):
    out = brotli_b200.compress_oneshot(d, q, w)
print(len(d), len(out), brotli_b200.last_stats())
