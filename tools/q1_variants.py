"""Tuning runs of the quality-1 batch (config-5 shape on a smaller source): warps per SM x first probe width."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from corpus import synth_web
total, count = 200_000_000, 10000
src = synth_web(total)
streams = [src[o:o + 65536] for o in [(i * 104729) % (total - 65536) for i in range(count)]]
nbytes = sum(len(s) for s in streams)
libs = [a for a in sys.argv[1:] if a.endswith(".so")] or ["libbrotlienc_b200.so"]
variants = [(48, 32, 1), (48, 32, 0)] if "--one" in sys.argv else [(48, 32, 1), (8, 32, 1), (16, 32, 1), (48, 32, 0), (48, 8, 0), (48, 16, 0)]
base = None
def run(w, f, k):
    global base
    for rep in range(2):
        got = brotli_b200.compress_batch(streams, 1, 22, threads=16)
    st = brotli_b200.last_stats_q1()
    if base is None: base = got
    print("kernel %s warps/SM %2d first width %2d: parse %.1f ms code %.1f ms total %.1f ms  same bytes as first variant: %s" % (
        "global " if k == 1 else "on-chip", w, f, st["ms_parse"], st["ms_code"], st["ms_total"], got == base), flush=True)
for so in libs:
    brotli_b200._lib = None
    brotli_b200.LIB_PATH = os.path.join(ROOT, "brotli_b200", so)
    print(so, flush=True)
    for w, f, k in variants:
        os.environ["BR_Q1_WARPS_PER_SM"] = str(w); os.environ["BR_Q1_FIRST_WIDTH"] = str(f); os.environ["BR_Q1_KERNEL"] = str(k)
        t = threading.Thread(target=run, args=(w, f, k)); t.start(); t.join()
