"""A small batch of streams through BrotliB200CompressBatch (quality 5 and 9) checked against the oracle: short enough to
run under compute-sanitizer.  usage: batch_sanity.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Oracle
from corpus import synth_web
ora = Oracle()
src = synth_web(400000, 5)
streams = [src[i * 9973:i * 9973 + 3000 + 2500 * i] for i in range(24)] + [src[:70000], b"x", src[:5]]
for q, w in ((5, 22), (9, 24)):
    got = brotli_b200.compress_batch(streams, q, w, threads=2)
    bad = [k for k, x in enumerate(streams) if got[k] != ora.compress(x, q, w)]
    print("quality %d lgwin %d: %d streams, differing: %s" % (q, w, len(streams), bad), flush=True)
    assert not bad
print("batch sanity ok")
