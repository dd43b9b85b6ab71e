import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
if os.environ.get("BR_LIB"):   # an alternative build of the library (e.g. the -DBR_DEBUG_KNOBS one: BR_TRACE=1 prints every launch)
    brotli_b200.LIB_PATH = os.path.join(ROOT, "brotli_b200", os.environ["BR_LIB"])
from corpus import synth_binary, synth_text, synth_web
kind, n, q, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
d = {"binary": synth_binary, "text": synth_text, "web": synth_web}[kind](n)
out = brotli_b200.compress_oneshot(d, q, w)
st = brotli_b200.last_stats()
print(kind, n, q, w, len(out), "total %.1f ms lz77 %.1f walk %.1f iters %d runs %d/%d" % (st["ms_total"], st["ms_lz77"], st["ms_walk"], st["lz77_iterations"], st["block_runs"], st["blocks"]))
