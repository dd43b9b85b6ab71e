"""Times alternative builds of the library (tuning experiments)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from corpus import synth_text
import brotli_b200
d = synth_text(100_000_000, seed=3)
for so in sys.argv[1:]:
    brotli_b200._lib = None
    brotli_b200.LIB_PATH = os.path.join(ROOT, "brotli_b200", so)
    for _ in range(3):
        out = brotli_b200.compress_oneshot(d, 5, 22)
    st = brotli_b200.last_stats()
    print(so, len(out), "total %.1f lz77 %.1f walk %.1f iters %d runs %d/%d" % (st["ms_total"], st["ms_lz77"], st["ms_walk"], st["lz77_iterations"], st["block_runs"], st["blocks"]), flush=True)
