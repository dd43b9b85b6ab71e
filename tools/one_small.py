import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import brotli_b200
from corpus import synth_web
src = synth_web(8_000_000)
streams = [src[o:o + 65536] for o in [(i * 104729) % (len(src) - 65536) for i in range(60)]]
for s in streams[:5]: brotli_b200.compress_oneshot(s, 5, 22)
t = time.time()
for s in streams: brotli_b200.compress_oneshot(s, 5, 22)
dt = time.time() - t
st = brotli_b200.last_stats()
print("wall per stream %.2f ms; gpu events: total %.2f index %.2f lz77 %.2f (walk %.2f, launches %d) entropy %.2f assemble %.2f; kernel launches %d" % (
    1e3 * dt / len(streams), st["ms_total"], st["ms_index"], st["ms_lz77"], st["ms_walk"], st["lz77_iterations"], st["ms_entropy"], st["ms_assemble"], st["launches"]))
for n in (4096, 1 << 20):
    s = src[:n]
    brotli_b200.compress_oneshot(s, 5, 22)
    t = time.time(); [brotli_b200.compress_oneshot(s, 5, 22) for _ in range(10)]; dt = (time.time() - t) / 10
    st = brotli_b200.last_stats()
    print("n=%d: wall %.2f ms; gpu total %.2f index %.2f lz77 %.2f entropy %.2f" % (n, 1e3 * dt, st["ms_total"], st["ms_index"], st["ms_lz77"], st["ms_entropy"]))
