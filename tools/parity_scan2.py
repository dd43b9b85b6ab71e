import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Ref
from corpus import synth_binary
ref = Ref()
full = synth_binary(200_000_000)
got = brotli_b200.compress_oneshot(full, 9, 24)
want = ref.compress(full, 9, 24)
k = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), -1)
print("full: first diff at compressed byte", k, "of", len(want), flush=True)
# estimate the input offset: compress growing prefixes with the reference until the size passes k
lo, hi = 0, len(full)
for _ in range(12):
    mid = (lo + hi) // 2
    if len(ref.compress(full[:mid], 9, 24)) < k: lo = mid
    else: hi = mid
print("diff is near input offset", lo, hi, flush=True)
start = max(0, ((lo - 24_000_000) >> 18) << 18)
for a, b in [(start, min(len(full), hi + 6_000_000)), (max(0, start - (8 << 20)), min(len(full), hi + 2_000_000))]:
    d = full[a:b]
    g = brotli_b200.compress_oneshot(d, 9, 24); w = ref.compress(d, 9, 24)
    print("slice", a, b, "parity", g == w, flush=True)
