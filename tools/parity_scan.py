import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Ref
from corpus import synth_binary
ref = Ref()
full = synth_binary(200_000_000)
for n, q, w in [(8_000_000, 9, 24), (20_000_000, 9, 24), (40_000_000, 9, 24), (40_000_000, 5, 22), (80_000_000, 9, 24)]:
    d = full[:n]
    res = []
    for rep in range(2):
        got = brotli_b200.compress_oneshot(d, q, w)
        res.append(got)
    want = ref.compress(d, q, w)
    st = brotli_b200.last_stats()
    k = next((i for i in range(min(len(res[0]), len(want))) if res[0][i] != want[i]), -1)
    print(n, q, w, "parity", res[0] == want, res[1] == want, "deterministic", res[0] == res[1], "first diff", k, "iters", st["lz77_iterations"], flush=True)
