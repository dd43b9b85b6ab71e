"""BASELINE.json configs 3 and 4 at full size on one GPU: parity against the reference run on the box."""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Ref
from corpus import synth_binary, synth_web
ref = Ref()
which = sys.argv[1:] or ["c4", "c3"]
for w in which:
    t = time.time()
    if w == "c4":
        d, q, lw = synth_binary(200_000_000), 9, 24
    elif w == "c3":
        d, q, lw = synth_web(1 << 30), 5, 22
    else:
        d, q, lw = synth_web(int(w)), 5, 22
    print(w, "generated", len(d), "in %.1fs" % (time.time() - t), flush=True)
    t = time.time(); got = brotli_b200.compress_oneshot(d, q, lw); t_gpu = time.time() - t
    st = brotli_b200.last_stats()
    print(w, "gpu: %d bytes, wall %.2fs, gpu %.1f ms [index %.1f lz77 %.1f (walk %.1f) entropy %.1f] iters %d runs %d/%d mbs %d" % (
        len(got), t_gpu, st["ms_total"], st["ms_index"], st["ms_lz77"], st["ms_walk"], st["ms_entropy"], st["lz77_iterations"],
        st["block_runs"], st["blocks"], st["metablocks"]), flush=True)
    print(w, "sha256 of the GPU stream:", hashlib.sha256(got).hexdigest(), flush=True)
    if os.environ.get("SKIP_REF"):
        print(w, "matches the given digest:", hashlib.sha256(got).hexdigest() == os.environ["SKIP_REF"], flush=True)
        continue
    t = time.time(); want = ref.compress(d, q, lw); t_cpu = time.time() - t
    print(w, "ref: %d bytes in %.2fs (%.1f MB/s); parity %s; gpu %.1f MB/s" % (
        len(want), t_cpu, len(d) / t_cpu / 1e6, got == want, len(d) / (st["ms_total"] / 1e3) / 1e6), flush=True)
