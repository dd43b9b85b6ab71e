"""BASELINE.json config 5 at full size on one GPU: 10 000 x 64 KiB streams (slices of the C3 web mix),
quality 1, lgwin 22, one device batch; parity of every stream against the reference run on the box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Ref
from corpus import synth_web
count = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
t = time.time(); src = synth_web(total); print("generated %d bytes in %.1fs" % (len(src), time.time() - t), flush=True)
offs = [(i * 104729) % (total - 65536) for i in range(count)]
streams = [src[o:o + 65536] for o in offs]
nbytes = sum(len(s) for s in streams)
for rep in range(3):
    t = time.time(); got = brotli_b200.compress_batch(streams, 1, 22, threads=16); wall = time.time() - t
    st = brotli_b200.last_stats_q1()
    print("gpu rep %d: wall %.3fs (incl. ctypes marshalling)  pipeline %.1f ms [h2d %.1f parse %.1f code %.1f pack %.1f d2h %.1f]  "
          "%.0f MB/s pipeline, %.0f MB/s kernels only; out %d bytes" % (
              rep, wall, st["ms_total"], st["ms_h2d"], st["ms_parse"], st["ms_code"], st["ms_pack"], st["ms_d2h"],
              nbytes / st["ms_total"] / 1e3, nbytes / (st["ms_parse"] + st["ms_code"] + st["ms_pack"]) / 1e3, st["out_bytes"]), flush=True)
ref = Ref()
t = time.time(); want = [ref.compress(s, 1, 22) for s in streams]; t_cpu = time.time() - t
bad = sum(1 for a, b in zip(got, want) if a != b)
print("reference: %.2fs on 1 core (%.1f MB/s); streams differing: %d of %d; parity %s" % (
    t_cpu, nbytes / t_cpu / 1e6, bad, count, bad == 0), flush=True)
# the reference with one stream per host core (ctypes releases the GIL)
import threading
ncpu = os.cpu_count() or 1
def work(k):
    for i in range(k, count, ncpu):
        ref.compress(streams[i], 1, 22)
t = time.time()
th = [threading.Thread(target=work, args=(k,)) for k in range(ncpu)]
[x.start() for x in th]; [x.join() for x in th]
t_all = time.time() - t
print("reference on all %d host cores: %.2fs (%.1f MB/s)" % (ncpu, t_all, nbytes / t_all / 1e6), flush=True)
