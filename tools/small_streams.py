"""Many small independent streams at quality 5 (the "web payload" case): K x 64 KiB slices of the C3 mix through
BrotliB200CompressBatch (host thread pool over the one-stream pipeline), bit-exact per stream, beside the reference with one
stream per host core.  usage: small_streams.py [count] [threads ...]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Ref
from corpus import synth_web
count = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
threads = [int(x) for x in sys.argv[2:]] or [8, 32]
total = 200_000_000
src = synth_web(total)
streams = [src[o:o + 65536] for o in [(i * 104729) % (total - 65536) for i in range(count)]]
nbytes = sum(len(s) for s in streams)
ref = Ref()
ncpu = os.cpu_count() or 1
want = [None] * count
def work(k):
    for i in range(k, count, ncpu):
        want[i] = ref.compress(streams[i], 5, 22)
t = time.time(); th = [threading.Thread(target=work, args=(k,)) for k in range(ncpu)]; [x.start() for x in th]; [x.join() for x in th]; t_ref = time.time() - t
print("reference, %d streams on %d host cores: %.3fs = %.1f MB/s" % (count, ncpu, t_ref, nbytes / t_ref / 1e6), flush=True)
for nt in threads:
    brotli_b200.compress_batch(streams[:64], 5, 22, threads=nt)
    t = time.time(); got = brotli_b200.compress_batch(streams, 5, 22, threads=nt); dt = time.time() - t
    bad = sum(1 for a, b in zip(got, want) if a != b)
    print("GPU, %d host threads: %.3fs = %.1f MB/s (%.2f ms per stream per thread), streams differing: %d" % (
        nt, dt, nbytes / dt / 1e6, 1e3 * dt * nt / count, bad), flush=True)
