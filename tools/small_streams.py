"""Many small independent streams at quality 5..9 (the "web payload" case): K x 64 KiB slices of the C3 mix through
BrotliB200CompressBatch (streams below 1 MiB run as groups, one device job per group: br_api.cc compress_stream_group),
bit-exact per stream, beside the reference with one stream per host core.  Times the C call only (host buffers in, host
buffers out).  usage: small_streams.py [count] [quality] [size] [threads]"""
import ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Ref
from corpus import synth_web
count = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 5
size = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
nt = int(sys.argv[4]) if len(sys.argv) > 4 else 16
total = 200_000_000
src = synth_web(total)
streams = [src[o:o + size] for o in [(i * 104729) % (total - size) for i in range(count)]]
nbytes = sum(len(s) for s in streams)
ref = Ref()
ncpu = os.cpu_count() or 1
want = [None] * count
def work(k):
    for i in range(k, count, ncpu):
        want[i] = ref.compress(streams[i], q, 22)
t = time.time(); th = [threading.Thread(target=work, args=(k,)) for k in range(ncpu)]; [x.start() for x in th]; [x.join() for x in th]; t_ref = time.time() - t
print("reference, %d x %d bytes at quality %d on %d host cores: %.3fs = %.1f MB/s" % (count, size, q, ncpu, t_ref, nbytes / t_ref / 1e6), flush=True)
L = brotli_b200.lib()
bufs = [C.create_string_buffer(s, len(s)) for s in streams]
sizes = (C.c_size_t * count)(*[len(s) for s in streams])
caps = [L.BrotliEncoderMaxCompressedSize(len(s)) + 16 for s in streams]
outs = [C.create_string_buffer(c) for c in caps]
in_ptrs = (C.c_void_p * count)(*[C.addressof(b) for b in bufs])
out_ptrs = (C.c_void_p * count)(*[C.addressof(b) for b in outs])
for rep in range(3):
    out_sizes = (C.c_size_t * count)(*caps)
    t = time.time(); good = L.BrotliB200CompressBatch(q, 22, count, in_ptrs, sizes, out_ptrs, out_sizes, nt); dt = time.time() - t
    bad = sum(1 for i in range(count) if outs[i].raw[:out_sizes[i]] != want[i])
    st = brotli_b200.last_stats()
    print("GPU batch (run %d): %.3fs = %.1f MB/s, ok %d of %d, streams differing: %d | last group: %s" % (
        rep, dt, nbytes / dt / 1e6, good, count, bad, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}), flush=True)
