"""One BrotliB200CompressBatch call over K x SIZE-byte slices of the web mix (no reference run, no checks): the thing to put
under ncu for a launch list of the batched quality-5..9 path.  usage: prof_batch.py [count] [quality] [size] [reps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from corpus import synth_web
count = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 5
size = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
total = 60_000_000
src = synth_web(total)
streams = [src[o:o + size] for o in [(i * 104729) % (total - size) for i in range(count)]]
L = brotli_b200.lib()
bufs = [C.create_string_buffer(s, len(s)) for s in streams]
sizes = (C.c_size_t * count)(*[len(s) for s in streams])
caps = [L.BrotliEncoderMaxCompressedSize(len(s)) + 16 for s in streams]
outs = [C.create_string_buffer(c) for c in caps]
in_ptrs = (C.c_void_p * count)(*[C.addressof(b) for b in bufs])
out_ptrs = (C.c_void_p * count)(*[C.addressof(b) for b in outs])
for rep in range(reps):
    out_sizes = (C.c_size_t * count)(*caps)
    t = time.time(); good = L.BrotliB200CompressBatch(q, 22, count, in_ptrs, sizes, out_ptrs, out_sizes, 16); dt = time.time() - t
    st = brotli_b200.last_stats()
    print("run %d: %.3fs = %.1f MB/s, ok %d | %s" % (rep, dt, count * size / dt / 1e6, good, {k: round(v, 2) for k, v in st.items()}), flush=True)
