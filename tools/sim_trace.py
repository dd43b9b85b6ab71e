"""Run the LZ77 fixpoint of the CPU sim (tests/sim) on a synthetic input and print the per-launch trace.
usage: sim_trace.py <binary|text|web|walk|random> <bytes> <quality> <lgwin> [seed]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from brotli_libs import TABLES, Oracle
from corpus import synth_binary, synth_text, synth_web
kind, n, q, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 20250924
if kind == "binary": d = synth_binary(n, seed)
elif kind == "text": d = synth_text(n, seed)
elif kind == "web": d = synth_web(n, seed)
elif kind == "walk":
    rs = np.random.RandomState(seed & 0x7fffffff)
    d = np.cumsum(rs.normal(0, 50, size=n // 4).astype(np.int64)).astype(np.int32).tobytes()
elif kind == "soup":   # the opcode soup of synth_binary: 64 patterns of 2-9 bytes in random order
    import random
    rs = np.random.RandomState(seed & 0x7fffffff); rng = random.Random(seed)
    pats = [bytes(rs.randint(0, 256, size=rng.randint(2, 9), dtype=np.uint8)) for _ in range(64)]
    d = b"".join(rng.choice(pats) for _ in range(n // 5))[:n]
else:
    d = np.random.RandomState(seed & 0x7fffffff).randint(0, 256, size=n, dtype=np.uint8).tobytes()
L = C.CDLL(os.environ.get("SIM_SO", os.path.join(ROOT, "tests", "sim", "libbrsim.so")))
L.sim_init.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
blob = open(TABLES, "rb").read(); L.sim_init(blob, len(blob), (1 << 22) + 2)
L.sim_compress.restype = C.c_long
L.sim_compress.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
cap = len(d) + len(d) // 2 + 4096
out = C.create_string_buffer(cap); st = np.zeros(8, np.uint32)
t = time.time(); r = L.sim_compress(q, w, d, len(d), out, cap, st.ctypes.data); t = time.time() - t
want = Oracle().compress(d, q, w)
print("sim: %d bytes, %d launches, %d chunk walks for %d chunks, rounds %d, serial-cost model %d chunk-times, %.1fs; parity %s" % (
    r, st[0], st[1], st[2], st[3], st[4], t, out.raw[:r] == want))
