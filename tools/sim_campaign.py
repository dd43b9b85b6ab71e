"""Randomized campaign in the CPU sim (tests/sim/libbrsim.so = the product's device functions compiled for the host): mixes of
text / web / binary / noise / runs, one-shot against the oracle (70 %) or as a random PROCESS / FLUSH / FINISH sequence against the
reference's own CompressStream (oracle/_ref), with FLUSH positions biased to the neighbourhood of input-block boundaries.
TEST TOOLING (like tests/): it found the lost stitch position behind a too-short block.

usage: python tools/sim_campaign.py <seed> <seconds> <qualities, e.g. 2,3,4 or 5,6,7,8,9>"""
import sys, os, ctypes as C, time, random
ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT_, 'tests')); sys.path.insert(0, ROOT_)
import numpy as np
from brotli_libs import ROOT, TABLES, Oracle, Ref, ref_stream_ops, zeroed_malloc
zeroed_malloc()
from corpus import synth_text, synth_binary, synth_web
L = C.CDLL(os.path.join(ROOT, "tests/sim/libbrsim.so"))
L.sim_init.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
L.sim_compress.restype = C.c_long
L.sim_compress.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
L.sim_compress_cuts.restype = C.c_long
L.sim_compress_cuts.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
blob = open(TABLES, "rb").read(); L.sim_init(blob, len(blob), (1 << 22) + 2)
seed = int(sys.argv[1]); tlimit = float(sys.argv[2]); qs = [int(x) for x in sys.argv[3].split(',')]
rnd = random.Random(seed); nrnd = np.random.default_rng(seed)
pool = [synth_text(1_500_000, seed), synth_web(1_500_000, seed + 1), synth_binary(1_500_000, seed + 2),
        nrnd.integers(0, 256, 400000, dtype=np.uint8).tobytes(), bytes(300000), bytes(range(256)) * 1000,
        nrnd.integers(0, 4, 300000, dtype=np.uint8).tobytes()]
ora = Oracle(); ref = Ref()
def mix():
    parts = []
    for _ in range(rnd.randint(1, 6)):
        src = rnd.choice(pool); n = int(min(len(src), rnd.choice([10, 300, 5000, 70000, 300000, 900000]) * rnd.uniform(0.3, 1.0)) ) or 1
        o = rnd.randint(0, len(src) - n)
        parts.append(src[o:o + n])
    if rnd.random() < 0.3: parts.append(parts[0])
    return b"".join(parts)
t0 = time.time(); cases = bad = 0
while time.time() - t0 < tlimit:
    d = mix(); q = rnd.choice(qs); w = rnd.randint(10 if q < 5 else 17, 24)
    if rnd.random() < 0.7:
        want = ora.compress(d, q, w)
        cap = len(d) + len(d)//2 + 4096
        out = C.create_string_buffer(cap); st = np.zeros(8, np.uint32)
        r = L.sim_compress(q, w, d, len(d), out, cap, st.ctypes.data)
        ok = r >= 0 and out.raw[:r] == want
        kind = "oneshot"
    else:
        # a random PROCESS / FLUSH sequence, against the reference's CompressStream
        n = len(d); k = rnd.randint(1, 4)
        def cutpos():
            if rnd.random() < 0.6 and n > 40000:
                b = 1 << rnd.choice([14, 16, 18]); m = rnd.randint(1, max(1, (n - 1) // b)); return min(n - 1, max(1, m * b + rnd.choice([-9, -8, -7, -4, -3, -2, -1, 0, 1, 2, 3, 4, 6, 7, 8, 9])))
            return rnd.randint(1, max(1, n - 1))
        cuts = sorted(set(cutpos() for _ in range(k)))
        sizes, ops, prev = [], [], 0
        for c in cuts:
            sizes.append(c - prev); ops.append(rnd.choice([0, 1])); prev = c
        sizes.append(n - prev); ops.append(2)
        want = ref_stream_ops(ref, d, q, w, sizes, ops)
        bs = 1 << (14 if q < 4 else (18 if q >= 9 and w >= 18 else (w if q >= 9 else 16)))
        pos, fl, acc, fixed, hint = 0, [], 0, False, 0
        for a, op in zip(sizes, ops):
            if not fixed and (op != 0 or acc + a >= bs):
                hint, fixed = acc + a, True
            acc += a; pos += a
            if op == 1 and pos > 0 and (not fl or fl[-1] != pos): fl.append(pos)
        cp = (C.c_uint32 * max(1, len(fl)))(*fl); ck = (C.c_uint32 * max(1, len(fl)))(*([1] * len(fl)))
        eb = (C.c_uint64 * max(1, len(fl)))()
        cap = n + n // 2 + 4096 + 64 * len(fl)
        out = C.create_string_buffer(cap); st = np.zeros(8, np.uint32)
        fin = 0 if (fl and fl[-1] == n) else 1
        r = L.sim_compress_cuts(q, w, hint, d, n, cp, ck, len(fl), fin, 1, 0, eb, out, cap, st.ctypes.data, 0, 0, 0)
        got = out.raw[:max(r, 0)] + (b"\x03" if not fin else b"")
        ok = r >= 0 and got == want
        kind = "stream %s %s" % (sizes, ops)
    cases += 1
    if not ok:
        bad += 1
        print("MISMATCH seed %d case %d: %s n=%d q=%d w=%d" % (seed, cases, kind, len(d), q, w), flush=True)
print("campaign seed %d q=%s: %d cases, %d bad, %.0fs" % (seed, qs, cases, bad, time.time() - t0), flush=True)
