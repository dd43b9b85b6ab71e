"""Randomized campaign in the CPU sim, part 2: (a) batches of independent streams as ONE device job (cuts of kind 3) against the
oracle per stream, (b) one-shot streams with BROTLI_PARAM_LGBLOCK / DISABLE_LITERAL_CONTEXT_MODELING and single FLUSHes against the
reference's CompressStream with the same parameters.  TEST TOOLING (like tests/).

usage: python tools/sim_campaign_batch.py <seed> <seconds> <qualities, e.g. 2,3,4 or 5,6,7,8,9>"""
import sys, os, ctypes as C, time, random
ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT_, 'tests')); sys.path.insert(0, ROOT_)
import numpy as np
from brotli_libs import ROOT, TABLES, Oracle, Ref, ref_stream_ops, zeroed_malloc
zeroed_malloc()
from corpus import synth_text, synth_binary, synth_web
L = C.CDLL(os.path.join(ROOT, "tests/sim/libbrsim.so"))
L.sim_init.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
L.sim_compress_multi.restype = C.c_long
L.sim_compress_multi.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
L.sim_compress_cuts.restype = C.c_long
L.sim_compress_cuts.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
blob = open(TABLES, "rb").read(); L.sim_init(blob, len(blob), (1 << 22) + 2)
seed = int(sys.argv[1]); tlimit = float(sys.argv[2]); qs = [int(x) for x in sys.argv[3].split(',')]
rnd = random.Random(seed); nrnd = np.random.default_rng(seed)
pool = [synth_text(800_000, seed), synth_web(800_000, seed + 1), synth_binary(800_000, seed + 2),
        nrnd.integers(0, 256, 200000, dtype=np.uint8).tobytes(), bytes(200000), bytes(range(256)) * 600,
        nrnd.integers(0, 4, 200000, dtype=np.uint8).tobytes()]
ora = Oracle(); ref = Ref()
def piece(maxn):
    src = rnd.choice(pool); n = max(1, int(min(len(src), maxn) * rnd.uniform(0.01, 1.0)))
    if rnd.random() < 0.25: n = rnd.choice([1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 100])
    o = rnd.randint(0, len(src) - n)
    return src[o:o + n]
t0 = time.time(); cases = bad = 0
while time.time() - t0 < tlimit:
    q = rnd.choice(qs); w = rnd.randint(10 if q < 5 else 17, 24)
    if rnd.random() < 0.6:
        k = rnd.randint(2, 30)
        ss = [piece(rnd.choice([3000, 70000, 70000, 300000])) for _ in range(k)]
        d = b"".join(ss); bounds = np.cumsum([len(x) for x in ss]).astype(np.uint32); ends = np.zeros(k + 1, np.uint64)
        cap = len(d) + len(d) // 2 + 4096 + 64 * k; out = C.create_string_buffer(cap); st = np.zeros(8, np.uint32)
        os.environ["BR_SIM_BATCH_CHUNK_BITS"] = rnd.choice(["9", "11"])
        r = L.sim_compress_multi(q, w, d, len(d), bounds.ctypes.data, k, ends.ctypes.data, out, cap, st.ctypes.data)
        ok = r >= 0 and int(ends[k - 1]) == r
        a = 0
        for j in range(k):
            if not ok: break
            ok = out.raw[a:int(ends[j])] == ora.compress(ss[j], q, w); a = int(ends[j])
        kind = "batch of %d (%s)" % (k, [len(x) for x in ss])
    else:
        if q < 4: continue
        d = piece(900000) + piece(300000); n = len(d)
        lgb = rnd.choice([0, 0, 16, 17, 19, 21, 24]); dis = rnd.choice([0, 1]) if q >= 5 else 0
        prm = {}
        if lgb: prm[3] = lgb
        if dis: prm[4] = 1
        fl = [rnd.randint(1, n - 1)] if (n > 2 and rnd.random() < 0.5) else []
        sizes = ([fl[0], n - fl[0]] if fl else [n]); ops = ([1, 2] if fl else [2])
        want = ref_stream_ops(ref, d, q, w, sizes, ops, params=prm)
        cp = (C.c_uint32 * max(1, len(fl)))(*fl); ck = (C.c_uint32 * max(1, len(fl)))(*([1] * len(fl))); eb = (C.c_uint64 * max(1, len(fl)))()
        cap = n + n // 2 + 4096; out = C.create_string_buffer(cap); st = np.zeros(8, np.uint32)
        r = L.sim_compress_cuts(q, w, sizes[0], d, n, cp, ck, len(fl), 1, 1, 0, eb, out, cap, st.ctypes.data, lgb, dis, 0)
        ok = r >= 0 and out.raw[:r] == want
        kind = "params lgblock %d disable_ctx %d flush %s n %d" % (lgb, dis, fl, n)
    cases += 1
    if not ok:
        bad += 1
        print("MISMATCH seed %d case %d: %s q=%d w=%d" % (seed, cases, kind, q, w), flush=True)
print("batch campaign seed %d q=%s: %d cases, %d bad, %.0fs" % (seed, qs, cases, bad, time.time() - t0), flush=True)
