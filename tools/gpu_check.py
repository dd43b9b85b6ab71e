"""First-contact GPU script: index sort check, parity against oracle/_ref, timings."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_b200
from brotli_libs import Oracle, Ref
from corpus import synth_binary, synth_text, synth_web

L = brotli_b200.lib()
print("available", brotli_b200.available(), flush=True)
ref = Ref()
ora = Oracle()


def check_sort(d, q, w):
    n = len(d)
    L.br_debug_sort.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
    S = np.zeros(n, np.uint32)
    seg = np.zeros(40000, np.uint32)
    ok = L.br_debug_sort(q, w, d, n, S.ctypes.data, seg.ctypes.data)
    assert ok
    # reference: stable sort of positions by key
    hash64 = n >= (1 << 20) and w >= 19
    a = np.frombuffer(d + bytes(16), np.uint8)
    if hash64:
        v = np.zeros(n, np.uint64)
        for i in range(8):
            v |= a[i:i + n].astype(np.uint64) << np.uint64(8 * i)
        mul = np.uint64((0x1FE35A7BD3579BD3 << 24) & 0xFFFFFFFFFFFFFFFF)
        key = ((v * mul) >> np.uint64(49)).astype(np.uint32)
        htl, nb = 8, 1 << 15
    else:
        v = np.zeros(n, np.uint32)
        for i in range(4):
            v |= a[i:i + n].astype(np.uint32) << np.uint32(8 * i)
        bb = 14 if q < 7 else 15
        key = ((v * np.uint32(0x1E35A7BD)) >> np.uint32(32 - bb)).astype(np.uint32)
        htl, nb = 4, 1 << bb
    hashable = n - htl + 1
    key[hashable:] = nb
    want = np.argsort(key, kind="stable").astype(np.uint32)
    good = np.array_equal(want, S)
    cnt = np.bincount(key, minlength=nb + 1)
    wseg = np.concatenate([[0], np.cumsum(cnt)])
    good2 = np.array_equal(wseg[:nb + 2], seg[:nb + 2])
    print("sort n=%d q=%d: S %s seg %s" % (n, q, good, good2), flush=True)
    return good and good2


def check(name, d, q, w, use_ref=True):
    t = time.time()
    got = brotli_b200.compress_oneshot(d, q, w)
    dt = time.time() - t
    st = brotli_b200.last_stats()
    want = ref.compress(d, q, w) if use_ref else ora.compress(d, q, w)
    ok = got == want
    k = -1
    if not ok:
        k = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), min(len(got), len(want)))
    print("%s n=%d q=%d w=%d -> %d (want %d) %s  wall %.1f ms  gpu %.1f ms [index %.1f lz77 %.1f entropy %.1f asm %.1f] iters %d runs %d/%d mbs %d walk %.1f ms in %d launches" % (
        name, len(d), q, w, len(got), len(want), "OK" if ok else "DIFF at %d" % k, dt * 1e3, st["ms_total"],
        st["ms_index"], st["ms_lz77"], st["ms_entropy"], st["ms_assemble"], st["lz77_iterations"],
        st["block_runs"], st["blocks"], st["metablocks"], st["ms_walk"], st["walk_launches"]), flush=True)
    return ok


bad = 0
t1 = synth_text(3_000_000, seed=1)
bad += not check_sort(t1[:200000], 5, 22)
bad += not check_sort(t1, 5, 22)
for name, d in [("tiny", b"x"), ("small", t1[:1000]), ("t64k", t1[:65536]), ("t300k", t1[:300000]), ("t3M", t1),
                ("zeros", bytes(400000)), ("web2M", synth_web(2_000_000)), ("bin2M", synth_binary(2_000_000)),
                ("rand", np.random.RandomState(1).randint(0, 256, 300000, dtype=np.uint8).tobytes())]:
    for q, w in [(5, 22), (9, 24), (7, 18)]:
        bad += not check(name, d, q, w)
print("bad", bad, flush=True)
if bad == 0 or os.environ.get("BIG"):
    t = synth_text(20_000_000, seed=2)
    check("t20M", t, 5, 22)
    check("t20M", t, 5, 22)
    t = synth_text(100_000_000)
    check("t100M", t, 5, 22)
    check("t100M", t, 5, 22)
