"""numpy restatement of the slot-sorted position index of qualities 2..4 (hash_longest_match_quickly_inc.h:27 HashBytes,
:96 Store): positions ordered by (slot, position), the positions without a full 8-byte load last; seg = segment starts.
Test infrastructure: tests/test_sim.py checks it against the sim's index, tests/test_zz_gpu_q234.py the GPU's against it."""
import numpy as np

# (quality, size) -> BUCKET_BITS, BUCKET_SWEEP_BITS, HASH_LEN (hash.h:251-338, quality.h:172)
def hasher_of(q, n):
    if q == 2:
        return 16, 0, 5
    if q == 3:
        return 16, 1, 5
    return (20, 2, 7) if n >= (1 << 20) else (17, 2, 5)


def slot_index(d, q):
    n = len(d)
    bits, sweep_bits, hash_len = hasher_of(q, n)
    a = np.frombuffer(bytes(d) + bytes(8), np.uint8)
    v = np.zeros(n, np.uint64)
    for k in range(8):
        v |= a[k:k + n].astype(np.uint64) << np.uint64(8 * k)
    with np.errstate(over="ignore"):
        key = ((v << np.uint64(64 - 8 * hash_len)) * np.uint64(0x1FE35A7BD3579BD3)) >> np.uint64(64 - bits)
    pos = np.arange(n, dtype=np.uint64)
    slot = (key + (pos & np.uint64(((1 << sweep_bits) - 1) << 3))) & np.uint64((1 << bits) - 1)
    if n >= 8:
        slot[n - 7:] = 1 << bits                               # no full 8-byte load: overflow key
    else:
        slot[:] = 1 << bits
    S = np.argsort(slot, kind="stable").astype(np.uint32)
    seg = np.searchsorted(slot[S], np.arange((1 << bits) + 2), side="left").astype(np.uint32)
    return S, seg, bits
