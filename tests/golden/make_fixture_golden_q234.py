"""Generates tests/golden/fixtures_q234.json: digests of the REFERENCE encoder's output (oracle/_ref, built from
/root/reference by oracle/Makefile) at qualities 2..4 for the reference's own test files under tests/golden/fixtures/
(see make_fixture_golden.py).  Kept in a file of its own: the quality 2..4 GPU tests run last (tests/test_zz_gpu_q234.py)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from brotli_libs import Ref

FIX = os.path.join(HERE, "fixtures")
ref = Ref()
out = []
for name in sorted(os.listdir(FIX)):
    d = open(os.path.join(FIX, name), "rb").read()
    variants = [(name, d)]
    if name == "alice29.txt":
        variants.append(("alice29.txt[:65536]", d[:65536]))
    for label, data in variants:
        for q, w in ((2, 22), (3, 22), (4, 22), (4, 16), (3, 12), (2, 18)):
            comp = ref.compress(data, q, w)
            out.append(dict(file=name, label=label, n=len(data), q=q, lgwin=w, in_sha256=hashlib.sha256(data).hexdigest(),
                            out_len=len(comp), out_sha256=hashlib.sha256(comp).hexdigest()))
            print(label, len(data), q, w, len(comp))
json.dump(out, open(os.path.join(HERE, "fixtures_q234.json"), "w"), indent=1)
