"""Generates tests/golden/fixtures.json: digests of the REFERENCE encoder's output (oracle/_ref, built from
/root/reference by oracle/Makefile) for the reference's own test files.  The files under tests/golden/fixtures/
are byte copies of /root/reference/tests/testdata/* (test DATA, not source; bb.binast is cut to its first 2 MiB),
so the GPU tests can read them on the box, where /root/reference does not exist.
BASELINE.json config C1 is alice29.txt[:65536] at quality 5, lgwin 22."""
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
        for q, w in ((1, 22), (5, 22), (9, 24), (6, 18), (9, 17)):
            comp = ref.compress(data, q, w)
            out.append(dict(file=name, label=label, n=len(data), q=q, lgwin=w, in_sha256=hashlib.sha256(data).hexdigest(),
                            out_len=len(comp), out_sha256=hashlib.sha256(comp).hexdigest()))
            print(label, len(data), q, w, len(comp))
json.dump(out, open(os.path.join(HERE, "fixtures.json"), "w"), indent=1)
