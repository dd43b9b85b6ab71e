"""Generates tests/golden/golden.json: digests of the REFERENCE encoder's output
(oracle/_ref/libbrotli_ref.so, built from /root/reference by oracle/Makefile) for seeded inputs
that regenerate identically anywhere (tests/corpus.py).  The reference ships no encoder golden
vectors of its own, so these pin the oracle and the GPU path."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from brotli_libs import Ref
from golden_cases import CASES, CASES_ORACLE_ONLY, make_case

ref = Ref()
for cases, name in ((CASES, "golden.json"), (CASES_ORACLE_ONLY, "golden_oracle_only.json")):
    out = []
    for c in cases:
        d = make_case(c)
        comp = ref.compress(d, c["q"], c["lgwin"])
        out.append(dict(c, in_sha256=hashlib.sha256(d).hexdigest(), out_len=len(comp),
                        out_sha256=hashlib.sha256(comp).hexdigest()))
        print(c, len(d), len(comp))
    json.dump(out, open(os.path.join(HERE, name), "w"), indent=1)
