"""Parity at the sizes BASELINE.json / SURVEY.md 8(d) state: C1 1000 reads x 7 unit lengths, C2 1000 nanopore stand-in
reads with -a, C3 500 pacbio stand-in reads with -p -m 0.7, C5 2048 synthetic long reads -- the CUDA pipeline against
digests taken from the reference itself (oracle/_ref/mTR_ref_det, tests/golden/make_golden_full.py; the inputs are
regenerated here from the same seeds, tests/golden_full_cases.py)."""
import hashlib
import json
import os

import pytest

import golden_cases
import golden_full_cases as gfc
from mtr_b200 import capi

pytestmark = pytest.mark.gpu
FULL = json.load(open(os.path.join(golden_cases.GOLDEN, "full_digests.json")))


@pytest.mark.parametrize("name", sorted(gfc.CASES))
def test_full_size_case_matches_the_reference(name):
    flags = gfc.CASES[name]["flags"]
    manhattan = "-p" not in flags
    ratio = float(flags[flags.index("-m") + 1]) if "-m" in flags else 0.6
    out = b""
    for text in gfc.chunks(name):                               # a fresh pipeline per chunk = a fresh process of the reference
        pipe = capi.Pipeline(0, manhattan=manhattan, min_match_ratio=ratio)
        assert pipe.load_fasta(text) == text.count(b">")
        out += pipe.run(print_alignment="-a" in flags)
        pipe.close()
    want = FULL[name]
    assert len(out) == want["bytes"], (name, len(out), want["bytes"])
    assert hashlib.md5(out).hexdigest() == want["md5"], name
