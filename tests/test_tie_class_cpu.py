"""The stock reference binary orders its alignment set by heap addresses, so on inputs with ties it
does not print the same bytes as the insertion-ordered variant this repo is bit-exact with (oracle/_ref/mTR_ref_det).
tests/golden/tie_flips.json lists EVERY difference on the full-size cases (tests/golden/make_golden_full.py: 32 k
records).  Checked here, per record replaced one-for-one:
  rotation   : columns 1-12 equal (read, span, period, counts) and the unit a rotation of the other  -- 93 % of the flips
  same locus : same read and the same rep_end: the other of two overlapping alternatives survived the chaining sweep
               (its multimaps are seeded in the set's order, chaining.cpp:251-259)                    -- the rest
and the differences that are not one-for-one (one binary prints two records where the other prints one) stay a listed,
small set.  Nothing else may appear."""
import json
import os

import pytest

import golden_cases

PATH = os.path.join(golden_cases.GOLDEN, "tie_flips.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/tie_flips.json not generated")


def is_rotation(a, b):
    return len(a) == len(b) and a in b + b


def test_every_flip_is_in_the_tie_class():
    flips = json.load(open(PATH))
    assert flips, "no cases"
    n_rot = n_locus = n_other = n_records = 0
    for name, d in flips.items():
        n_records += d["records"]
        for det, stock in d["flips"]:
            a, b = det.split("\t"), stock.split("\t")
            assert len(a) == len(b) == 13, (name, det, stock)
            assert a[:2] == b[:2], (name, det, stock)                       # same read, same length
            if a[:12] == b[:12]:
                assert is_rotation(a[12], b[12]) and a[12] != b[12], (name, det, stock)
                n_rot += 1
            else:
                assert a[3] == b[3] or a[2] == b[2], (name, det, stock)      # same locus: one end of the repeat shared
                n_locus += 1
        for o in d["other"]:                                               # listed, same read on both sides
            ids = {l.split("\t")[0] for l in o["det"] + o["stock"] if l}
            assert len(ids) == 1, (name, o)
        n_other += len(d["other"])
        assert len(d["other"]) <= max(8, d["records"] // 100), (name, len(d["other"]), d["records"])
    assert n_rot > 0 and n_rot >= 9 * n_locus, (n_rot, n_locus)
    print("tie class: %d rotation flips, %d same-locus flips, %d other differences in %d records" % (n_rot, n_locus, n_other, n_records))


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(golden_cases.HERE), "oracle", "_ref", "mTR_ref_O3")), reason="reference binaries not built")
def test_flips_are_reproducible_on_a_small_case(tmp_path):
    """Live re-check on the 'mixed' case: the stock binary's differences from the canonical variant are rotation flips."""
    import subprocess
    root = os.path.dirname(golden_cases.HERE)
    reads, lw = golden_cases.synthetic_cases()["mixed"]
    p = str(tmp_path / "mixed.fa")
    golden_cases.write_case(p, reads, lw)
    det = subprocess.run([os.path.join(root, "oracle", "_ref", "mTR_ref_det"), p], stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    stock = subprocess.run([os.path.join(root, "oracle", "_ref", "mTR_ref_O3"), p], stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    assert len(det) == len(stock)
    for a, b in zip(det, stock):
        if a != b:
            x, y = a.split("\t"), b.split("\t")
            assert x[:12] == y[:12] and is_rotation(x[12], y[12]), (a, b)
