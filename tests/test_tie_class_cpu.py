"""The stock reference binary orders its alignment set by heap addresses, so on inputs with ties it
does not print the same bytes as the insertion-ordered variant this repo is bit-exact with (oracle/_ref/mTR_ref_det).
tests/golden/tie_flips.json lists every difference on the full-size cases (tests/golden/make_golden_full.py).  Checked
here: a record replaced one-for-one differs ONLY in the unit string, and the two units are rotations of each other
(columns 1-12 equal: same read, span, period, counts, k, penalties); whatever is not one-for-one is listed and small."""
import json
import os

import pytest

import golden_cases

PATH = os.path.join(golden_cases.GOLDEN, "tie_flips.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/tie_flips.json not generated")


def is_rotation(a, b):
    return len(a) == len(b) and a in b + b


def test_every_flip_is_a_rotation_of_the_unit():
    flips = json.load(open(PATH))
    assert flips, "no cases"
    n_flips = n_other = n_records = 0
    for name, d in flips.items():
        n_records += d["records"]
        for det, stock in d["flips"]:
            a, b = det.split("\t"), stock.split("\t")
            assert len(a) == len(b) == 13, (name, det, stock)
            assert a[:12] == b[:12], (name, det, stock)
            assert is_rotation(a[12], b[12]) and a[12] != b[12], (name, det, stock)
            n_flips += 1
        n_other += len(d["other"])
        # records only one of the two binaries prints: the same tie decides which of two equal-score chains survives the
        # chaining pass; listed, not explained away -- and rare
        assert len(d["other"]) <= max(3, d["records"] // 50), (name, len(d["other"]), d["records"])
    assert n_flips > 0
    print("tie class: %d rotation flips, %d other differences in %d records" % (n_flips, n_other, n_records))


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
