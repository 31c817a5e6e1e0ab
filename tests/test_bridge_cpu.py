"""The reference's C <-> C++ bridge (mTR.h:146-175, chaining.h:30-56: insert_an_alignment_into_set, chaining,
pretty_print_alignment, print_freq) as exported by the product, against the reference's own functions.

tests/golden/bridge_pa{0,1}.txt are what oracle/_ref/libmtr_ref.so (the unmodified reference sources, oracle/Makefile ref)
prints for tests/bridge_driver.py; they are re-derived here whenever that library is present."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "bridge_driver.py")
REF = os.path.join(ROOT, "oracle", "_ref", "libmtr_ref.so")
SIM = os.path.join(ROOT, "tests", "hostsim", "_build", "libmtr_hostsim.so")
LIB = os.path.join(ROOT, "mtr_b200", "lib", "libmtr_b200.so")


def golden(pa):
    return open(os.path.join(ROOT, "tests", "golden", "bridge_pa%d.txt" % pa), "rb").read()


def drive(lib, kind, pa, *more):
    p = subprocess.run([sys.executable, DRIVER, lib, kind, str(pa)] + list(more), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout


def without_dp_section(text):
    a = text.index(b"-- pretty_print_alignment alone")
    b = text.index(b"-- print_freq")
    return text[:a] + text[b:]


@pytest.mark.parametrize("pa", [0, 1])
def test_golden_is_what_the_reference_prints(pa):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libmtr_ref.so not built (no /root/reference here)")
    assert drive(REF, "ref", pa) == golden(pa)


@pytest.mark.parametrize("pa", [0, 1])
def test_bridge_on_the_simulated_device(pa):
    """The product's host sources with the DP answered by the oracle: records, chain choice, alignment text, k-mer digits."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    got = drive(SIM, "ours", pa)
    assert got == golden(pa)
    assert got.count(b"match gain") == (5 if pa else 1)        # three chained repeats + the one-repeat read + the call of its own


def test_bridge_of_the_product_library_without_a_gpu():
    """chaining(0) and print_freq are host code of libmtr_b200.so itself: no device needed, same bytes as the reference."""
    assert drive(LIB, "ours", 0, "nodp") == without_dp_section(golden(0))
