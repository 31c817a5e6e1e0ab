"""A bounded run of tools/fuzz_vs_reference.py: random small FASTA files, flags and scheduling knobs through the product's
host sources + engine device code on the simulated device (tests/hostsim) against the reference binary built from the
reference's own sources (oracle/_ref/mTR_ref_det).  Skipped where that binary does not exist."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_files_flags_and_schedules_match_the_reference():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mTR_ref_det")):
        pytest.skip("oracle/_ref/mTR_ref_det not built (no /root/reference here)")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_vs_reference.py"), "4242", "20"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    last = p.stdout.decode().strip().splitlines()[-1].split()
    assert last[0] == "done" and int(last[last.index("cases") + 1]) >= 20 and int(last[last.index("bad") + 1]) == 0, p.stdout.decode()[-3000:]
