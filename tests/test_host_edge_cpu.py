"""The edge cases of tests/test_pipeline_edge_gpu.py (ragged and tiny reads, CRLF / lower case / wrapped lines, empty and
zero-length inputs, the reference's abort messages, one 120 kb read) run against the host pipeline linked to the
test-only simulated device (tests/hostsim): FASTA reader, batching and abort paths are host code and are checked here
without a GPU.  Same test bodies, other binary."""
import os
import subprocess

import pytest

import test_pipeline_edge_gpu as edge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMDIR = os.path.join(ROOT, "tests", "hostsim")
SIM = os.path.join(SIMDIR, "_build", "mTR_hostsim")


@pytest.fixture(autouse=True)
def _host_pipeline_on_the_simulated_device(monkeypatch):
    subprocess.check_call(["make", "-s", "-C", SIMDIR])
    monkeypatch.setattr(edge, "MTR", SIM)


test_tiny_and_ragged_reads = edge.test_tiny_and_ragged_reads
test_line_endings_case_and_wrapping = edge.test_line_endings_case_and_wrapping
test_empty_inputs_and_zero_length_read = edge.test_empty_inputs_and_zero_length_read
test_abort_behaviour = edge.test_abort_behaviour
test_long_single_read = edge.test_long_single_read


def test_overlong_read_aborts_like_the_reference(tmp_path):
    """MAX_INPUT_LENGTH (mTR.h:31): a read of 10^6 bases or more is fatal, with the reference's message."""
    p = os.path.join(str(tmp_path), "huge.fa")
    with open(p, "wb") as f:
        f.write(b">big\n" + b"ACGT" * 250000 + b"\n")
    a = subprocess.run([SIM, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.returncode == 1 and a.stderr.startswith(b"fatal error: The length 1000000 is tentatively at most 1000000.")
    with open(p, "wb") as f:                      # ... also when the length is reached in the middle of a line
        f.write(b">big\n" + b"ACGTA" * 150000 + b"\n" + b"C" * 300000 + b"\n")
    a = subprocess.run([SIM, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.returncode == 1 and a.stderr.startswith(b"fatal error: The length 1000000 is tentatively at most 1000000.")
