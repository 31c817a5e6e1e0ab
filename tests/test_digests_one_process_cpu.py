"""The digest tests of tests/test_pipeline_gpu.py (all shipped and synthetic files x three modes through handle_one_file() in
ONE process, look-ahead depths) with the library and the command line of the simulated device in place of the CUDA ones:
the test bodies themselves are checked here without a GPU, and so is the claim that consecutive files in one process do
not influence each other."""
import os
import subprocess

import pytest

import golden_cases
import test_pipeline_gpu as gpu
import test_zz_one_process_gpu as last

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMDIR = os.path.join(ROOT, "tests", "hostsim")


@pytest.fixture(autouse=True)
def _simulated_device(monkeypatch):
    subprocess.check_call(["make", "-s", "-C", SIMDIR])
    monkeypatch.setattr(gpu, "LIB", os.path.join(SIMDIR, "_build", "libmtr_hostsim.so"))
    monkeypatch.setattr(gpu, "MTR", os.path.join(SIMDIR, "_build", "mTR_hostsim"))


@pytest.fixture(scope="module")
def shipped_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("shipped")
    golden_cases.extract_shipped(str(d))
    return str(d)


@pytest.fixture(scope="module")
def synthetic_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("synthetic")
    for name, (reads, lw) in golden_cases.synthetic_cases().items():
        golden_cases.write_case(os.path.join(str(d), name + ".fa"), reads, lw)
    return str(d)


def test_every_digest_in_one_process(shipped_dir, synthetic_dir):
    last.test_consecutive_files_in_one_process(shipped_dir, synthetic_dir)
    for name in ("10_20.fasta", "5_10.fasta"):              # (one with, one without the extra command-line run)
        gpu.test_shipped_multiple_TRs(shipped_dir, synthetic_dir, name)
    for name in ("mixed", "single_TR_5"):
        gpu.test_synthetic_cases(shipped_dir, synthetic_dir, name)


def test_digest_tests_fall_back_to_the_command_line(shipped_dir, synthetic_dir, monkeypatch):
    """A one-process run that fails (here: a library that does not exist) must not take the digest tests with it."""
    monkeypatch.setattr(gpu, "LIB", "/nonexistent/libmtr.so")
    monkeypatch.setattr(gpu, "_MANY_ERRORS", [])
    gpu.test_shipped_multiple_TRs(shipped_dir, synthetic_dir, "5_10.fasta")
    assert len(gpu._MANY_ERRORS) == 1
    gpu._MANY.pop(("/nonexistent/libmtr.so", (), None), None)


@pytest.mark.parametrize("spec", ["0", "24"])
def test_look_ahead_depths_in_one_process(shipped_dir, synthetic_dir, spec):
    gpu.test_speculative_look_ahead_does_not_change_the_output(synthetic_dir, shipped_dir, spec)
