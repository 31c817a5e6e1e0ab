"""Boundary proof: the reference's OWN main.c (/root/reference/main.c, unmodified, compiled with -fcommon as its Makefile
implies) links against this repo's library -- handle_one_file, the option globals and the -c timers are the whole
interface (mTR.h:121-143) -- and prints the reference's bytes.  Here against the host-logic build (tests/hostsim: the
product's pipeline.cpp with the simulated device); __graft_entry__.build() links the same object against
libmtr_b200.so as bin/mTR_refmain, which tests/test_pipeline_gpu.py runs on the GPU."""
import hashlib
import json
import os
import subprocess

import pytest

import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = "/root/reference/main.c"
SIMDIR = os.path.join(ROOT, "tests", "hostsim")
DIGESTS = json.load(open(os.path.join(golden_cases.GOLDEN, "digests.json")))

pytestmark = pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="the reference sources are not on this machine")


@pytest.fixture(scope="module")
def refmain(tmp_path_factory):
    subprocess.check_call(["make", "-s", "-C", SIMDIR])
    d = tmp_path_factory.mktemp("refmain")
    obj, exe = str(d / "main.o"), str(d / "mTR_refmain_sim")
    subprocess.check_call(["gcc", "-O2", "-w", "-fcommon", "-I/root/reference", "-c", REF_MAIN, "-o", obj])
    # linked like the reference links its own objects: the option globals and timers that main.c declares (tentative
    # definitions, hence -fcommon) are the very objects the library defines
    b = os.path.join(SIMDIR, "_build")
    objs = [os.path.join(b, o) for o in ("pipeline.o", "sim_device.o", "sim_engine.o", "mtr_oracle.o", "mtr_oracle_chain.o")]
    subprocess.check_call(["g++", "-o", exe, obj] + objs + ["-pthread", "-lm"])
    return exe


def test_reference_main_links_and_reproduces_the_digests(refmain, tmp_path):
    golden_cases.extract_shipped(str(tmp_path))
    for name in ("10_20.fasta", "2_5_10_20_50_100_200_set.fasta", "worm_chrII_1.fasta"):
        for mode, flags in golden_cases.MODES.items():
            p = subprocess.run([refmain] + flags + [os.path.join(str(tmp_path), name)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert p.returncode == 0, p.stderr.decode()[-500:]
            assert hashlib.md5(p.stdout).hexdigest() == DIGESTS["shipped"][name][mode]["md5"], (name, mode)


def test_reference_main_error_paths(refmain, tmp_path):
    """main.c's own messages and exit codes (main.c:40-100) come out of the reference's code; the library only has to
    keep handle_one_file's contract for a missing file (handle_one_file.c:275-278)."""
    p = subprocess.run([refmain], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 or b"Usage" in p.stderr + p.stdout or b"mTR" in p.stderr + p.stdout
    p = subprocess.run([refmain, str(tmp_path / "does_not_exist.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and p.stdout == b""
