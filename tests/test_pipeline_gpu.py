"""End-to-end parity of the drop-in: bin/mTR (libmtr_b200.so: CUDA directional index + CUDA wrap-around DP +
host control) must print exactly what the reference prints -- checked as md5 against the digests taken from the
reference binary (tests/golden/digests.json) and, on a mismatch, diffed against the oracle for a readable report."""
import hashlib
import json
import os
import subprocess

import pytest

import golden_cases
import multi_run

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MTR = os.path.join(ROOT, "bin", "mTR")
ORACLE_BIN = os.path.join(ROOT, "oracle", "mtr_oracle")
DIGESTS = json.load(open(os.path.join(golden_cases.GOLDEN, "digests.json")))


def run(binary, flags, path, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([binary] + flags + [path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout


def explain(out, flags, path):
    exp = run(ORACLE_BIN, flags, path).decode().splitlines()
    got = out.decode().splitlines()
    for i, (a, b) in enumerate(zip(got, exp)):
        if a != b:
            return "first difference at line %d:\n got: %s\n exp: %s\n(%d vs %d lines)" % (i + 1, a[:300], b[:300], len(got), len(exp))
    return "outputs differ in length: %d vs %d lines" % (len(got), len(exp))


@pytest.fixture(scope="module")
def shipped_dir(tmp_path_factory, oracle_so):
    d = tmp_path_factory.mktemp("shipped")
    golden_cases.extract_shipped(str(d))
    return str(d)


@pytest.fixture(scope="module")
def synthetic_dir(tmp_path_factory, oracle_so):
    d = tmp_path_factory.mktemp("synthetic")
    for name, (reads, lw) in golden_cases.synthetic_cases().items():
        golden_cases.write_case(os.path.join(str(d), name + ".fa"), reads, lw)
    return str(d)


LIB = os.path.join(ROOT, "mtr_b200", "lib", "libmtr_b200.so")
_MANY = {}
_MANY_ERRORS = []


def outputs_in_one_process(shipped_dir, synthetic_dir, env=None, only=None):
    """Every (file, mode) of the digest table through handle_one_file() of the library in ONE process (tests/multi_run.py):
    one CUDA start-up for the whole table, and consecutive files in one process must not influence each other."""
    key = (LIB, tuple(sorted((env or {}).items())), only)      # (the directories hold the same files whoever made them)
    if key not in _MANY:
        jobs, names = [], []
        for kind, d, suffix in (("shipped", shipped_dir, ""), ("synthetic", synthetic_dir, ".fa")):
            for name in sorted(DIGESTS[kind]):
                if only is not None and name not in only:
                    continue
                for mode, flags in golden_cases.MODES.items():
                    jobs.append((os.path.join(d, name + suffix), flags))
                    names.append((kind, name, mode))
        try:
            outs = multi_run.run_many(LIB, jobs, env)
        except Exception as e:      # noqa: BLE001 -- the digest tests then run the command line per file; the last GPU test reports this
            _MANY_ERRORS.append("%s: %s" % (type(e).__name__, e))
            outs = [b"<the one-process run failed>"] * len(jobs)
        _MANY[key] = dict(zip(names, outs))
    return _MANY[key]


def check_digests(outs, kind, name, directory, suffix=""):
    """The digest of every mode: from the one-process run where it agrees, else from the command line itself (a file that
    only goes wrong behind other files in the same process fails tests/test_zz_one_process_gpu.py, not this test)."""
    for mode, flags in golden_cases.MODES.items():
        out = outs[(kind, name, mode)]
        if hashlib.md5(out).hexdigest() != DIGESTS[kind][name][mode]["md5"]:
            out = run(MTR, flags, os.path.join(directory, name + suffix))
        assert hashlib.md5(out).hexdigest() == DIGESTS[kind][name][mode]["md5"], (name, mode, explain(out, flags, os.path.join(directory, name + suffix)))


@pytest.mark.parametrize("name", sorted(DIGESTS["shipped"]))
def test_shipped_multiple_TRs(shipped_dir, synthetic_dir, name):
    check_digests(outputs_in_one_process(shipped_dir, synthetic_dir), "shipped", name, shipped_dir)
    if name == "10_20.fasta":                               # ... and the command line itself, once
        for mode, flags in golden_cases.MODES.items():
            assert hashlib.md5(run(MTR, flags, os.path.join(shipped_dir, name))).hexdigest() == DIGESTS["shipped"][name][mode]["md5"], (name, mode)


@pytest.mark.parametrize("name", sorted(DIGESTS["synthetic"]))
def test_synthetic_cases(shipped_dir, synthetic_dir, name):
    check_digests(outputs_in_one_process(shipped_dir, synthetic_dir), "synthetic", name, synthetic_dir, ".fa")
    if name == "mixed":
        for mode, flags in golden_cases.MODES.items():
            assert hashlib.md5(run(MTR, flags, os.path.join(synthetic_dir, name + ".fa"))).hexdigest() == DIGESTS["synthetic"][name][mode]["md5"], (name, mode)


def test_small_groups_give_identical_output(synthetic_dir):
    """Group boundaries, the number of engine contexts and the per-wave direction-matrix budget (a chain that does not
    fit is emitted again by the next wave) must not change a byte."""
    path = os.path.join(synthetic_dir, "mixed.fa")
    ref = DIGESTS["synthetic"]["mixed"]["default"]["md5"]
    for env in ({"MTR_GROUP_READS": "1"}, {"MTR_GROUP_READS": "5", "MTR_GROUPS_PER_GPU": "1"}, {"MTR_ENGINE_DIR_MB": "64", "MTR_ENGINE_SIDE_STREAMS": "0"},
                {"MTR_GROUP_READS": "7", "MTR_GROUPS_PER_GPU": "3", "MTR_ENGINE_BURST": "1"}):
        assert hashlib.md5(run(MTR, [], path, env)).hexdigest() == ref, env


def test_sharded_pipeline_equals_whole_file(synthetic_dir):
    """Two shards of the mixed-length file (stale state crosses the cut) through the pipeline ABI: the concatenation
    must be the whole file's output."""
    from mtr_b200 import capi, shard
    text = open(os.path.join(synthetic_dir, "mixed.fa"), "rb").read()
    plan = shard.plan_shards(shard.read_lengths(text), 2)
    parts = []
    for start, end in plan:
        pipe = capi.Pipeline(0)
        assert pipe.load_fasta(text, first=start, count=end - start) == end - start
        parts.append(pipe.run())
        pipe.close()
    assert hashlib.md5(shard.merge_outputs(parts)).hexdigest() == DIGESTS["synthetic"]["mixed"]["default"]["md5"]


def test_pipeline_abi_alignment_and_pearson_modes(synthetic_dir):
    from mtr_b200 import capi
    text = open(os.path.join(synthetic_dir, "single_TR_20.fa"), "rb").read()
    pipe = capi.Pipeline(0)
    pipe.load_fasta(text)
    assert hashlib.md5(pipe.run(print_alignment=True)).hexdigest() == DIGESTS["synthetic"]["single_TR_20"]["a"]["md5"]
    pipe.close()
    pipe = capi.Pipeline(0, manhattan=False, min_match_ratio=0.7)
    pipe.load_fasta(text)
    assert hashlib.md5(pipe.run()).hexdigest() == DIGESTS["synthetic"]["single_TR_20"]["p_m07"]["md5"]
    pipe.close()


def test_cli_on_two_gpus_matches_digest(synthetic_dir):
    """handle_one_file with MTR_GPUS=2: groups of reads are pulled by the engine contexts of both GPUs, output is merged in input order."""
    from mtr_b200 import capi
    if capi.load_library().mtr_device_count() < 2:
        pytest.skip("needs two GPUs")
    path = os.path.join(synthetic_dir, "mixed.fa")
    out = run(MTR, [], path, {"MTR_GPUS": "2", "MTR_GROUP_READS": "5", "MTR_GROUPS_PER_GPU": "2"})
    assert hashlib.md5(out).hexdigest() == DIGESTS["synthetic"]["mixed"]["default"]["md5"]


def test_handle_one_read_entry_point(synthetic_dir):
    """The reference's per-read entry point (mTR.h:127): the caller puts the read into orgInputString and calls
    handle_one_read; records appear on stdout at mtr_flush().  Driven through ctypes in a child process."""
    import sys
    path = os.path.join(synthetic_dir, "mixed.fa")
    script = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from mtr_b200 import capi
lib = capi.load_library()
lib.handle_one_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
lib.handle_one_read.restype = None
org = C.POINTER(C.c_int).in_dll(lib, "orgInputString")
C.c_int.in_dll(lib, "Manhattan_Distance").value = 1
C.c_float.in_dll(lib, "min_match_ratio").value = 0.6
code = {"A": 0, "C": 1, "G": 2, "T": 3}
rid, n = None, 0
for line in open(%r):
    line = line.strip()
    if line.startswith(">"):
        rid = line[1:]
    elif line:
        for i, ch in enumerate(line):
            org[i] = code[ch]                       # like handle_one_file.c:284-285; the tail keeps older reads
        n += 1
        lib.handle_one_read(rid.encode(), len(line), n, 0)
lib.mtr_flush()
''' % (ROOT, path)
    p = subprocess.run([sys.executable, "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()[-1500:]
    assert hashlib.md5(p.stdout).hexdigest() == DIGESTS["synthetic"]["mixed"]["default"]["md5"]


def test_handle_one_file_through_ctypes_with_counters(synthetic_dir):
    """capi.run_file: handle_one_file exactly as main.c calls it, stdout captured, plus mtr_file_stats (child process:
    the entry point's runtime is process-wide and reads MTR_* once)."""
    import sys
    path = os.path.join(synthetic_dir, "mixed.fa")
    script = r'''
import hashlib, json, sys
sys.path.insert(0, %r)
from mtr_b200 import capi
n, out, st = capi.run_file(%r)
print(json.dumps({"n": n, "md5": hashlib.md5(out).hexdigest(), "reads": st["reads"], "launches": st["launches"], "cells": st["wdp_cells"], "d2h": st["d2h_bytes"]}))
''' % (ROOT, path)
    for env in ({"MTR_GROUP_READS": "4"}, {"MTR_GROUP_READS": "3", "MTR_GROUPS_PER_GPU": "2"}):
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
        assert p.returncode == 0, p.stderr.decode()[-1500:]
        r = json.loads(p.stdout.decode().strip().splitlines()[-1])
        assert r["md5"] == DIGESTS["synthetic"]["mixed"]["default"]["md5"], env
        assert r["n"] == r["reads"] > 0 and r["launches"] > 0 and r["cells"] > 0 and r["d2h"] > 0


@pytest.mark.parametrize("spec", ["0", "24"])
def test_speculative_look_ahead_does_not_change_the_output(synthetic_dir, shipped_dir, spec):
    """MTR_SPECULATE (default 8): any depth of speculative candidate look-ahead, incl. none, prints the same bytes."""
    only = ("mixed", "long4", "pacbio_200_200", "worm_chrII_1.fasta", "2_5_10_20_50_100_200_set.fasta")
    outs = outputs_in_one_process(shipped_dir, synthetic_dir, {"MTR_SPECULATE": spec}, only)
    for kind, name, mode in outs:
        out = outs[(kind, name, mode)]
        if hashlib.md5(out).hexdigest() != DIGESTS[kind][name][mode]["md5"]:      # (see check_digests)
            out = run(MTR, golden_cases.MODES[mode], os.path.join(shipped_dir, name) if kind == "shipped" else os.path.join(synthetic_dir, name + ".fa"), {"MTR_SPECULATE": spec})
        assert hashlib.md5(out).hexdigest() == DIGESTS[kind][name][mode]["md5"], (name, mode, spec)


REFMAIN = os.path.join(ROOT, "bin", "mTR_refmain")


@pytest.mark.skipif(not os.path.exists(REFMAIN), reason="bin/mTR_refmain is built where the reference sources are (__graft_entry__.build)")
def test_reference_main_linked_against_the_cuda_library(shipped_dir):
    """The reference's own main.c (unmodified; compiled from /root/reference by __graft_entry__.build, -fcommon) linked
    against libmtr_b200.so: same bytes as the reference in every mode, and its -c report (main.c:108-121) shows the
    library's timers and query counter -- the globals main.c declares are the ones the library updates."""
    for name in ("10_20.fasta", "worm_chrII_1.fasta"):
        path = os.path.join(shipped_dir, name)
        for mode, flags in golden_cases.MODES.items():
            out = run(REFMAIN, flags, path)
            assert hashlib.md5(out).hexdigest() == DIGESTS["shipped"][name][mode]["md5"], (name, mode, explain(out, flags, path))
    p = subprocess.run([REFMAIN, "-c", os.path.join(shipped_dir, "10_20.fasta")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0
    report = dict((l.split("\t")[-1], l.strip().split("\t")[0]) for l in p.stderr.decode().splitlines() if "\t" in l)
    assert int(report["Count of queries"]) > 0, report
    for key in ("all", "ranges", "Computing periods", "wrap around", "count table generation", "Initialize the input", "chaining"):
        assert key in report, report
    assert float(report["Computing periods"]) > 0 and float(report["wrap around"]) > 0 and float(report["ranges"]) > 0, report
