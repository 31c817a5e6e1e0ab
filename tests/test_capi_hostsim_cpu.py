"""The ctypes layer (mtr_b200/capi.py: Pipeline, run_file, shard loading, counters) against the host pipeline on the
test-only simulated device.  Each case runs in a child interpreter that points capi at tests/hostsim/_build/
libmtr_hostsim.so, so the real libmtr_b200.so of this process (tests/test_abi_cpu.py) is never mixed with it."""
import json
import os
import subprocess
import sys

import pytest

import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMDIR = os.path.join(ROOT, "tests", "hostsim")
DIGESTS = json.load(open(os.path.join(golden_cases.GOLDEN, "digests.json")))

PRELUDE = r'''
import hashlib, json, os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from mtr_b200 import capi, shard
capi.LIB_PATH = %r
''' % (ROOT, os.path.join(ROOT, "tests"), os.path.join(SIMDIR, "_build", "libmtr_hostsim.so"))


@pytest.fixture(scope="module")
def mixed(tmp_path_factory):
    subprocess.check_call(["make", "-s", "-C", SIMDIR])
    d = tmp_path_factory.mktemp("capi")
    reads, lw = golden_cases.synthetic_cases()["mixed"]
    p = os.path.join(str(d), "mixed.fa")
    golden_cases.write_case(p, reads, lw)
    return p


def child(body, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, "-c", PRELUDE + body], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return json.loads(p.stdout.decode().strip().splitlines()[-1])


def test_pipeline_object_modes_and_counters(mixed):
    r = child(r'''
text = open(%r, "rb").read()
out = {}
pipe = capi.Pipeline(0)
n = pipe.load_fasta(text); a = pipe.run(); st = pipe.stats()
out["default"] = hashlib.md5(a).hexdigest(); out["n"] = n; out["reads"] = st["reads"]; out["jobs"] = st["jobs"]; out["cells"] = st["wdp_cells"]
out["again"] = hashlib.md5(pipe.run()).hexdigest()                      # the resident batch can be run repeatedly
out["a"] = hashlib.md5(pipe.run(print_alignment=True)).hexdigest()
pipe.close()
pipe = capi.Pipeline(0, manhattan=False, min_match_ratio=0.7)
pipe.load_fasta(text); out["p"] = hashlib.md5(pipe.run()).hexdigest(); pipe.close()
print(json.dumps(out))
''' % mixed)
    want = DIGESTS["synthetic"]["mixed"]
    assert r["default"] == r["again"] == want["default"]["md5"] and r["a"] == want["a"]["md5"] and r["p"] == want["p_m07"]["md5"]
    assert r["n"] == r["reads"] > 0 and r["jobs"] > 0 and r["cells"] > 0


def test_shards_concatenate_to_the_whole_file(mixed):
    r = child(r'''
text = open(%r, "rb").read()
parts = []
for start, end in shard.plan_shards(shard.read_lengths(text), 3):
    pipe = capi.Pipeline(0)
    assert pipe.load_fasta(text, first=start, count=end - start) == end - start
    parts.append(pipe.run()); pipe.close()
print(json.dumps({"md5": hashlib.md5(shard.merge_outputs(parts)).hexdigest()}))
''' % mixed)
    assert r["md5"] == DIGESTS["synthetic"]["mixed"]["default"]["md5"]


@pytest.mark.parametrize("env", [{}, {"MTR_GROUP_READS": "4", "MTR_GROUPS_PER_GPU": "1"}, {"MTR_GROUP_READS": "3", "MTR_GROUPS_PER_GPU": "3"},
                                 {"MTR_GROUP_READS": "2", "MTR_GPUS": "2", "MTR_GROUPS_PER_GPU": "2"}])
def test_handle_one_file_through_ctypes(mixed, env):
    """capi.run_file = handle_one_file as main.c calls it (stdout captured) + mtr_file_stats; group size, number of
    engine contexts and the two-device round-robin must not change a byte, and the counters must cover every read."""
    r = child(r'''
n, out, st = capi.run_file(%r)
n2, out2, st2 = capi.run_file(%r)                                        # counters reset between calls
print(json.dumps({"n": n, "md5": hashlib.md5(out).hexdigest(), "reads": st["reads"], "jobs": st["jobs"], "same": out == out2 and st2["reads"] == st["reads"]}))
''' % (mixed, mixed), env)
    assert r["md5"] == DIGESTS["synthetic"]["mixed"]["default"]["md5"]
    assert r["n"] == r["reads"] > 0 and r["jobs"] > 0 and r["same"]


def test_speculative_cells_are_counted_apart(mixed):
    """Cells spent on speculative candidates that were pruned after all are launched but are not algorithmic cells:
    mtr_pipeline_stats.spec_cells reports them, and the text is the same with and without speculation."""
    body = r'''
text = open(%r, "rb").read()
pipe = capi.Pipeline(0); pipe.load_fasta(text); out = pipe.run(); st = pipe.stats(); pipe.close()
print(json.dumps({"md5": hashlib.md5(out).hexdigest(), "cells": st["wdp_cells"], "spec": st["spec_cells"], "jobs": st["jobs"]}))
''' % mixed
    a = child(body, {"MTR_SPECULATE": "0"})
    b = child(body, {"MTR_SPECULATE": "24"})
    assert a["md5"] == b["md5"] == DIGESTS["synthetic"]["mixed"]["default"]["md5"]
    assert b["jobs"] >= a["jobs"] and b["spec"] >= a["spec"] >= 0
    assert b["cells"] - b["spec"] <= a["cells"]          # what is left after removing the waste is at most the reference's work
