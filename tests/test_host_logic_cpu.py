"""Host logic and engine logic on the CPU: the product's pipeline.cpp + main.c (unmodified) linked against
tests/hostsim/sim_device.cpp, a TEST-ONLY stand-in that answers the kernel-level ABI (mtr_di_run, mtr_wdp_run) with the
oracle, and against tests/hostsim/sim_engine.cpp, which runs the resident engine's own code (mtr_b200/csrc/eng_core.h,
one lane per warp) wave by wave on host memory.  Everything but the DP cells and the directional index -- FASTA reader,
stale-state tracker, group dispatcher, per-read scheduler with candidate look-ahead, k-mer count tables, de Bruijn
walks, polish / vote, DP task emission, chaining, TSV and alignment formatting -- must reproduce the digests taken from
the reference binary (tests/golden/digests.json).  The GPU twins of these tests are in
tests/test_pipeline_gpu.py; the simulator is never part of the product (tests/test_abi_cpu.py)."""
import hashlib
import json
import os
import subprocess

import pytest

import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMDIR = os.path.join(ROOT, "tests", "hostsim")
SIM = os.path.join(SIMDIR, "_build", "mTR_hostsim")
DIGESTS = json.load(open(os.path.join(golden_cases.GOLDEN, "digests.json")))


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-s", "-C", SIMDIR])
    return SIM


def run(binary, flags, path, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([binary] + flags + [path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout


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


CLI_TOO = ("10_20.fasta", "worm_chrII_1.fasta", "mixed", "long4")      # these also go through the command line, one process each


def one_process(shipped_dir, synthetic_dir):
    """All 81 (file, mode) runs through handle_one_file() of the simulated-device library in one process (the same helper
    the GPU suite uses, tests/test_pipeline_gpu.py)."""
    import test_pipeline_gpu as gpu
    old = gpu.LIB
    gpu.LIB = os.path.join(SIMDIR, "_build", "libmtr_hostsim.so")
    try:
        return gpu.outputs_in_one_process(shipped_dir, synthetic_dir)
    finally:
        gpu.LIB = old


@pytest.mark.parametrize("name", sorted(DIGESTS["shipped"]))
def test_host_logic_on_shipped_files(sim, shipped_dir, synthetic_dir, name):
    path = os.path.join(shipped_dir, name)
    outs = one_process(shipped_dir, synthetic_dir)
    for mode, flags in golden_cases.MODES.items():
        assert hashlib.md5(outs[("shipped", name, mode)]).hexdigest() == DIGESTS["shipped"][name][mode]["md5"], (name, mode)
        if name in CLI_TOO:
            assert hashlib.md5(run(sim, flags, path)).hexdigest() == DIGESTS["shipped"][name][mode]["md5"], (name, mode)


@pytest.mark.parametrize("name", sorted(DIGESTS["synthetic"]))
def test_host_logic_on_synthetic_cases(sim, shipped_dir, synthetic_dir, name):
    path = os.path.join(synthetic_dir, name + ".fa")
    outs = one_process(shipped_dir, synthetic_dir)
    for mode, flags in golden_cases.MODES.items():
        assert hashlib.md5(outs[("synthetic", name, mode)]).hexdigest() == DIGESTS["synthetic"][name][mode]["md5"], (name, mode)
        if name in CLI_TOO:
            assert hashlib.md5(run(sim, flags, path)).hexdigest() == DIGESTS["synthetic"][name][mode]["md5"], (name, mode)


def test_grouping_contexts_and_budgets_do_not_change_the_output(sim, synthetic_dir):
    """Group boundaries (the stale state crosses them), the number of engine contexts, the two-device round-robin and
    the per-wave budgets of the engine (direction-matrix bytes, task slots: a chain that does not fit is emitted again
    by the next wave) and the number of read slots (reads at work at the same time) are scheduling only: not a byte may
    change."""
    path = os.path.join(synthetic_dir, "mixed.fa")
    ref = DIGESTS["synthetic"]["mixed"]["default"]["md5"]
    for env in ({"MTR_GROUP_READS": "1"}, {"MTR_GROUP_READS": "5", "MTR_GROUPS_PER_GPU": "1"}, {"MTR_GROUP_READS": "3", "MTR_GPUS": "2"},
                {"MTR_GROUP_READS": "7", "MTR_GROUPS_PER_GPU": "3", "MTR_GPUS": "2"}, {"MTR_ENGINE_DIR_KB": "600"},
                {"MTR_ENGINE_TASK_CAP": "6", "MTR_GROUP_READS": "4"},
                # read slots: a slot takes the next read of the group when its read has finished
                {"MTR_ENGINE_SLOTS": "1"}, {"MTR_ENGINE_SLOTS": "2", "MTR_SIM_READ_ORDER": "1"}, {"MTR_ENGINE_SLOTS": "3", "MTR_GROUPS_PER_GPU": "2", "MTR_GROUP_READS": "11"},
                {"MTR_ENGINE_SLOTS": "2", "MTR_ENGINE_LONG_ROWS": "40", "MTR_SIM_LONG_EVERY": "3", "MTR_SIM_WALK_LAG": "2"},
                # DP queues whose results arrive waves later (and waves that find no free queue at all)
                {"MTR_SIM_SHORT_EVERY": "3", "MTR_SIM_LONG_EVERY": "5", "MTR_ENGINE_LONG_ROWS": "60"},
                {"MTR_SIM_SHORT_EVERY": "2", "MTR_SIM_WALK_LAG": "1", "MTR_ENGINE_SLOTS": "4", "MTR_SPECULATE": "3"},
                # the file reader: blocks of any size cut at record boundaries and parsed in parallel / the reference's own
                # line-by-line loop
                {"MTR_READ_BLOCK_BYTES": "64", "MTR_GROUP_READS": "3"}, {"MTR_READ_BLOCK_BYTES": "1000", "MTR_PARSE_THREADS": "3"},
                {"MTR_READ_BLOCK_BYTES": "70000"}, {"MTR_SERIAL_READER": "1", "MTR_GROUP_READS": "4"}):
        assert hashlib.md5(run(sim, [], path, env)).hexdigest() == ref, env


@pytest.mark.parametrize("spec", ["0", "1", "24"])
def test_speculative_look_ahead_does_not_change_the_output(sim, synthetic_dir, shipped_dir, spec):
    """MTR_SPECULATE (default 8): candidates evaluated ahead of ones that could still prune them are dropped when they
    are pruned after all -- the reference would never have visited them.  Any depth, incl. none, gives the same bytes
    in every mode (the -a path indexes its PATH jobs behind jobs a dropped candidate may have queued)."""
    cases = [(os.path.join(synthetic_dir, n + ".fa"), DIGESTS["synthetic"][n]) for n in ("mixed", "single_TR_10", "pacbio_200_200", "long4")]
    cases += [(os.path.join(shipped_dir, n), DIGESTS["shipped"][n]) for n in ("worm_chrII_1.fasta", "2_5_10_20_50_100_200_set.fasta")]
    for path, want in cases:
        for mode, flags in golden_cases.MODES.items():
            assert hashlib.md5(run(sim, flags, path, {"MTR_SPECULATE": spec})).hexdigest() == want[mode]["md5"], (path, mode, spec)
