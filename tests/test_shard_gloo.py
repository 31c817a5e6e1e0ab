"""N > 1 host logic on CPU: two gloo ranks shard one FASTA, each processes its contiguous shard (the oracle stands in
for the GPU worker here), rank 0 merges in rank order and must reproduce the single-process output."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtr_b200 import shard, synth  # noqa: E402

WORKER = r'''
import os, sys, subprocess, tempfile
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from mtr_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[4], rank=int(sys.argv[2]), world_size=2)
rank = dist.get_rank()
text = open(sys.argv[3], "rb").read()
plan = shard.plan_shards(shard.read_lengths(text), dist.get_world_size())
start, end = plan[rank]
records = text.split(b">")[1:]
with tempfile.NamedTemporaryFile(suffix=".fa", delete=False) as f:
    f.write(b"".join(b">" + r for r in records[start:end]))
out = subprocess.run([os.path.join(sys.argv[1], "oracle", "mtr_oracle"), f.name], stdout=subprocess.PIPE, check=True).stdout
os.unlink(f.name)
parts = [None, None]
dist.all_gather_object(parts, out)
if rank == 0:
    open(sys.argv[3] + ".merged", "wb").write(shard.merge_outputs(parts))
    open(sys.argv[3] + ".plan", "w").write(repr(plan))
dist.barrier()
dist.destroy_process_group()
'''


def test_plan_shards_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        lengths = [int(x) for x in rng.integers(100, 20000, 57)]
        plan = shard.plan_shards(lengths, world)
        assert len(plan) == world and plan[0][0] == 0 and plan[-1][1] == len(lengths)
        assert all(plan[k][1] == plan[k + 1][0] for k in range(world - 1))
        loads = [sum(lengths[a:b]) for a, b in plan]
        assert max(loads) <= sum(lengths) / world + max(lengths)
    small = shard.plan_shards([5, 5], 4)                        # more ranks than reads: some shards are empty
    assert small[0][0] == 0 and small[-1][1] == 2 and all(small[k][1] == small[k + 1][0] for k in range(3))
    assert shard.read_lengths(b">a\nACGT\nAC\r\n>b\n\n>c\nA\n") == [6, 0, 1]


def test_two_gloo_ranks_reproduce_the_single_process_output(oracle_so):
    # equal-length reads: no cross-read stale state, so the oracle can process a shard as a file of its own
    reads, _ = synth.rand_seq_reads(12, 14, 0.02, 0.05, 0.04, 80, 80, 9, seed=77)
    reads = [r[:300] for r in reads]
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "in.fa")
        synth.write_fasta(fa, reads)
        whole = subprocess.run([os.path.join(ROOT, "oracle", "mtr_oracle"), fa], stdout=subprocess.PIPE, check=True).stdout
        assert whole.count(b"\n") >= 5
        script = os.path.join(tmp, "worker.py")
        open(script, "w").write(WORKER)
        port = str(29600 + os.getpid() % 300)
        procs = [subprocess.Popen([sys.executable, script, ROOT, str(r), fa, port]) for r in range(2)]
        assert all(p.wait(timeout=300) == 0 for p in procs)
        assert open(fa + ".merged", "rb").read() == whole
        plan = eval(open(fa + ".plan").read())
        assert plan[0][1] == plan[1][0] and plan[1][1] == len(reads)
