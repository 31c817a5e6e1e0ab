"""Generates the golden fixtures.  Run ONCE in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

  shipped_multiple_TRs.tar.gz : the 15 deterministic inputs of /root/reference/test_multiple_TRs/data
  digests.json                : md5 of the reference's stdout for every (input, mode)
       shipped   : stock -O3 binary AND the insertion-ordered variant (they agree on all 45: hazard-free)
       synthetic : oracle/_ref/mTR_ref_det (canonical tie-break) on the seeded cases of
                   tests/golden_cases.py, plus the stock binary's digest for reference
  min_missing_table.npy       : the 10x10x20 table parsed from consensus.c:714-785
"""
import hashlib
import io
import json
import os
import re
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases  # noqa: E402

REF = "/root/reference"
DET = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_det")
STOCK = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_O3")


def md5_of(cmd):
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    return hashlib.md5(out).hexdigest(), len(out)


def main():
    data = os.path.join(REF, "test_multiple_TRs", "data")
    names = sorted(f for f in os.listdir(data) if f.endswith(".fasta"))
    with tarfile.open(os.path.join(HERE, "shipped_multiple_TRs.tar.gz"), "w:gz") as t:
        for f in names:
            t.add(os.path.join(data, f), arcname=f)
    dig = {"shipped": {}, "synthetic": {}}
    for f in names:
        dig["shipped"][f] = {}
        for mode, flags in golden_cases.MODES.items():
            a, n = md5_of([STOCK] + flags + [os.path.join(data, f)])
            b, _ = md5_of([DET] + flags + [os.path.join(data, f)])
            assert a == b, (f, mode)
            dig["shipped"][f][mode] = {"md5": a, "bytes": n}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (reads, lw) in golden_cases.synthetic_cases().items():
            path = os.path.join(tmp, name + ".fa")
            golden_cases.write_case(path, reads, lw)
            dig["synthetic"][name] = {"n_reads": len(reads), "bases": int(sum(len(r) for r in reads))}
            for mode, flags in golden_cases.MODES.items():
                a, n = md5_of([DET] + flags + [path])
                s, _ = md5_of([STOCK] + flags + [path])
                dig["synthetic"][name][mode] = {"md5": a, "bytes": n, "stock_md5": s}
            print(name, dig["synthetic"][name], flush=True)
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(dig, f, indent=1, sort_keys=True)
    src = open(os.path.join(REF, "consensus.c")).read()
    s = src[src.index("int min_missing_bases[10][10][20] ={"):]
    s = re.sub(r"//.*", "", s[:s.index("};")])
    nums = [int(x) for x in re.findall(r"\d+", s)][3:]
    np.save(os.path.join(HERE, "min_missing_table.npy"), np.array(nums, dtype=np.int8).reshape(10, 10, 20))


if __name__ == "__main__":
    main()
