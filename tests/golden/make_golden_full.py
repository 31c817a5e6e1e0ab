"""Golden digests at the sizes BASELINE.json / SURVEY.md 8(d) state, taken from the reference itself.  Run ONCE in the
build container (needs oracle/_ref, i.e. /root/reference compiled by `make -C oracle ref`):

    python tests/golden/make_golden_full.py

  full_digests.json : per case, md5 of the concatenated stdout of oracle/_ref/mTR_ref_det (the reference with its
                      alignment set in insertion order) over the case's CHUNKS (every chunk is run as
                      a process of its own, so the cross-read stale state starts fresh per chunk -- the tests load the
                      same chunks one by one), plus the number of records
  tie_flips.json    : where the STOCK binary (oracle/_ref/mTR_ref_O3: std::set ordered by heap addresses) prints
                      something else than mTR_ref_det: per case every one-for-one replaced record ("flips": the two
                      records) and every other difference ("other": records only one of the two prints -- the alignment
                      set's iteration order also decides which of two equal-score chains survives);
                      tests/test_tie_class_cpu.py checks that every flip is the H1 tie class -- columns 1-12 equal, unit
                      a rotation of the other -- and that the rest stays a listed, small set
Cases (tests/golden_full_cases.py): C1 1000 reads x unit lengths {2,5,10,20,50,100,200} (test_single_TR/test.sh shapes),
C2 1000 nanopore stand-in reads with -a, C3 500 pacbio stand-in reads with -p -m 0.7, C5 2048 synthetic long reads.
"""
import difflib
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_full_cases as gfc  # noqa: E402

DET = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_det")
STOCK = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_O3")


def run(binary, flags, path):
    return subprocess.run([binary] + flags + [path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout


def main():
    dig, flips = {}, {}
    only = set(sys.argv[1:])
    for name in gfc.CASES:
        if only and name not in only:
            continue
        flags = gfc.CASES[name]["flags"]
        with tempfile.TemporaryDirectory() as tmp:
            paths = []
            for ci, text in enumerate(gfc.chunks(name)):
                p = os.path.join(tmp, "c%d.fa" % ci)
                open(p, "wb").write(text)
                paths.append(p)
            with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
                det = list(ex.map(lambda p: run(DET, flags, p), paths))
                stock = list(ex.map(lambda p: run(STOCK, flags, p), paths))
        out = b"".join(det)
        dig[name] = {"md5": hashlib.md5(out).hexdigest(), "bytes": len(out), "records": sum(1 for l in out.split(b"\n") if l.count(b"\t") >= 12),
                     "reads": gfc.CASES[name]["reads"], "chunks": len(paths), "flags": flags,
                     "stock_md5": hashlib.md5(b"".join(stock)).hexdigest()}
        fl, other = [], []
        if "-a" not in flags:                                   # (the alignment text follows the unit: records only)
            for a, b in zip(det, stock):
                la, lb = a.decode().split("\n"), b.decode().split("\n")
                for tag, i1, i2, j1, j2 in difflib.SequenceMatcher(None, la, lb, autojunk=False).get_opcodes():
                    if tag == "equal":
                        continue
                    if tag == "replace" and i2 - i1 == j2 - j1:
                        fl += [[x, y] for x, y in zip(la[i1:i2], lb[j1:j2])]
                    else:
                        other.append({"det": la[i1:i2], "stock": lb[j1:j2]})
        flips[name] = {"records": dig[name]["records"], "differing": len(fl), "flips": fl, "other": other}
        print(name, {k: v for k, v in dig[name].items()}, "flips", len(fl), "other", len(other), flush=True)
    for fn, new in (("full_digests.json", dig), ("tie_flips.json", flips)):
        path = os.path.join(HERE, fn)
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.update(new)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
