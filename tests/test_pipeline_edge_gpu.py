"""Edge cases of the drop-in (bin/mTR) against the pinned oracle: ragged and tiny inputs, line-ending and case
variants, the reference's abort behaviours (handle_one_file.c:169-293)."""
import os
import subprocess

import numpy as np
import pytest

from mtr_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MTR = os.path.join(ROOT, "bin", "mTR")
ORACLE_BIN = os.path.join(ROOT, "oracle", "mtr_oracle")


def both(path, flags=()):
    a = subprocess.run([MTR, *flags, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    b = subprocess.run([ORACLE_BIN, *flags, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return a, b


def write(tmp_path, name, text):
    p = os.path.join(str(tmp_path), name)
    with open(p, "wb") as f:
        f.write(text)
    return p


def test_tiny_and_ragged_reads(tmp_path, oracle_so):
    rng = np.random.default_rng(3)
    reads = [rng.integers(0, 4, n).astype(np.int8) for n in (1, 2, 5, 9, 10, 11, 19, 20, 21, 40, 99, 100, 101, 250, 999, 1000, 1001)]
    reads += [np.tile(np.array([0, 1, 2], np.int8), 30), np.zeros(64, np.int8), np.tile(np.array([3, 2], np.int8), 12)]
    reads += synth.rand_seq_reads(7, 25, 0.02, 0.03, 0.03, 30, 0, 2, seed=4)[0]      # repeat runs into the read end
    p = os.path.join(str(tmp_path), "ragged.fa")
    synth.write_fasta(p, reads)
    lines = {}
    for flags in ((), ("-a",), ("-p", "-m", "0.7"), ("-m", "0.0"), ("-m", "1")):
        a, b = both(p, flags)
        assert a.returncode == 0 and a.stdout == b.stdout, flags
        lines[flags] = b.stdout.count(b"\n")
    assert lines[()] >= 3 and lines[("-m", "0.0")] >= lines[()] >= lines[("-m", "1")]


def test_line_endings_case_and_wrapping(tmp_path, oracle_so):
    rd = synth.rand_seq_reads(11, 20, 0.02, 0.04, 0.04, 60, 60, 3, seed=8)[0]
    plain = os.path.join(str(tmp_path), "plain.fa")
    synth.write_fasta(plain, rd, ids=["r one", "r2", "r3 extra words"])
    want = subprocess.run([ORACLE_BIN, plain], stdout=subprocess.PIPE, check=True).stdout
    text = open(plain, "rb").read()
    variants = {
        "crlf.fa": text.replace(b"\n", b"\r\n"),
        "lower.fa": b"\n".join(l if l.startswith(b">") else l.lower() for l in text.split(b"\n")),
        "nofinalnl.fa": text.rstrip(b"\n"),
    }
    wrapped = os.path.join(str(tmp_path), "wrapped.fa")
    synth.write_fasta(wrapped, rd, ids=["r one", "r2", "r3 extra words"], line_width=60)
    for name, t in variants.items():
        a, b = both(write(tmp_path, name, t))
        assert a.returncode == 0 and a.stdout == b.stdout == want, name
    a, b = both(wrapped)
    assert a.stdout == b.stdout == want


def test_empty_inputs_and_zero_length_read(tmp_path, oracle_so):
    rd = synth.rand_seq_reads(5, 30, 0.0, 0.03, 0.03, 40, 40, 3, seed=9)[0]
    t0, t1, t2 = (">%d\n%s\n" % (i, synth.to_text(r)) for i, r in enumerate(rd))
    for name, text in (("empty.fa", ""), ("header_only.fa", ">x\n"), ("zero_mid.fa", t0 + ">empty\n" + t1 + t2),
                       ("zero_first.fa", ">e\n" + t0), ("blank_lines.fa", t0 + "\n\n" + t1)):
        a, b = both(write(tmp_path, name, text.encode()))
        assert a.returncode == 0 and a.stdout == b.stdout, name
    # a zero-length read ends the run (handle_one_file.c:283): only the first read is reported
    a, _ = both(os.path.join(str(tmp_path), "zero_mid.fa"))
    assert set(l.split(b"\t")[0] for l in a.stdout.splitlines()) <= {b"0"}


def test_abort_behaviour(tmp_path):
    p = write(tmp_path, "bad.fa", b">x\nACGTNACGT\n")
    a = subprocess.run([MTR, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.returncode == 1 and a.stderr.startswith(b"Invalid character: N") and a.stdout == b""
    # a bad read in the middle of the file: the reference has printed the reads in front of it when it aborts
    # (handle_one_file.c:277-289 parses a read only after the one before it is done); also across group boundaries
    rd = synth.rand_seq_reads(5, 30, 0.0, 0.03, 0.03, 40, 40, 4, seed=9)[0]
    t = [">%d\n%s\n" % (i, synth.to_text(r)) for i, r in enumerate(rd)]
    p = write(tmp_path, "bad_mid.fa", (t[0] + t[1] + t[2] + ">bad\nACGTNACGT\n" + t[3]).encode())
    b = subprocess.run([ORACLE_BIN, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert b.returncode == 1 and b.stdout.count(b"\n") >= 3
    for env in ({}, {"MTR_GROUP_READS": "2"}, {"MTR_GROUP_READS": "1", "MTR_GROUPS_PER_GPU": "1"}):
        a = subprocess.run([MTR, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env))
        assert a.returncode == 1 and a.stderr.startswith(b"Invalid character: N") and a.stdout == b.stdout, env
    # a header line of more than BLK - 1 = 4095 characters: the reference's fgets(s, BLK, fp) reads it in pieces and takes the
    # rest for bases (handle_one_file.c:208)
    p = write(tmp_path, "long_header.fa", (">" + "x" * 5000 + "\n" + synth.to_text(rd[0]) + "\n").encode())
    a = subprocess.run([MTR, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    b = subprocess.run([ORACLE_BIN, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.returncode == 1 and b.returncode == 1 and a.stderr.startswith(b"Invalid character: x") and b.stderr.startswith(b"Invalid character: x") and a.stdout == b""
    a = subprocess.run([MTR, os.path.join(str(tmp_path), "missing.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.returncode == 1 and b"fatal error: cannot open" in a.stderr


def test_long_single_read(tmp_path, oracle_so):
    """One 120 kb read with three repeats of very different unit lengths (the C4 regime: a single long read)."""
    rng = np.random.default_rng(12)
    parts = [rng.integers(0, 4, 20000).astype(np.int8)]
    for ul, cp, seed in ((3, 900, 1), (57, 400, 2), (410, 40, 3)):
        parts.append(synth.rand_seq_reads(ul, cp, 0.02, 0.03, 0.03, 0, 0, 1, seed=seed)[0][0])
        parts.append(rng.integers(0, 4, 15000).astype(np.int8))
    p = os.path.join(str(tmp_path), "long.fa")
    synth.write_fasta(p, [np.concatenate(parts)])
    a, b = both(p)
    assert a.returncode == 0 and a.stdout == b.stdout and b.stdout.count(b"\n") >= 3
