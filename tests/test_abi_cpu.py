"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol include/mtr_b200.h declares,
the ctypes mirror agrees with the header, and the product fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from mtr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "mtr_b200.h")).read()


def header_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(mtr_[a-z_0-9]+|handle_one_file|handle_one_read|insert_an_alignment_into_set|chaining|pretty_print_alignment|print_freq)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    declared = header_functions()
    assert set(declared) == set(capi.ABI_FUNCTIONS), (declared, capi.ABI_FUNCTIONS)
    for name in declared:
        assert hasattr(lib, name), name
    for g in capi.ABI_GLOBALS:
        assert g in HEADER
        C.c_int.in_dll(lib, g)


def test_struct_layouts_match_header():
    assert C.sizeof(capi.WdpJob) == 48 and capi.JOB_DTYPE.itemsize == 48
    assert C.sizeof(capi.WdpResult) == 48
    assert capi.JOB_DTYPE.fields["aux_off"][1] == 32 and capi.JOB_DTYPE.fields["n_param"][1] == 26


def test_pack_reads_layout():
    rng = np.random.default_rng(0)
    reads = [rng.integers(0, 4, n).astype(np.int8) for n in (1, 15, 16, 17, 63, 64, 1000)]
    tails = [(int(rng.integers(4)), int(rng.integers(4))) for _ in reads]
    packed, woff, lens = capi.pack_reads(reads, tails)
    assert np.all(woff % 4 == 0)
    for r, rd in enumerate(reads):
        ext = np.concatenate([rd, tails[r]])
        for b in range(len(ext)):
            assert (int(packed[woff[r] + b // 16]) >> (2 * (b % 16))) & 3 == ext[b]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.MtrError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)
    p = subprocess.run([os.path.join(ROOT, "bin", "mTR"), os.path.join(ROOT, "README.md")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"no CPU fallback" in p.stderr and p.stdout == b""


def test_product_does_not_reference_the_oracle():
    """Nothing under mtr_b200/ or include/ may include, link or import oracle/ code."""
    for base in ("mtr_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".cu", ".cpp", ".c", ".h", ".py")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f)).read()
                    assert "mtr_oracle" not in text and "liboracle" not in text and "oracle_lib" not in text, os.path.join(dirpath, f)
    out = subprocess.run(["ldd", capi.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    assert "oracle" not in out


def test_cli_usage_errors_match_reference():
    mtr = os.path.join(ROOT, "bin", "mTR")
    p = subprocess.run([mtr], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 1 and p.stderr == b"The input file name is expected argument after options\n"
    p = subprocess.run([mtr, "-m", "1.5", "x.fa"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 1 and p.stderr == b"The input minimum match ratio must range from 0 to 1.\n"
    p = subprocess.run([mtr, "-z"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 1 and b"mTR [-acp] [-m ratio] <fasta file name>" in p.stderr
