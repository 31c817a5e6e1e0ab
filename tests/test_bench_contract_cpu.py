"""bench.py's JSON contract, exercised on the CPU: the whole bench flow (resident leg, end-to-end leg through
handle_one_file, CPU baseline, parity sample, --quick, --impl reference) runs against the host pipeline and the engine's
CPU twin (tests/hostsim) with a handful of reads; only the numbers are meaningless there."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMDIR = os.path.join(ROOT, "tests", "hostsim")

DRIVER = r'''
import runpy, sys
sys.path.insert(0, %r)
from mtr_b200 import capi
capi.LIB_PATH = %r
import torch
torch.cuda.synchronize = lambda *a, **k: None
sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path(%r, run_name="__main__")
''' % (ROOT, os.path.join(SIMDIR, "_build", "libmtr_hostsim.so"), os.path.join(ROOT, "bench.py"))


def run_bench(*args):
    subprocess.check_call(["make", "-s", "-C", SIMDIR])
    p = subprocess.run([sys.executable, "-c", DRIVER, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=ROOT)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return json.loads(p.stdout.decode().strip().splitlines()[-1])


def test_bench_line_has_every_contract_key():
    d = run_bench("--reads", "5", "--steps", "2", "--warmup", "1", "--cpu-sample", "4")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "reads_per_s" and d["unit"] == "reads/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
    # both legs processed the same reads: the resident loop's digest and handle_one_file's are digests of the same text
    assert d["output_md5"] == d["e2e"]["output_md5"]
    # ... and the bench-time parity sample against the reference sources is byte-identical
    assert d["parity"].get("identical") is True and d["parity"]["reads"] == 5 and d["parity"]["records"] >= 1, d["parity"]


def test_quick_mode():
    d = run_bench("--reads", "5", "--steps", "3", "--warmup", "1", "--quick")
    assert d["quick"] is True and d["groups"] >= 1 and d["reads_per_s"] > 0 and len(d["md5"]) == 32


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--reads", "8"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=ROOT)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    d = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "reads_per_s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
