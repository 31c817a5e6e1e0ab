"""Last GPU test of the suite: the 81 handle_one_file() calls that tests/test_pipeline_gpu.py makes in one process (every
shipped and synthetic file in every mode, one after the other on the same engine contexts) each print what they print
alone.  (The digest tests themselves fall back to the command line for a file that differs here.)"""
import hashlib

import pytest

import test_pipeline_gpu as tp
from test_pipeline_gpu import shipped_dir, synthetic_dir      # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def test_consecutive_files_in_one_process(shipped_dir, synthetic_dir):      # noqa: F811
    outs = tp.outputs_in_one_process(shipped_dir, synthetic_dir)
    bad = [k for k, out in outs.items() if hashlib.md5(out).hexdigest() != tp.DIGESTS[k[0]][k[1]][k[2]]["md5"]]
    assert not tp._MANY_ERRORS, tp._MANY_ERRORS
    assert len(outs) == 81 and not bad, bad
