"""The exported C <-> C++ bridge of libmtr_b200.so on the GPU: chaining(1) and pretty_print_alignment run their DP as a K3
PATH job; the bytes must be the reference's (tests/golden/bridge_pa1.txt, see tests/test_bridge_cpu.py)."""
import pytest

from test_bridge_cpu import LIB, drive, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pa", [0, 1])
def test_bridge_of_the_product_library(pa):
    assert drive(LIB, "ours", pa) == golden(pa)
