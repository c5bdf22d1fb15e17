"""GPU parity of the tcgen05 conv / wgrad kernels (through the C ABI) against torch fp32 convolutions."""
import pytest

import conv_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in conv_cases.CASES if not n.startswith("perf")])
def test_conv_case(name):
    r = conv_cases.CASES[name]()
    assert r["ok"], r
