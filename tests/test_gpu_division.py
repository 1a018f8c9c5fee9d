"""-m gpu: the kernels replace IEEE divisions by x*RN(1/b) plus two exact FMA residual corrections
(div_rcp in adv_kernels.cuh).  That is only legitimate if it is bit-identical to `/`; check it on
2^28 pseudo-random operand pairs for each divisor class the kernels use."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2])   # /6.0 (MUSCL/MFCT), /3.0 (QR4C), / areasvol, hnode_new
def test_reciprocal_division_is_ieee_exact(mode):
    from fesom2_b200.driver import selftest_div
    assert selftest_div(1 << 28, 20261017 + mode, mode) == 0
