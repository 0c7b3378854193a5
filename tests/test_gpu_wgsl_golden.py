"""The CUDA path held DIRECTLY against what the reference's own WGSL computes (tests/golden/wgsl_*.npz, see tests/test_wgsl_reference.py
and tools/make_wgsl_golden.py): every buffer and texture of every frame of the shared sequences, bit for bit, through the C ABI on the
exact flavour of the library — no oracle in between."""
import pytest

from tests import wgsl_cases as WC
from tests.test_wgsl_reference import run_and_compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(set(WC.CASES) - WC.LATE_CASES))
def test_cuda_path_reproduces_what_the_reference_shaders_compute(case):
    def make(bench):
        dev = bench.device()
        dev.set_keep_intermediates(True)
        return dev
    run_and_compare(case, make, lambda dev, b: dev.update_instances(b.world))


def _device(bench):
    dev = bench.device()
    dev.set_keep_intermediates(True)
    return dev


def test_cuda_gbuffer_agrees_with_the_rasterised_prepass_cases():
    """the ray-cast G-buffer of the CUDA path against what the reference's prepass.wgsl rasterises
    (tests/golden/wgsl_prepass_*.npz; bounds and method in tests/test_wgsl_prepass.py)"""
    from tests import test_wgsl_prepass as TP
    for case in TP.FIXTURE_CASES:
        bench, g = TP.render_gbuffer(case, _device, lambda dev, b: dev.update_instances(b.world))
        rep = TP.compare(TP.fixture(case), g, bench.width, bench.height, case)
        assert rep["coverage_mismatch"] == 0, (case, rep)
