"""The CUDA path against the WGSL fixtures that were added after the round's last GPU minute (tests/wgsl_cases.py LATE_CASES: BASELINE
configs[2] = examples/scene.rs, and five corners of HikariSettings).  Same test as tests/test_gpu_wgsl_golden.py; in a file of its own that
sorts last, because these sequences have run on the emulated kernels only: every code path they take has been on the device in other
tests (tests/test_gpu_variants.py, test_gpu_zz_examples.py), but not these exact sequences."""
import pytest

from tests import wgsl_cases as WC
from tests.test_wgsl_reference import run_and_compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(WC.LATE_CASES))
def test_cuda_path_reproduces_the_late_fixtures(case):
    def make(bench):
        dev = bench.device()
        dev.set_keep_intermediates(True)
        return dev
    run_and_compare(case, make, lambda dev, b: dev.update_instances(b.world))
