"""The CUDA path against the WGSL fixtures that were added after the round's last full device run (tests/wgsl_cases.py LATE_CASES: BASELINE
configs[2] = examples/scene.rs, and five corners of HikariSettings).  Same test as tests/test_gpu_wgsl_golden.py, in a file of its own that
sorts last; green on a B200 in the round's last call (call 18: 6 passed in 4 s)."""
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
