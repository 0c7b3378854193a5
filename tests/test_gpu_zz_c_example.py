"""examples/c/box.c on the device: a plain-C consumer of the boundary (include/hikari_host.h + include/hikari_b200.h only) builds a scene,
renders 8 frames and reads the image back.  (Named zz: written after the round's GPU budget was spent; host logic validated against the
kernel-logic emulation in tests/test_c_example.py.)"""
import os
import subprocess

import pytest

from tests.conftest import ROOT
from tests.test_c_example import NOISE, build

pytestmark = pytest.mark.gpu


def test_c_example_runs_on_the_device(tmp_path):
    from tests.conftest import needs_real_gpu
    needs_real_gpu()
    exe = build(tmp_path, os.path.join(ROOT, "bevy_hikari_b200"), "hikari_b200")
    r = subprocess.run([exe, NOISE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert int(r.stdout.split("kernel_launches/frame=")[1].split()[0]) >= 10, r.stdout
