"""The cooperative half of the kernel-logic emulation (tests/emu/cuda_emu.h: threads of a block as fibers, __syncthreads, ballots, shuffles,
shared memory, waits on another thread's write) against closed-form answers, in ascending and descending thread order — the kc_* kernels'
CPU validation rests on it."""
import os
import subprocess

from tests.conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu")


def test_cooperative_emulation_selftest(tmp_path):
    exe = str(tmp_path / "emu_selftest")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-I" + EMU, "-I" + os.path.join(EMU, "include"), os.path.join(EMU, "selftest.cpp"),
                        os.path.join(EMU, "emu_runtime.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for env in ({}, {"HK_EMU_REVERSE": "1"}):
        r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, **env), timeout=120)
        assert r.returncode == 0 and "selftest ok" in r.stdout, (env, r.stdout, r.stderr)
