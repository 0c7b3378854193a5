"""examples/c/box.c: the boundary consumed from plain C (no Python, no torch in the consumer).  CPU: it compiles and links against
libhikari_b200.so with nothing but the two public headers, fails loudly where there is no CUDA device (no CPU fallback), and its
host logic (scene, camera matrices, frame loop, read-back) runs to completion against the kernel-logic emulation.  On a device: tests/test_gpu_zz_c_example.py."""
import os
import shutil
import subprocess
import sys

import pytest

from tests.conftest import ROOT, has_gpu

SRC = os.path.join(ROOT, "examples", "c", "box.c")
NOISE = os.path.join(ROOT, "data", "noise_rgba8_64x64x16.bin")


def build(tmp_path, libdir, libname):
    """the product is two libraries: libhikari_b200.so (C ABI + HikariPlugin) and its dependency libhikari_host.so (scene
    preparation, settings); the emulated build is one self-contained library"""
    exe = str(tmp_path / ("box_" + libname))
    libs = ["-l" + libname] + (["-lhikari_host"] if libname == "hikari_b200" else [])
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + libdir] + libs + ["-lm",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(not shutil.which("gcc"), reason="no C compiler")
def test_c_example_links_and_fails_loudly_without_a_device(tmp_path):
    from bevy_hikari_b200 import _ffi
    exe = build(tmp_path, os.path.dirname(_ffi.LIB_PATH) if not os.environ.get("HK_EMULATE_KERNELS") else os.path.join(ROOT, "bevy_hikari_b200"),
                "hikari_b200")
    if has_gpu():
        pytest.skip("a device is present: covered by the gpu test")
    r = subprocess.run([exe, NOISE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.skipif(not shutil.which("gcc") or not os.path.exists("/usr/bin/g++"), reason="no host compilers")
def test_c_example_host_logic_on_the_emulated_kernels(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    lib = build_emu.build()
    exe = build(tmp_path, os.path.dirname(lib), "hikari_emu")
    r = subprocess.run([exe, NOISE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    # SmaaTu4x{ratio 1} + Taa::None through HikariPlugin::run_frame: the path + smaa_tu4x + smaa_tu4x_extrapolate, as the reference runs it
    launches = int(r.stdout.split("kernel_launches/frame=")[1].split()[0])
    assert launches >= 10 and "covered=0.5" in r.stdout, r.stdout
