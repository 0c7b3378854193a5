"""bench.py's single-GPU flow (replays, the two timed arms, the one-kernel ring timing, the JSON line) dry-run on the kernel-logic
emulation with a stand-in for the sliver of torch it touches: a Python error in bench.py must not cost a GPU call.  The numbers of
such a run mean nothing and are not looked at; the keys of the contract are."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_bench_line_has_the_contract_keys_on_the_emulated_kernels():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    lib = build_emu.build()
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "emu", "fake_torch"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--lib", lib, "--config", "cornell_256", "--steps", "3", "--warmup", "3",
                        "--no-cpu-baseline", "--print-frame-hash"], capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "frame_ms", "value_vs_e2e", "kernel_ms"):
        assert key in d, key
    assert d["steps"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert d["roofline"]["kernel_ms_source"] == "live, timed region" and d["roofline"]["kernel"] in d["kernel_ms"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["d2h_bytes_per_step"] == 256 * 256 * 8
    assert len(d["frame_check"]["unsharded_sha256"]) == 64
    # animated-scene block: the host path and the device-side rebuild (hk_scene_update_transforms) were both exercised
    assert d["scene_update"]["device_path_taken"] is True and d["scene_update"]["device_rebuild_done_ms"] > 0


def test_reference_arm_never_maps_the_cuda_library():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cornell_256", "--steps", "1",
                        "--warmup", "3"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["mapped_cuda_library"] is False
    assert d["config"]["rendered_width"] == 256 and d["config"]["rendered_pixel_fraction"] == 1.0
