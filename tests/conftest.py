import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


EMULATED = bool(os.environ.get("HK_EMULATE_KERNELS"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if EMULATED:
        # Development / CI aid (tests/emu/): run the `-m gpu` parity tests against the kernel SOURCES compiled for the host,
        # to check kernel logic without a GPU.  Only ever enabled by this environment variable, only from tests.
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        from bevy_hikari_b200 import _ffi
        _ffi.LIB_PATH = _ffi.HOST_LIB_PATH = build_emu.build()   # one self-contained library: kernels + the whole host mirror


    else:
        # The 0-ulp parity suite runs on the exact flavour of the library (same sources, exact arithmetic in every translation
        # unit); tests of the product's tolerance contract ask for flavor="product" explicitly (tests/test_gpu_tolerance.py).
        from bevy_hikari_b200 import _ffi
        _ffi.DEFAULT_FLAVOR = "exact"


def pytest_collection_modifyitems(config, items):
    """A hung kernel or a dead-locked peer must fail ONE test, not take the whole GPU tier (and the box) with it: every `gpu`
    test gets a wall-clock limit when pytest-timeout is installed (it is in this image)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900))


def needs_real_gpu():
    """for the few tests that exercise the CUDA runtime itself (pinned memory, IPC) rather than kernel logic"""
    if EMULATED:
        pytest.skip("needs a real CUDA device (kernel-logic emulation is active)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Both shared libraries must exist; build them if a fresh checkout has not yet (nvcc cross-compiles on CPU)."""
    from bevy_hikari_b200 import _ffi
    from oracle import oracle
    if not os.path.exists(_ffi.LIB_PATH) or not os.path.exists(_ffi.EXACT_LIB_PATH) or not os.path.exists(oracle.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return True


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class Bench:
    """A scene + camera + settings, ready to feed both the oracle and the CUDA path with identical inputs."""

    def __init__(self, scene_name, width, height, config=None, **settings):
        from bevy_hikari_b200 import plugin, scenes
        self.scene = scenes.SCENE_BUILDERS[scene_name]()
        self.width, self.height = width, height
        self.world = self.scene.populate(plugin.World())
        self.view, self.previous_view, self.lights = self.scene.view_inputs(width, height)
        if config:
            self.settings = scenes.config_settings(config, **settings)
        else:
            kw = dict(taa=plugin.TAA_NONE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=1.0)
            kw.update(settings)
            self.settings = plugin.HikariSettings(**kw)

    def inputs(self, frame):
        from bevy_hikari_b200 import plugin
        return plugin.make_frame_inputs(self.settings, frame, self.view, self.previous_view, self.lights)

    def moving_inputs(self, frame, step=(0.03, 0.01, -0.02)):
        """Camera translating by `step` per frame; previous_view = the view of frame - 1 (view.rs:47-73)."""
        from bevy_hikari_b200 import camera as cam
        from bevy_hikari_b200 import plugin

        def view_at(f):
            eye = tuple(e + s * (f - 1) for e, s in zip(self.scene.eye, step))
            tgt = tuple(t + s * (f - 1) for t, s in zip(self.scene.target, step))
            proj = cam.perspective_infinite_reverse_rh(self.scene.fov, self.width / self.height, self.scene.near)
            return cam.make_view(cam.look_at(eye, tgt), proj, self.width, self.height)
        view = view_at(frame)
        prev = cam.make_previous_view(view_at(max(frame - 1, 1)))
        return plugin.make_frame_inputs(self.settings, frame, view, prev, self.lights)

    def oracle(self, threads=None):
        from bevy_hikari_b200 import plugin
        from oracle import oracle
        o = oracle.Oracle(self.width, self.height, plugin.load_noise(), threads)
        o.upload_scene_desc(self.world.scene_desc())
        return o

    def device(self, row_begin=0, row_end=None, col_begin=0, col_end=None, flavor=None):
        from bevy_hikari_b200 import plugin
        p = plugin.HikariPlugin(self.width, self.height, 0, row_begin, row_end, None, col_begin, col_end, flavor=flavor)
        p.upload_scene(self.world)
        return p


@pytest.fixture(scope="session")
def cornell64():
    return Bench("cornell", 64, 64, config="cornell_256")


def rotation_y_about(angle, center, translate=(0.0, 0.0, 0.0)):
    """Column-major mat4 (16 floats): rotate by `angle` about the vertical axis through `center`, then translate."""
    c, s = np.float32(np.cos(angle)), np.float32(np.sin(angle))
    r = np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], np.float64)   # rows = columns of the matrix
    cx = np.array(center, np.float64)
    t = cx - cx @ r[:3, :3] + np.array(translate, np.float64)
    r[3, :3] = t
    return r.astype(np.float32).reshape(16)


class Animation:
    """A user system moving entities every frame + the per-frame part of MeshMaterialPlugin (transform.rs:31-44,
    instance.rs:352-437) on a Bench's world.  `tracks` maps instance id -> f(frame) -> world-space mat4 applied on top
    of the instance's original transform."""

    def __init__(self, bench, tracks):
        from bevy_hikari_b200 import scenes
        self.bench, self.tracks = bench, tracks
        self.compose = scenes._compose
        self.base = {i: np.array(bench.scene.inst_transform[i], np.float32) for i in tracks}

    def step(self, frame):
        w = self.bench.world
        for i, f in self.tracks.items():
            w.set_instance_transform(i, self.compose(f(frame), self.base[i]))
        w.previous_transform_system()
        w.prepare_instances()
        return w


def cornell_animation(bench):
    """the short box spins and drifts, the ceiling light slides"""
    return Animation(bench, {6: lambda f: rotation_y_about(0.06 * f, (0.33, 0.3, -0.37), (0.01 * f, 0.0, 0.0)),
                             4: lambda f: rotation_y_about(0.0, (0, 0, 0), (0.03 * np.sin(0.7 * f), 0.0, 0.02 * f))})


def city_animation(bench):
    """the emissive earth sphere rotates and bobs (the reference's animated light), one house part slides"""
    return Animation(bench, {1: lambda f: rotation_y_about(0.1 * f, (0.0, 1.0, 0.0), (0.0, 0.05 * np.sin(0.5 * f), 0.0)),
                             2: lambda f: rotation_y_about(0.0, (0, 0, 0), (0.02 * f, 0.0, 0.0))})


def orthographic_inputs(bench, frame, half_height=1.3, shift=(0.0, 0.0, 0.0)):
    """frame inputs for an OrthographicProjection camera at the scene's eye / target (+ `shift` * (frame - 1))"""
    from bevy_hikari_b200 import camera as cam
    from bevy_hikari_b200 import plugin

    def view_at(f):
        eye = tuple(e + s * (f - 1) for e, s in zip(bench.scene.eye, shift))
        tgt = tuple(t + s * (f - 1) for t, s in zip(bench.scene.target, shift))
        proj = cam.orthographic_reverse_rh(half_height, bench.width / bench.height, 0.1, 50.0)
        return cam.make_view(cam.look_at(eye, tgt), proj, bench.width, bench.height)
    view = view_at(frame)
    return plugin.make_frame_inputs(bench.settings, frame, view, cam.make_previous_view(view_at(max(frame - 1, 1))), bench.lights)
