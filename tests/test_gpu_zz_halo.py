"""Exact tiling under camera motion (hk_context_set_motion_margin + hk_halo_pull): with the ghost reservoirs refreshed from
their owners after every frame, 2 x 2 tiles reproduce the unsharded frame bit for bit while the camera moves; without the
pulls they drift apart (the limit DESIGN.md 5 documents).  Validated on the emulated kernels when written; (zz: runs last)."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from tests.conftest import Bench
from tests.test_gpu_parity import mismatch

pytestmark = pytest.mark.gpu

PLANES = [L.OUT_TONE_MAPPED, L.OUT_RENDER_DIRECT, L.OUT_RENDER_EMISSIVE, L.OUT_RENDER_INDIRECT] + [L.OUT_RESERVOIR_0 + i for i in range(10)]


def run(scene, config, size, rects, margin, pulls, frames=8, step=(0.04, 0.01, -0.02)):
    b = Bench(scene, size[0], size[1], config=config)
    full = b.device()
    tiles = [b.device(r[2], r[3], r[0], r[1]) for r in rects]
    for t in tiles:
        t.set_motion_margin(margin)                    # re-allocates the planes; the uploaded scene stays
    differing = 0
    for f in range(1, frames + 1):
        inp = b.moving_inputs(f, step=step)
        full.render_frame(inp)
        for t in tiles:
            t.render_frame(inp)
        for t in tiles:
            t.sync()
        for k in PLANES:
            whole = full.readback(k)
            for r, t in zip(rects, tiles):
                differing += mismatch(t.readback(k), whole[r[2]:r[3], r[0]:r[1]])
        if pulls:
            for t in tiles:
                for other in tiles:
                    if other is not t:
                        t.halo_pull(other)
            for t in tiles:
                t.sync()
    return differing


@pytest.mark.parametrize("scene,config", [("cornell", "cornell_1080p"), ("city", "city_4k")])
def test_tiles_with_halo_pull_equal_unsharded_under_camera_motion(scene, config):
    rects = [(0, 72, 0, 40), (72, 144, 0, 40), (0, 72, 40, 96), (72, 144, 40, 96)]       # (col_begin, col_end, row_begin, row_end)
    assert run(scene, config, (144, 96), rects, margin=12, pulls=True) == 0


def test_without_pulls_the_same_tiles_drift():
    rects = [(0, 72, 0, 96), (72, 144, 0, 96)]
    assert run("cornell", "cornell_1080p", (144, 96), rects, margin=12, pulls=False) > 0
    # static camera: exact without any exchange, with or without a margin
    assert run("cornell", "cornell_1080p", (144, 96), rects, margin=0, pulls=False, frames=5, step=(0.0, 0.0, 0.0)) == 0


def test_halo_pull_refusals():
    from bevy_hikari_b200 import _ffi
    b = Bench("cornell", 64, 48, config="cornell_256")
    a, c = b.device(0, 48, 0, 32), b.device(0, 48, 16, 64)           # overlapping columns
    with pytest.raises(_ffi.HikariError, match="overlap"):
        a.halo_pull(c)
    other = Bench("cornell", 32, 48, config="cornell_256").device()
    with pytest.raises(_ffi.HikariError, match="different frames"):
        a.halo_pull(other)
    with pytest.raises(_ffi.HikariError, match="256"):
        a.set_motion_margin(1000)
