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


def test_halo_descriptor_path_in_one_process_on_the_emulator():
    """hk_halo_export / hk_halo_import / hk_halo_pull_peer with both tiles in one process.  Real CUDA refuses to open an IPC
    handle inside the process that exported it, so this form only runs on the emulated kernels (logic of the descriptor,
    the mapped-plane bookkeeping and the copy rectangle); the two-process form below runs on the device."""
    from tests.conftest import EMULATED
    if not EMULATED:
        pytest.skip("same-process IPC open exists only in the kernel-logic emulation")
    from bevy_hikari_b200 import plugin
    rects = [(0, 72, 0, 96), (72, 144, 0, 96)]
    for upscalers in (False, True):
        b = Bench("cornell", 144, 96, config="cornell_1080p", taa=plugin.TAA_JASMINE if upscalers else plugin.TAA_NONE)
        full = b.device()
        tiles = [b.device(r[2], r[3], r[0], r[1]) for r in rects]
        for t in tiles:
            t.set_motion_margin(12)
            if upscalers:
                t.enable_tile_upscalers()
        peers = {0: tiles[0].halo_import(tiles[1].halo_export()), 1: tiles[1].halo_import(tiles[0].halo_export())}
        scaled = [(L.OUT_UPSCALED, 2), (L.OUT_TAA, 2)] if upscalers else []
        for f in range(1, 8):
            inp = b.moving_inputs(f, step=(0.04, 0.01, -0.02))
            inp.temporal_upscalers = 1 if upscalers else 0
            full.render_frame(inp)
            for t in tiles:
                t.render_frame(inp)
            for k in PLANES:
                whole = full.readback(k)
                for r, t in zip(rects, tiles):
                    assert mismatch(t.readback(k), whole[r[2]:r[3], r[0]:r[1]]) == 0, (f, k)
            for k, s in scaled:
                whole = full.readback(k)
                for r, t in zip(rects, tiles):
                    assert mismatch(t.readback(k), whole[s * r[2]:s * r[3], s * r[0]:s * r[1]]) == 0, (f, k)
            for i, t in enumerate(tiles):
                t.halo_pull_peer(peers[i])


def _halo_worker(rect, frames, conn):
    """second process: renders its tile and the unsharded frame, exchanges halos with the parent after every frame"""
    try:
        from bevy_hikari_b200 import _ffi
        _ffi.DEFAULT_FLAVOR = "exact"     # a spawned process does not run conftest.pytest_configure: same flavour as the parent's contexts
        b = Bench("cornell", 144, 96, config="cornell_1080p")
        full, tile = b.device(), b.device(rect[2], rect[3], rect[0], rect[1])
        tile.set_motion_margin(12)
        def recv():                       # never block for ever: a parent that failed must not leave this process behind
            if not conn.poll(180):
                raise TimeoutError("parent went silent")
            return conn.recv()
        conn.send(tile.halo_export())
        peer = tile.halo_import(recv())
        bad = 0
        for f in range(1, frames + 1):
            inp = b.moving_inputs(f, step=(0.04, 0.01, -0.02))
            full.render_frame(inp)
            tile.render_frame(inp)
            tile.sync()
            for k in PLANES:
                bad += mismatch(tile.readback(k), full.readback(k)[rect[2]:rect[3], rect[0]:rect[1]])
            conn.send("rendered"); assert recv() == "rendered"           # both tiles have finished frame f
            tile.halo_pull_peer(peer)
            tile.sync()
            conn.send("pulled"); assert recv() == "pulled"               # nobody starts frame f + 1 before both pulls are done
        conn.send(("done", bad))
    except Exception as e:   # pragma: no cover
        try:
            conn.send(("error", repr(e)))
        except Exception:
            pass


def test_halo_exchange_between_two_processes_cuda_ipc():
    """One process per tile (the bench's shape): descriptors cross a pipe, every frame ends with pull + barrier on both sides;
    both tiles stay bit-identical to the unsharded render under camera motion.  (Both processes use cuda:0 here; across GPUs
    the pulls go over NVLink.)"""
    import multiprocessing as mp
    from tests.conftest import needs_real_gpu
    needs_real_gpu()
    rects = [(0, 72, 0, 96), (72, 144, 0, 96)]
    frames = 6
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    p = ctx.Process(target=_halo_worker, args=(rects[1], frames, child), daemon=True)   # daemon: never waited for at interpreter exit
    p.start()
    try:
        _halo_parent(parent, p, rects, frames)
    finally:                              # whatever happened above, the worker does not outlive the test
        if p.is_alive():
            p.terminate()
        p.join(10)


def _halo_parent(parent, p, rects, frames):
    b = Bench("cornell", 144, 96, config="cornell_1080p")
    full, tile = b.device(), b.device(rects[0][2], rects[0][3], rects[0][0], rects[0][1])
    tile.set_motion_margin(12)
    assert parent.poll(240), "worker did not start"
    theirs = parent.recv()
    assert not isinstance(theirs, tuple), theirs
    parent.send(tile.halo_export())
    peer = tile.halo_import(theirs)
    bad = 0
    for f in range(1, frames + 1):
        inp = b.moving_inputs(f, step=(0.04, 0.01, -0.02))
        full.render_frame(inp)
        tile.render_frame(inp)
        tile.sync()
        for k in PLANES:
            bad += mismatch(tile.readback(k), full.readback(k)[rects[0][2]:rects[0][3], rects[0][0]:rects[0][1]])
        assert parent.poll(120); msg = parent.recv(); assert msg == "rendered", msg
        parent.send("rendered")
        tile.halo_pull_peer(peer)
        tile.sync()
        assert parent.poll(120); msg = parent.recv(); assert msg == "pulled", msg
        parent.send("pulled")
    assert parent.poll(120)
    status, their_bad = parent.recv()
    p.join(30)
    assert status == "done", their_bad
    assert bad == 0 and their_bad == 0, (bad, their_bad)


@pytest.mark.parametrize("smaa,taa,rects", [
    (True, True, [(0, 72, 0, 40), (72, 144, 0, 40), (0, 72, 40, 96), (72, 144, 40, 96)]),
    (True, False, [(0, 60, 0, 96), (60, 144, 0, 96)]),
    (False, True, [(0, 144, 0, 50), (0, 144, 50, 96)]),
])
def test_temporal_upscalers_on_tiles_equal_unsharded(smaa, taa, rects):
    """The default pipeline's smaa_tu4x / taa_jasmine on tiles: with the upscaler planes enabled, a motion margin and halo
    pulls (which also carry the tone-mapped and TAA history of the ghost ring), every tile's part of the upscaled and the TAA
    image equals the full-frame render bit for bit under camera motion."""
    from bevy_hikari_b200 import plugin
    b = Bench("cornell", 144, 96, config="cornell_1080p", taa=plugin.TAA_JASMINE if taa else plugin.TAA_NONE,
              upscale_kind=plugin.UPSCALE_SMAA_TU4X if smaa else plugin.UPSCALE_FSR1, upscale_ratio=1.0)
    full = b.device()
    tiles = [b.device(r[2], r[3], r[0], r[1]) for r in rects]
    for t in tiles:
        t.set_motion_margin(16)
        t.enable_tile_upscalers()
    outputs = ([L.OUT_UPSCALED] if smaa else []) + ([L.OUT_TAA] if taa else [])
    k_of = {L.OUT_UPSCALED: 2, L.OUT_TAA: 2 if smaa else 1}
    for f in range(1, 9):
        inp = b.moving_inputs(f, step=(0.04, 0.01, -0.02))
        inp.temporal_upscalers = 1
        inp.fsr1 = 0        # not SMAA = "TAA only" here: FSR1 itself runs on full-frame contexts only (tests/test_gpu_zz_fsr.py)
        full.render_frame(inp)
        for t in tiles:
            t.render_frame(inp)
        for t in tiles:
            t.sync()
        for k in PLANES:
            whole = full.readback(k)
            for r, t in zip(rects, tiles):
                assert mismatch(t.readback(k), whole[r[2]:r[3], r[0]:r[1]]) == 0, (f, k)
        for k in outputs:
            whole, s = full.readback(k), k_of[k]
            for r, t in zip(rects, tiles):
                part = t.readback(k)
                assert part.shape[:2] == (s * (r[3] - r[2]), s * (r[1] - r[0]))
                assert mismatch(part, whole[s * r[2]:s * r[3], s * r[0]:s * r[1]]) == 0, (f, k, r)
        for t in tiles:
            for other in tiles:
                if other is not t:
                    t.halo_pull(other)
        for t in tiles:
            t.sync()


def test_tile_upscalers_need_their_planes_and_a_margin():
    from bevy_hikari_b200 import _ffi
    b = Bench("cornell", 64, 48, config="cornell_256")
    t = b.device(0, 24)
    inp = b.inputs(1)
    inp.temporal_upscalers = 1
    with pytest.raises(_ffi.HikariError, match="enable_tile_upscalers"):
        t.render_frame(inp)
    t.enable_tile_upscalers()
    with pytest.raises(_ffi.HikariError, match="motion margin"):
        t.render_frame(inp)                       # planes are there, the margin is not
    t.set_motion_margin(4)
    t.render_frame(inp)
    t.sync()
