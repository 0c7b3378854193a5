"""Frame assembly for tiled rendering (hk_set_frame_target): tiles store their tone-mapped pixels straight into one
full-frame buffer — on the same GPU, on a peer GPU of the same process, or in another process's allocation through CUDA
IPC — and the assembled frame is byte-identical to the unsharded render."""
import multiprocessing as mp

import numpy as np
import pytest

from bevy_hikari_b200 import _ffi
from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_gpu_parity import mismatch

pytestmark = pytest.mark.gpu

W, H = 160, 96
TILES = [(0, 88, 0, 96), (88, 160, 0, 96)]   # (col_begin, col_end, row_begin, row_end)


def test_tiles_assemble_into_one_frame():
    b = Bench("cornell", W, H, config="cornell_1080p")
    full = b.device()
    tiles = [b.device(t[2], t[3], t[0], t[1]) for t in TILES]
    frames = [full.frame_alloc()[0], full.frame_alloc()[0]]            # double-buffered, owned by the full context
    for f in range(1, 7):
        inp = b.inputs(f)      # static camera: tiles are bit-identical to the unsharded frame
        full.render_frame(inp)
        for t in tiles:
            t.set_frame_target(frames[f & 1], W)
            t.render_frame(inp)
        for t in tiles:
            t.sync()
        n_bad = mismatch(full.frame_read(frames[f & 1]), full.readback(L.OUT_TONE_MAPPED))
        assert n_bad == 0, (f, n_bad)
    # node-by-node path (unfused tone mapping kernel) writes the target too
    for t in tiles:
        t.set_frame_target(frames[0], W)
        t.prepass(inp); t.light(inp); t.post_process(inp); t.sync()
    full.prepass(inp); full.light(inp); full.post_process(inp)
    assert mismatch(full.frame_read(frames[0]), full.readback(L.OUT_TONE_MAPPED)) == 0
    # clearing the target stops the writes; a too-small pitch is refused
    tiles[0].set_frame_target(None)
    with pytest.raises(_ffi.HikariError, match="pitch"):
        tiles[0].set_frame_target(frames[0], W - 1)


def test_peer_gpu_frame_target_same_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    b = Bench("cornell", W, H, config="cornell_1080p")
    full = b.device()
    remote = plugin.HikariPlugin(W, H, 1, TILES[1][2], TILES[1][3], None, TILES[1][0], TILES[1][1])   # tile rendered on cuda:1
    remote.upload_scene(b.world)
    local = b.device(TILES[0][2], TILES[0][3], TILES[0][0], TILES[0][1])
    frame = full.frame_alloc()[0]                                      # lives on cuda:0
    for f in range(1, 5):
        inp = b.inputs(f)
        full.render_frame(inp)
        for t in (local, remote):
            t.set_frame_target(frame, W)
            t.render_frame(inp)
        local.sync(); remote.sync()
        assert mismatch(full.frame_read(frame), full.readback(L.OUT_TONE_MAPPED)) == 0, f


def _ipc_worker(handle, tile, frames, conn):
    try:
        _ffi.DEFAULT_FLAVOR = "exact"      # a spawned process does not run conftest.pytest_configure: same flavour as the parent's contexts
        b = Bench("cornell", W, H, config="cornell_1080p")
        dev = b.device(tile[2], tile[3], tile[0], tile[1])
        target = dev.frame_open(handle)
        for f in range(1, frames + 1):
            dev.set_frame_target(target, W)
            dev.render_frame(b.inputs(f))
        dev.sync()
        conn.send("ok")
    except Exception as e:   # pragma: no cover
        conn.send(repr(e))


def test_frame_target_across_processes_cuda_ipc():
    """The owner process allocates the frame and renders tile 0; a second process maps the frame through its IPC handle
    and renders tile 1 into it (both on cuda:0 here; across GPUs the same calls go over NVLink)."""
    from tests.conftest import needs_real_gpu
    needs_real_gpu()
    b = Bench("cornell", W, H, config="cornell_1080p")
    full = b.device()
    local = b.device(TILES[0][2], TILES[0][3], TILES[0][0], TILES[0][1])
    frame, handle = full.frame_alloc()
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    p = ctx.Process(target=_ipc_worker, args=(handle, TILES[1], 3, child), daemon=True)
    p.start()
    try:
        for f in range(1, 4):
            inp = b.inputs(f)
            full.render_frame(inp)
            local.set_frame_target(frame, W)
            local.render_frame(inp)
        local.sync()
        assert parent.poll(180), "worker did not answer"
        assert parent.recv() == "ok"
        p.join(30)
    finally:                          # the worker never outlives the test, whatever failed
        if p.is_alive():
            p.terminate()
        p.join(10)
    assert mismatch(full.frame_read(frame), full.readback(L.OUT_TONE_MAPPED)) == 0
