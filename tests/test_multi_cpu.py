"""N > 1 host logic on CPU (gloo, world_size 2): the row-band partition and the all-gather assembly bench.py uses.
The CUDA side of sharding (ghost rows) is covered on the GPU by test_gpu_parity.py::test_row_bands_equal_unsharded."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import Bench


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, width, full_bytes, out_dir):
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, x1, r0, r1 = bench.tile(width, height, rank, world)
    full = torch.frombuffer(bytearray(full_bytes), dtype=torch.float16).reshape(height, width, 4)
    tile = full[r0:r1, x0:x1].contiguous().reshape(-1)       # what hk_get_output(HK_OUT_TONE_MAPPED) holds on this rank
    frame = torch.empty(world * tile.numel(), dtype=torch.float16)
    dist.all_gather_into_tensor(frame, tile)                  # the one collective of the path
    # the gathered buffer is tile-major: [rank][row][col]; re-assemble and compare with the unsharded frame
    tiles = frame.reshape(world, r1 - r0, x1 - x0, 4)
    rebuilt = torch.empty_like(full)
    for r in range(world):
        a0, a1, b0, b1 = bench.tile(width, height, r, world)
        rebuilt[b0:b1, a0:a1] = tiles[r]
    ok = torch.equal(rebuilt.view(torch.int16), full.view(torch.int16))
    t = torch.tensor([1.0 if ok else 0.0, float((r1 - r0) * (x1 - x0))])
    dist.all_reduce(t)
    if rank == 0:
        np.save(os.path.join(out_dir, "result.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_band_partition_and_all_gather_reassemble_the_frame(tmp_path):
    import bench
    # partition properties
    for w, h, n in ((1920, 1080, 1), (1920, 1080, 2), (3840, 2160, 4), (7680, 4320, 8), (1920, 1080, 8), (1920, 1080, 4)):
        tiles = [bench.tile(w, h, r, n) for r in range(n)]
        cover = np.zeros((h, w), np.uint8)                    # every pixel owned exactly once
        for x0, x1, y0, y1 in tiles:
            cover[y0:y1, x0:x1] += 1
        assert (cover == 1).all()
        assert len({(t[1] - t[0], t[3] - t[2]) for t in tiles}) == 1     # equal contributions for the all-gather
    assert bench.tile_grid(1920, 1080, 2) == (2, 1) and bench.tile_grid(1920, 1080, 8) == (4, 2)
    with pytest.raises(AssertionError):
        bench.tile(1920, 1080, 0, 7)
    # a real frame from the oracle, split in two bands, reassembled over gloo
    b = Bench("cornell", 48, 64, config="cornell_256")
    orc = b.oracle(threads=2)
    orc.render_frame(b.inputs(1))
    from bevy_hikari_b200 import layout as L
    full = np.ascontiguousarray(orc.readback(L.OUT_TONE_MAPPED))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 64, 48, full.tobytes(), str(tmp_path)), nprocs=2, join=True)
    res = np.load(tmp_path / "result.npy")
    assert res[0] == 2.0 and res[1] == 2 * 24 * 64.0
