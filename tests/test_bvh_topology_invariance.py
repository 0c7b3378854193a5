"""How much of the result hangs on the one dependency that could only be restated from memory — `bvh = "= 0.7.1"` (`BVH::build`:
recursive SAH with 6 buckets)?  The traversal is order-preserving on purpose: visit order decides
equal-distance ties, which occluder an any-hit ray reports and the order emissive leaves are streamed in.  This test measures what
that amounts to: the same scenes are built with DIFFERENT valid BVH topologies (2 and 12 SAH buckets instead of 6, and every inner node's children flattened in
the opposite order, through the numpy builder) and rendered by the oracle for eight frames with the full pipeline.  Every image — G-buffer, the three radiance planes,
denoised output, tone-mapped frame — is bit-identical across topologies; the only records that differ are temporal reservoirs that
stored the position of an occluder for a zero-radiance sample (any-hit order), which no image depends on.  So a wrong guess about the
crate's bucket count or split rule would not change a pixel in these scenes.  With SEVERAL emissive instances in range the order
of the emissive leaves is observable — the streaming 1 / count pick (light.wgsl:628-656) consumes its random numbers in leaf order —
so simple.rs (two emissive spheres) gets a different, statistically equivalent sample sequence; the second test states exactly that.
CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from oracle import oracle, scene_build
from tests.conftest import Bench

IMAGES = [L.OUT_TONE_MAPPED, L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_INSTANCE_MATERIAL, L.OUT_GBUFFER_VELOCITY_UV,
          L.OUT_RENDER_DIRECT, L.OUT_RENDER_EMISSIVE, L.OUT_RENDER_INDIRECT, L.OUT_ALBEDO]


def render(scene, config, size, frames, buckets, swap_children=False):
    previous = scene_build.NUM_BUCKETS
    scene_build.NUM_BUCKETS, scene_build.SWAP_CHILDREN = buckets, swap_children
    try:
        b = Bench(scene, size[0], size[1], config=config)
        sc = b.scene
        bufs = scene_build.build_scene(sc.meshes, sc.inst_mesh, sc.inst_material, sc.inst_transform, sc.materials)
    finally:
        scene_build.NUM_BUCKETS, scene_build.SWAP_CHILDREN = previous, False
    orc = oracle.Oracle(size[0], size[1], plugin.load_noise())
    orc.upload_scene_desc(plugin.scene_desc_from_buffers(bufs, sc.textures))
    out = []
    for f in range(1, frames + 1):
        orc.render_frame(b.moving_inputs(f) if f > 4 else b.inputs(f))
        out.append({k: orc.readback(k).copy() for k in IMAGES + [L.OUT_RESERVOIR_0 + r for r in range(10)]})
    return out, bufs


@pytest.mark.parametrize("scene,config,size", [("cornell", "cornell_1080p", (80, 64)), ("minimal", "cornell_1080p", (80, 56))])
def test_images_do_not_depend_on_the_bvh_topology(scene, config, size):
    frames = 8
    reference, ref_bufs = render(scene, config, size, frames, 6)
    for buckets, swap in ((2, False), (12, True), (6, True)):
        other, bufs = render(scene, config, size, frames, buckets, swap)
        assert len(bufs["asset_nodes"]) == len(ref_bufs["asset_nodes"])
        if swap:                    # really a different walk: every inner node visits its children in the opposite order
            assert bufs["asset_nodes"].tobytes() != ref_bufs["asset_nodes"].tobytes() and bufs["instance_nodes"].tobytes() != ref_bufs["instance_nodes"].tobytes()
            assert bufs["primitives"].tobytes() == ref_bufs["primitives"].tobytes()
        reservoir_records = 0
        for f in range(frames):
            for k in IMAGES:
                assert reference[f][k].tobytes() == other[f][k].tobytes(), (scene, buckets, f + 1, k)
            for r in range(10):
                a, b_ = reference[f][L.OUT_RESERVOIR_0 + r], other[f][L.OUT_RESERVOIR_0 + r]
                differ = (a.view(np.uint8).reshape(-1, 64) != b_.view(np.uint8).reshape(-1, 64)).any(1)
                if differ.any():
                    # only the stored sample position / normal may differ, and only for samples that carry no radiance
                    ar, br = a.reshape(-1)[differ], b_.reshape(-1)[differ]
                    for field in ("radiance", "random", "visible_position", "visible_normal", "reservoir"):
                        assert np.array_equal(ar[field], br[field]), (scene, buckets, f + 1, r, field)
                    assert not (ar["radiance"][:, 0] & 0x7FFF7FFF).any() and not (ar["radiance"][:, 1] & 0x7FFF).any()   # r, g, b == 0
                    reservoir_records += int(differ.sum())
        assert reservoir_records < 0.01 * frames * 10 * size[0] * size[1]


def test_several_emissives_make_the_leaf_order_observable_but_not_the_estimate():
    """simple.rs: geometry planes are invariant; the light samples differ with the emissive-BVH leaf order, the time average does not"""
    size, frames = (96, 60), 8
    reference, _ = render("simple", "cornell_1080p", size, frames, 6)
    other, _ = render("simple", "cornell_1080p", size, frames, 6, swap_children=True)
    for f in range(frames):
        for k in (L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_INSTANCE_MATERIAL, L.OUT_ALBEDO, L.OUT_RENDER_DIRECT):
            assert reference[f][k].tobytes() == other[f][k].tobytes(), (f + 1, k)       # the sun pass does not walk the emissive BVH
    a = np.mean([reference[f][L.OUT_RENDER_EMISSIVE].astype(np.float32)[..., :3] for f in range(frames)], axis=0)
    b_ = np.mean([other[f][L.OUT_RENDER_EMISSIVE].astype(np.float32)[..., :3] for f in range(frames)], axis=0)
    assert (reference[0][L.OUT_RENDER_EMISSIVE] != other[0][L.OUT_RENDER_EMISSIVE]).any()                 # observable ...
    assert abs(float(a.mean()) - float(b_.mean())) <= 0.05 * float(a.mean())                            # ... but the same estimator
