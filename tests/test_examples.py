"""The reference's shape-only examples restated as scenes (examples/minimal.rs, examples/simple.rs): host builder vs the
independent numpy builder, and what the oracle renders for them.  simple.rs is the only scene with MORE THAN ONE emissive
instance (emissive BVH of 4 records, two alias tables) — light selection must reach both.  CPU only."""
import numpy as np

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin, scenes
from tests.conftest import Bench
from tests.test_oracle import brute_force, world_triangles
from tests.test_scene_build import check_bvh, compare_builds

SETTINGS = dict(taa=plugin.TAA_NONE, upscale_ratio=1.0)


def test_example_scenes_builders_agree():
    for name in ("minimal", "simple"):
        pb = compare_builds(scenes.SCENE_BUILDERS[name]())
        inst = pb["instances"]
        check_bvh(pb["instance_nodes"], inst["min"], inst["max"])
    # simple.rs: two emissive spheres -> emissive BVH over two bounding spheres, two alias tables back to back
    em = pb["emissives"]
    assert len(em) == 2 and len(pb["emissive_nodes"]) == 4
    assert list(em["instance"]) == [6, 7]
    assert list(em["alias_table_count"]) == [1224, 1224] and list(em["alias_table_offset"]) == [0, 1224]
    assert np.allclose(em["position"], [[2.0, 1.0, 0.0], [-2.0, 1.0, 0.0]], atol=1e-6)
    # radius = half diagonal + sqrt(255 * alpha * |rgb|) (instance.rs:411-413)
    expect = 0.5 * np.sqrt(3.0) + np.sqrt(255.0 * np.array([0.5, 0.1]) * np.sqrt(3.0))
    assert np.allclose(em["radius"], expect, rtol=1e-5)
    # a box scaled (8,1,8) keeps bevy's 12 triangles; the unit cube mesh is shared by five instances
    assert len({tuple(i["mesh"].tolist()) for i in inst[[0, 2, 3, 4, 5]]}) == 1


def gbuffer_matches_brute_force(b, eye):
    orc = b.oracle()
    orc.prepass(b.inputs(1))
    pos = orc.readback(L.OUT_GBUFFER_POSITION)
    im = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)
    covered = pos[..., 3] > 0
    tris, owner = world_triangles(b.world.buffers())
    ys, xs = np.nonzero(covered)
    sel = np.random.default_rng(0).choice(len(ys), 250, replace=False)
    target = pos[ys[sel], xs[sel], :3].astype(np.float64)
    d = target - eye
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t_ref, i_ref = brute_force(tris, owner, np.tile(eye, (len(sel), 1)), d, np.full(len(sel), -1))
    got = np.floor(im[ys[sel], xs[sel], 0]).astype(int)
    assert (got == i_ref).mean() > 0.98
    assert np.allclose(np.linalg.norm(target - eye, axis=1), t_ref, rtol=1e-4)
    return covered


def test_minimal_example():
    b = Bench("minimal", 96, 64, **SETTINGS)
    covered = gbuffer_matches_brute_force(b, np.array([-2.0, 2.5, 5.0]))
    assert 0.3 < covered.mean() < 0.6
    orc = b.oracle()
    for f in range(1, 6):
        orc.render_frame(b.inputs(f))
    direct = orc.readback(L.OUT_RENDER_DIRECT).astype(np.float32)[..., :3].sum(axis=2)
    inst = np.floor(orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)[..., 0]).astype(int)
    ground = covered & (inst == 0)
    # the cube shadows part of the plane (sun from (+x, +y, +z) side): both lit and unlit ground pixels exist
    assert (direct[ground] > 0.05).mean() > 0.3 and (direct[ground] == 0).mean() > 0.03
    assert not orc.readback(L.OUT_RENDER_EMISSIVE).astype(np.float32)[..., :3].any()     # no emissive instance in this scene
    tm = orc.readback(L.OUT_TONE_MAPPED).astype(np.float32)
    assert np.isfinite(tm).all() and np.allclose(tm[~covered][:, :3], 0.4, atol=2e-3)    # clear_color rgb(0.4,0.4,0.4)


def test_simple_example_samples_both_emissives():
    b = Bench("simple", 128, 80, emissive_spatial_reuse=1, indirect_bounces=2, **SETTINGS)
    gbuffer_matches_brute_force(b, np.array([-10.0, 2.5, 20.0]))
    orc = b.oracle()
    for f in range(1, 7):
        orc.render_frame(b.inputs(f))
    res = orc.readback(L.OUT_RESERVOIR_0 + 2 + (6 & 1))     # emissive temporal reservoir written this frame
    res = res if res["reservoir"].any() else orc.readback(L.OUT_RESERVOIR_0 + 2 + 1 - (6 & 1))
    sample = res["sample_position"][..., :3].reshape(-1, 3)
    count = res["reservoir"].reshape(-1, 2)
    lit = np.abs(sample).sum(axis=1) > 0
    near_a = np.linalg.norm(sample[lit] - np.array([2.0, 1.0, 0.0]), axis=1) < 0.6
    near_b = np.linalg.norm(sample[lit] - np.array([-2.0, 1.0, 0.0]), axis=1) < 0.6
    assert lit.sum() > 1000 and count.any()
    # the sphere with emissive alpha 0.5 wins far more often than the one with 0.1, but both are reached through the emissive BVH
    assert near_a.mean() > 0.15 and near_b.sum() > 50 and near_a.sum() > 3 * near_b.sum(), (near_a.sum(), near_b.sum())
    emissive = orc.readback(L.OUT_RENDER_EMISSIVE).astype(np.float32)[..., :3]
    assert emissive.mean() > 0.01 and np.isfinite(emissive).all()


def test_town_scene_of_config_3():
    """examples/scene.rs (BASELINE configs[2]): assets/models/scene.gltf — the largest BLAS set of the configs (121 666
    primitives, 364 826 asset nodes) — host builder against the numpy builder, counts against the asset, and the oracle's
    G-buffer against brute force over all world triangles."""
    sc = scenes.town()
    pb = compare_builds(sc)
    inst = pb["instances"]
    check_bvh(pb["instance_nodes"], inst["min"], inst["max"])
    assert len(inst) == 86 and len(pb["instance_nodes"]) == 3 * 86 - 2
    assert len(pb["primitives"]) == 120440 + 1224 + 2 and len(pb["asset_nodes"]) == 3 * len(pb["primitives"]) - 2 * 86
    assert len(pb["materials"]) == 68 and len(pb["emissives"]) == 1 and int(pb["emissives"]["instance"][0]) == 1
    assert np.allclose(pb["emissives"]["position"][0], [2.0, 2.0, 0.0], atol=1e-6)
    b = Bench("town", 96, 54, config="scene_1080p")
    covered = gbuffer_matches_brute_force(b, np.array([-20.0, 10.0, 20.0]))
    assert covered.mean() > 0.9            # the 10 km ground plane fills the frame below the horizon
    orc = b.oracle()
    for f in range(1, 4):
        orc.render_frame(b.inputs(f))
    tm = orc.readback(L.OUT_TONE_MAPPED).astype(np.float32)
    assert np.isfinite(tm).all() and 0.05 < tm[..., :3].mean() < 0.6
    direct = orc.readback(L.OUT_RENDER_DIRECT).astype(np.float32)[..., :3].sum(axis=2)
    assert (direct[covered] > 0.05).mean() > 0.2 and (direct[covered] == 0).mean() > 0.05    # sun-lit and shadowed parts
