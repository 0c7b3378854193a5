"""Animated instances (the host half): the transform queue (transform.rs:31-44), the per-frame
rebuild of instances / TLAS / emissives / alias tables with the per-entity alias-table cache (instance.rs:352-437), and
the motion vectors of moving instances (prepass.wgsl:52,99) in the oracle.  CPU only."""
import numpy as np

from bevy_hikari_b200 import layout as L
from tests.conftest import Bench, cornell_animation, city_animation, rotation_y_about


def test_transform_queue_semantics():
    b = Bench("cornell", 16, 16, config="cornell_256")
    w = b.world
    base = w.buffers()["instances"]["model"].copy()
    assert w.previous_models().shape == (8, 16) and np.array_equal(w.previous_models(), base.reshape(-1, 16))   # no queue yet: previous == current
    m1 = rotation_y_about(0.3, (0, 0, 0), (0.1, 0, 0))
    m2 = rotation_y_about(0.6, (0, 0, 0), (0.2, 0, 0))
    w.set_instance_transform(6, m1)
    w.previous_transform_system()          # first sighting: queue = [m1, m1]
    w.prepare_instances()
    assert np.array_equal(w.buffers()["instances"]["model"][6].reshape(16), m1)
    assert np.array_equal(w.previous_models()[6], m1)
    w.set_instance_transform(6, m2)
    w.previous_transform_system()          # queue = [m2, m1]
    w.prepare_instances()
    assert np.array_equal(w.buffers()["instances"]["model"][6].reshape(16), m2)
    assert np.array_equal(w.previous_models()[6], m1)
    w.previous_transform_system()          # nothing moved this frame: queue = [m2, m2]
    w.prepare_instances()
    assert np.array_equal(w.previous_models()[6], m2)
    untouched = [i for i in range(8) if i != 6]
    assert np.array_equal(w.previous_models()[untouched], base.reshape(-1, 16)[untouched])


def test_instance_rebuild_matches_independent_builder():
    """After moving instances the host mirror's TLAS / instance records equal the numpy restatement of the same builder
    (oracle/scene_build.py) fed with the moved transforms."""
    from oracle import scene_build
    b = Bench("cornell", 16, 16, config="cornell_256")
    an = cornell_animation(b)
    for f in (1, 2, 3):
        w = an.step(f)
    got = w.buffers()
    sc = b.scene
    xf = [np.array(t, np.float32) for t in sc.inst_transform]
    for i, fn in an.tracks.items():
        xf[i] = an.compose(fn(3), an.base[i])
    ref = scene_build.build_scene(sc.meshes, sc.inst_mesh, sc.inst_material, xf, sc.materials)
    from tests.test_scene_build import fields_equal
    # (the alias table itself is exempt: it is cached per entity, see test_alias_table_cache)
    for name, dt in L.SCENE_BUFFERS:
        if name in ("instances", "instance_nodes", "emissive_nodes", "emissives"):
            assert len(got[name]) == len(ref[name]), name
            fields_equal(got[name], ref[name], dt)


def test_alias_table_cache():
    """instance.rs:385-397: the table of an emissive entity is rebuilt only when its scale moves by more than 0.01."""
    b = Bench("city", 16, 16, config="city_4k")
    w = b.world
    first = w.buffers()["alias_table"].copy()
    base = np.array(b.scene.inst_transform[1], np.float32)
    from bevy_hikari_b200 import scenes
    w.set_instance_transform(1, scenes._compose(rotation_y_about(0.7, (0, 1, 0)), base))
    w.previous_transform_system(); w.prepare_instances()
    rotated = w.buffers()["alias_table"]
    assert rotated.tobytes() == first.tobytes()                       # cache hit: rotation keeps the scale
    from bevy_hikari_b200 import plugin
    spawned_rotated = scenes.city()          # an entity first seen with the rotated transform gets a table built from it
    spawned_rotated.inst_transform[1] = scenes._compose(rotation_y_about(0.7, (0, 1, 0)), base)
    fresh = spawned_rotated.populate(plugin.World()).buffers()["alias_table"]
    assert fresh.tobytes() != first.tobytes() and len(fresh) == len(first)   # same areas up to rounding, a different table
    scale = np.diag([1.5, 1.0, 1.0, 1.0]).astype(np.float32).reshape(16)
    w.set_instance_transform(1, scenes._compose(scale, base))
    w.previous_transform_system(); w.prepare_instances()
    assert w.buffers()["alias_table"].tobytes() != first.tobytes()    # scale changed by 0.5 > 0.01: rebuilt
    assert len(w.buffers()["alias_table"]) == len(first)


def project(view_proj, p):
    clip = np.concatenate([p, np.ones((len(p), 1))], axis=1) @ view_proj.reshape(4, 4).astype(np.float64)
    ndc = clip[:, :2] / clip[:, 3:4]
    uv = (ndc + 1.0) * 0.5
    uv[:, 1] = 1.0 - uv[:, 1]
    return uv


def test_motion_vectors_of_moving_instances():
    """velocity = uv(view_proj * p) - uv(previous_view_proj * previous_model * model^-1 * p) on moved instances, exactly 0 with a
    static camera elsewhere; checked against float64 numpy."""
    b = Bench("cornell", 96, 72, config="cornell_256")
    orc = b.oracle()
    an = cornell_animation(b)
    for f in range(1, 4):
        w = an.step(f)
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f)
        orc.prepass(inp)
    vel = orc.readback(L.OUT_GBUFFER_VELOCITY_UV)[..., :2].astype(np.float64)
    pos = orc.readback(L.OUT_GBUFFER_POSITION).astype(np.float64)
    inst = np.floor(orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)[..., 0]).astype(int)
    hit = pos[..., 3] > 0
    moved = np.isin(inst, (4, 6)) & hit
    assert moved.sum() > 150
    assert (vel[hit & ~moved] == 0).all()
    models = w.buffers()["instances"]["model"].reshape(-1, 4, 4).astype(np.float64)
    prev = w.previous_models().reshape(-1, 4, 4).astype(np.float64)
    vp = np.array(inp.view.view_proj[:], np.float64)
    pvp = np.array(inp.previous_view.view_proj[:], np.float64)
    for i in (4, 6):
        sel = moved & (inst == i)
        p = pos[sel][:, :3]
        local = np.concatenate([p, np.ones((len(p), 1))], axis=1) @ np.linalg.inv(models[i])
        previous_world = (local @ prev[i])[:, :3]
        expect = project(vp, p) - project(pvp, previous_world)
        assert np.abs(expect).max() > 1e-3
        assert np.abs(vel[sel] - expect).max() < 2e-5, i


def test_update_instances_equals_full_upload_in_oracle():
    b = Bench("city", 64, 36, config="city_4k")
    a, c = b.oracle(), b.oracle()
    an = city_animation(b)
    for f in range(1, 4):
        w = an.step(f)
        a.update_instances_desc(w.scene_desc())
        c.upload_scene_desc(w.scene_desc())
        inp = b.inputs(f)
        a.render_frame(inp); c.render_frame(inp)
        for k in (L.OUT_TONE_MAPPED, L.OUT_GBUFFER_VELOCITY_UV, L.OUT_RESERVOIR_0 + 9):
            assert np.array_equal(np.ascontiguousarray(a.readback(k)).view(np.uint8), np.ascontiguousarray(c.readback(k)).view(np.uint8)), (f, k)


def test_visibility_changes_rebuild_the_instance_list():
    """instance.rs:357: invisible entities are dropped before the TLAS is built; instance ids are ranks among the visible ones"""
    b = Bench("cornell", 48, 36, config="cornell_256")
    w = b.world
    n = len(w.buffers()["instances"])
    w.set_instance_visible(6, False)
    w.previous_transform_system(); w.prepare_instances()
    bufs = w.buffers()
    assert len(bufs["instances"]) == n - 1 and len(bufs["instance_nodes"]) == 3 * (n - 1) - 2
    assert w.previous_models().shape == (n - 1, 16)
    orc = b.oracle()
    orc.update_instances_desc(w.scene_desc())
    orc.prepass(b.inputs(1))
    ids = np.floor(orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)[..., 0]).astype(int)
    materials = np.floor(orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)[..., 1]).astype(int)
    hit = orc.readback(L.OUT_GBUFFER_POSITION)[..., 3] > 0
    assert ids[hit].max() == n - 2 and 6 not in set(materials[hit])        # the short box (material 6) is gone, the tall box moved up a rank
    w.set_instance_visible(6, True)
    w.previous_transform_system(); w.prepare_instances()
    assert len(w.buffers()["instances"]) == n
