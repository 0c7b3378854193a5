"""SURVEY 8(f) rank 2 — the per-frame half of the scene rebuilt on the device (hk_scene_update_transforms, csrc/kernels_scene.cu):
instances' world AABBs and matrices, the TLAS (bvh 0.7.1's bucketed SAH build + flatten_custom, one warp per tree node), the emissives'
bounding spheres / surface areas and the emissive BVH, from one model matrix per instance.

Held against the host mirror's prepare_instances (host/hikari.cpp = the reference's CPU path, instance.rs:352-437; itself held against an
independent numpy builder in tests/test_scene_build.py and tests/test_dynamic_scene.py): EVERY RECORD of every rebuilt buffer is
bit-identical, on the example scenes under animation and on random instance soups that exercise the builder's corners (coincident
centres -> the half split, flat boxes -> zero surface areas, a single instance, many instances per warp pass); and frames rendered
after a device update equal the oracle's frames after the host update, bit for bit."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench, cornell_animation, city_animation, rotation_y_about
from tests.test_gpu_parity import ALL_PLANES, compare_all

pytestmark = pytest.mark.gpu

BUFFERS = [(L.SCENE_INSTANCES, "instances"), (L.SCENE_INSTANCE_NODES, "instance_nodes"), (L.SCENE_EMISSIVES, "emissives"),
           (L.SCENE_EMISSIVE_NODES, "emissive_nodes")]


def assert_scene_equals_host(dev, world, what):
    host = world.buffers()
    for which, name in BUFFERS:
        got, want = dev.scene_readback(which), host[name]
        assert got.shape == want.shape, (what, name, got.shape, want.shape)
        if len(got) == 0:
            continue
        for field in got.dtype.names:       # field by field, bitwise (numpy's copies of padded records do not carry the padding bytes)
            g = np.ascontiguousarray(got[field]).view(np.uint32).reshape(len(got), -1)
            h = np.ascontiguousarray(want[field]).view(np.uint32).reshape(len(want), -1)
            bad = np.nonzero((g != h).any(axis=1))[0]
            if len(bad):
                raise AssertionError(f"{what}: {name}.{field} differs in {len(bad)} of {len(got)} records, first {bad[0]}:\n device {got[bad[0]]}\n host   {want[bad[0]]}")


@pytest.mark.parametrize("scene,config,animation", [("cornell", "cornell_1080p", cornell_animation), ("city", "city_4k", city_animation)])
def test_device_rebuild_equals_the_host_mirror_record_for_record(scene, config, animation):
    b = Bench(scene, 48, 32, config=config)
    dev = b.device()
    anim = animation(b)
    for f in range(1, 7):
        w = anim.bench.world
        for i, track in anim.tracks.items():
            w.set_instance_transform(i, anim.compose(track(f), anim.base[i]))
        w.previous_transform_system()
        assert dev.update_transforms(w), "only transforms changed: the device path must have been taken"
        w.prepare_instances()                       # the reference's CPU path on the same transforms, for comparison only
        assert_scene_equals_host(dev, w, f"{scene} frame {f}")
        prev = dev.scene_readback(L.SCENE_PREVIOUS_MODELS)
        assert np.array_equal(prev.view(np.uint32), w.previous_models().view(np.uint32))
        moved = dev.scene_readback(L.SCENE_INSTANCE_MOVED)
        assert np.array_equal(moved != 0, (prev.view(np.uint32) != host_models(w).view(np.uint32)).any(axis=1))
    dev.close()


def host_models(world):
    return world.buffers()["instances"]["model"].reshape(-1, 16)


@pytest.mark.parametrize("scene,config,animation,size", [("cornell", "cornell_1080p", cornell_animation, (96, 64)), ("city", "city_4k", city_animation, (96, 54))])
def test_frames_after_device_updates_equal_the_oracle(scene, config, animation, size):
    b = Bench(scene, size[0], size[1], config=config)
    dev, orc = b.device(), b.oracle()
    anim = animation(b)
    for f in range(1, 8):
        w = anim.bench.world
        for i, track in anim.tracks.items():
            w.set_instance_transform(i, anim.compose(track(f), anim.base[i]))
        w.previous_transform_system()
        assert dev.update_transforms(w)
        w.prepare_instances()
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f)
        dev.render_frame(inp); orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES, f)
    dev.close()


def soup_world(rng, n, kind):
    """n unit cubes (a few of them emissive) under random transforms; `kind` picks the corner of the builder that is exercised"""
    from bevy_hikari_b200 import scenes
    w = plugin.World()
    pos, nrm, uv, idx = scenes.cube_mesh() if hasattr(scenes, "cube_mesh") else unit_cube()
    mesh = w.add_mesh(pos, nrm, uv, idx)
    flat = w.add_mesh(*unit_quad())
    plain = w.add_material(material())
    glow = w.add_material(material(emissive=(1.0, 0.8, 0.6, 0.5)))
    for i in range(n):
        t = np.eye(4, dtype=np.float32)
        if kind == "coincident":
            centre = np.zeros(3) if i % 3 else rng.uniform(-1e-7, 1e-7, 3)      # centres within EPSILON -> the half split
        elif kind == "line":
            centre = np.array([rng.uniform(-50, 50), 0.0, 0.0])
        else:
            centre = rng.uniform(-20, 20, 3)
        s = rng.uniform(0.2, 2.0, 3) if kind != "coincident" else np.full(3, 1.0 + 0.25 * (i % 4))
        ang = rng.uniform(0, 6.28)
        c, sn = np.cos(ang), np.sin(ang)
        rot = np.array([[c, 0, sn], [0, 1, 0], [-sn, 0, c]])
        t[:3, :3] = (rot * s).astype(np.float32)
        t[:3, 3] = centre.astype(np.float32)
        m = flat if (kind == "flat" and i % 2 == 0) else mesh
        w.add_instance(m, glow if i % 7 == 3 else plain, t.T.reshape(16).copy())       # column-major
    w.prepare()
    return w


def unit_cube():
    p, n, u, idx = [], [], [], []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            a, b_ = (axis + 1) % 3, (axis + 2) % 3
            base = len(p)
            for da, db in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                v = [0.0, 0.0, 0.0]; v[axis] = 0.5 * sgn; v[a] = 0.5 * da; v[b_] = 0.5 * db
                nn = [0.0, 0.0, 0.0]; nn[axis] = sgn
                p.append(v); n.append(nn); u.append([(da + 1) / 2, (db + 1) / 2])
            idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.array(p, np.float32), np.array(n, np.float32), np.array(u, np.float32), np.array(idx, np.uint32)


def unit_quad():
    p = np.array([[-0.5, 0, -0.5], [0.5, 0, -0.5], [0.5, 0, 0.5], [-0.5, 0, 0.5]], np.float32)
    n = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    u = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    return p, n, u, np.array([0, 1, 2, 0, 2, 3], np.uint32)


def material(emissive=(0.0, 0.0, 0.0, 1.0)):
    m = np.zeros(1, L.MATERIAL)
    m["base_color"] = (0.8, 0.8, 0.8, 1.0)
    m["emissive"] = emissive
    m["perceptual_roughness"] = 0.5; m["metallic"] = 0.0; m["reflectance"] = 0.5
    for k in ("base_color_texture", "emissive_texture", "metallic_roughness_texture", "normal_map_texture", "occlusion_texture"):
        m[k] = 0xFFFFFFFF
    return m


@pytest.mark.parametrize("n,kind,seed", [(1, "random", 1), (2, "random", 2), (3, "coincident", 3), (33, "random", 4), (70, "coincident", 5),
                                         (200, "random", 6), (97, "line", 7), (64, "flat", 8), (600, "random", 9)])
def test_random_instance_soups_rebuild_bit_identically(n, kind, seed):
    rng = np.random.default_rng(seed)
    w = soup_world(rng, n, kind)
    dev = plugin.HikariPlugin(32, 24, 0, 0, None, None, 0, None)
    dev.upload_scene(w)
    for step in range(3):
        for i in range(n):
            if rng.random() < 0.6:
                cur = w.buffers()["instances"]["model"][i].reshape(4, 4).T.astype(np.float64)     # row-major view of the column-major model
                ang = rng.uniform(-0.3, 0.3)
                c, s = np.cos(ang), np.sin(ang)
                rot = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
                cur = rot @ cur
                if kind not in ("coincident",):
                    cur[:3, 3] += rng.uniform(-1.0, 1.0, 3) * (0 if kind == "line" else 1) + (np.array([rng.uniform(-2, 2), 0, 0]) if kind == "line" else 0)
                w.set_instance_transform(i, cur.T.astype(np.float32).reshape(16).copy())
        w.previous_transform_system()
        assert dev.update_transforms(w), (n, kind, step)
        w.prepare_instances()
        assert_scene_equals_host(dev, w, f"{kind} n={n} step {step}")
    dev.close()


def test_set_changes_and_rescaled_lights_take_the_host_path():
    b = Bench("cornell", 32, 24, config="cornell_1080p")
    dev = b.device()
    w = b.world
    assert dev.update_transforms(w)                       # nothing changed at all: still the device path
    w.set_instance_visible(6, False)
    assert not dev.update_transforms(w)                   # the kept set changed -> prepare_instances + hk_scene_update_instances
    assert len(dev.scene_readback(L.SCENE_INSTANCES)) == len(w.buffers()["instances"])
    assert dev.update_transforms(w)                       # ... and from that state on the device path again
    # an emissive instance scaled by more than 0.01: its alias table must be rebuilt (instance.rs:385-397)
    em = int(w.buffers()["emissives"]["instance"][0])
    model = w.buffers()["instances"]["model"][em].reshape(4, 4).copy()
    model[:3, :3] *= np.float32(1.05)
    w.set_instance_transform(kept_entity(w, em, hidden=6), model.reshape(16))
    assert not dev.update_transforms(w)
    assert_scene_equals_host(dev, w, "after the host path")
    dev.close()


def kept_entity(world, buffer_index, hidden):
    """entity id of the instance at `buffer_index` of the instance buffer when entity `hidden` is invisible"""
    return buffer_index if buffer_index < hidden else buffer_index + 1


def test_wrong_instance_count_is_refused():
    from bevy_hikari_b200._ffi import HikariError
    b = Bench("cornell", 32, 24, config="cornell_1080p")
    dev = b.device()
    models, prev, aabbs = b.world.prepare_instance_transforms()
    with pytest.raises(HikariError):
        dev.update_transforms_arrays(models[:-1], aabbs[:-1])
    dev.update_transforms_arrays(models, aabbs)           # previous = NULL: the models the records held
    moved = dev.scene_readback(L.SCENE_INSTANCE_MOVED)
    assert not moved.any()
    dev.close()
