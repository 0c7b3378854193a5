"""Host-side scene preparation (H1/H2 of SURVEY.md 8(a)): the C++ host mirror (product) against the independent numpy
restatement in oracle/scene_build.py, bit-for-bit, plus structural invariants of the flattened skip-link BVH and of
the alias table, plus the reference's error behaviour (meshes without attributes are dropped, mod.rs:301-308)."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin, scenes
from oracle import scene_build


def fields_equal(a, b, dt):
    for n in dt.names:
        x, y = a[n], b[n]
        if x.dtype.names:
            fields_equal(x, y, x.dtype)
        else:
            assert x.tobytes() == y.tobytes(), n


def compare_builds(sc):
    w = sc.populate(plugin.World())
    pb = w.buffers()
    ob = scene_build.build_scene(sc.meshes, sc.inst_mesh, sc.inst_material, sc.inst_transform, sc.materials)
    for name, dt in L.SCENE_BUFFERS:
        assert len(pb[name]) == len(ob[name]), name
        fields_equal(pb[name], ob[name], dt)
    return pb


def walk_all_leaves(nodes):
    """Follow entry links unconditionally: visits every record once, in index order; returns leaf shape ids."""
    leaves, index, steps = [], 0, 0
    while index < len(nodes):
        n = nodes[index]
        if n["entry_index"] >= 0x80000000:
            leaves.append(int(n["entry_index"]) - 0x80000000)
            assert n["exit_index"] == index + 1
            index = int(n["exit_index"])
        else:
            assert n["entry_index"] == index + 1 and n["exit_index"] > index
            index = int(n["entry_index"])
        steps += 1
    assert steps == len(nodes)
    return leaves


def check_bvh(nodes, shape_min, shape_max):
    n_shapes = len(shape_min)
    assert len(nodes) == (3 * n_shapes - 2 if n_shapes > 1 else 1)
    leaves = walk_all_leaves(nodes)
    assert sorted(leaves) == list(range(n_shapes))
    # every navigator box contains the shapes of its subtree (records index+1 .. exit-1)
    for i, n in enumerate(nodes):
        if n["entry_index"] < 0x80000000:
            sub = nodes[i + 1:int(n["exit_index"])]
            ids = [int(e) - 0x80000000 for e in sub["entry_index"] if e >= 0x80000000]
            assert np.all(shape_min[ids] >= n["min"]) and np.all(shape_max[ids] <= n["max"])


def test_cornell_buffers_match_oracle_builder_bit_for_bit():
    pb = compare_builds(scenes.cornell())
    assert len(pb["primitives"]) == 32 and len(pb["vertices"]) == 78 and len(pb["instances"]) == 8
    assert len(pb["asset_nodes"]) == 80 and len(pb["instance_nodes"]) == 22
    assert len(pb["emissives"]) == 1 and len(pb["alias_table"]) == 2
    assert pb["emissives"][0]["instance"] == 4


def test_cornell_bvh_invariants():
    pb = scenes.cornell().populate(plugin.World()).buffers()
    tri = pb["primitives"]["vertices"]["position"]
    for inst in pb["instances"]:
        m = inst["mesh"]
        nodes = pb["asset_nodes"][m["node_offset"]:m["node_offset"] + m["node_count"]]
        n_prims = (m["node_count"] + 2) // 3
        t = tri[m["primitive"]:m["primitive"] + n_prims]
        check_bvh(nodes, t.min(axis=1), t.max(axis=1))
    check_bvh(pb["instance_nodes"], pb["instances"]["min"], pb["instances"]["max"])


@pytest.mark.parametrize("seed,n_tris", [(0, 1), (1, 2), (2, 7), (3, 64), (4, 500)])
def test_random_meshes_match_oracle_builder(seed, n_tris):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1, 1, (n_tris * 3, 3)).astype(np.float32)
    if seed == 3:   # many coincident centroids: exercises the split-in-half branch (split_axis_size < EPSILON)
        pos = np.tile(pos[:3], (n_tris, 1))
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (len(pos), 1))
    uv = rng.uniform(0, 1, (len(pos), 2)).astype(np.float32)
    idx = np.arange(len(pos), dtype=np.uint32)
    mats = np.zeros(1, L.MATERIAL)
    mats["base_color"], mats["emissive"] = (1, 1, 1, 1), (0.5, 0.25, 1.0, 1.0)
    for k in ("base_color_texture", "emissive_texture", "metallic_roughness_texture", "normal_map_texture", "occlusion_texture"):
        mats[k] = 0xFFFFFFFF
    xf = np.eye(4, dtype=np.float32)
    xf[0, 0], xf[1, 1], xf[3, :3] = 2.0, 0.5, (0.3, -0.2, 1.0)
    xf2 = np.eye(4, dtype=np.float32)
    xf2[3, :3] = (5.0, 0.0, 0.0)
    sc = scenes.SceneData([(pos, nrm, uv, idx)], mats, [], [0, 0], [0, 0], [xf.reshape(16), xf2.reshape(16)])
    pb = compare_builds(sc)
    tri = pb["primitives"]["vertices"]["position"]
    check_bvh(pb["asset_nodes"], tri.min(axis=1), tri.max(axis=1))
    # alias table: every entry's probability in [0,1]; expected pick frequency proportional to area
    at = pb["alias_table"][:n_tris]
    assert np.all(at["prob"] >= 0) and np.all(at["prob"] <= 1.0 + 1e-6) and np.all(at["index"] < n_tris)
    if n_tris > 1 and seed != 3:
        areas = 0.5 * np.linalg.norm(np.cross((tri[:, 1] - tri[:, 0]) * [2, .5, 1], (tri[:, 2] - tri[:, 0]) * [2, .5, 1]), axis=1)
        p = np.zeros(n_tris)
        for i, e in enumerate(at):
            p[i] += (1 - e["prob"]) / n_tris
            p[e["index"]] += e["prob"] / n_tris
        assert np.allclose(p, areas / areas.sum(), atol=2e-3)


def test_meshes_without_required_attributes_are_dropped_like_the_reference():
    w = plugin.World()
    pos = np.zeros((3, 3), np.float32); pos[1, 0] = pos[2, 1] = 1
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (3, 1))
    uv = np.zeros((3, 2), np.float32)
    good = w.add_mesh(pos, nrm, uv, np.arange(3, dtype=np.uint32))
    no_uv = w.add_mesh(pos, nrm, None, np.arange(3, dtype=np.uint32))
    no_normal = w.add_mesh(pos, None, uv, np.arange(3, dtype=np.uint32))
    lines = w.add_mesh(pos, nrm, uv, np.arange(3, dtype=np.uint32), topology=2)
    empty = w.add_mesh(pos, nrm, uv, np.zeros(0, np.uint32))
    strip = w.add_mesh(np.vstack([pos, [[1, 1, 0]]]).astype(np.float32), np.vstack([nrm, nrm[:1]]), np.vstack([uv, uv[:1]]),
                       np.arange(4, dtype=np.uint32), topology=1)
    m = np.zeros((), L.MATERIAL)
    w.add_material(m)
    for me in (good, no_uv, no_normal, lines, empty, strip):
        w.add_instance(me, 0, np.eye(4, dtype=np.float32).reshape(16))
    w.prepare()
    assert [w.mesh_error(i) for i in (good, no_uv, no_normal, lines, empty, strip)] == [0, 3, 2, 4, 5, 0]
    b = w.buffers()
    assert len(b["instances"]) == 2 and len(b["primitives"]) == 1 + 2      # list + 2-triangle strip
    # odd strip triangles swap the first two indices (mod.rs:436-440)
    assert list(b["primitives"][2]["vertices"]["index"]) == [2, 1, 3]


def test_city_scene_builds_and_matches_oracle_builder():
    sc = scenes.city()
    pb = compare_builds(sc)
    assert len(pb["instances"]) == 54 and len(pb["instance_nodes"]) == 3 * 54 - 2
    assert len(pb["primitives"]) == 19136
    assert len(pb["emissives"]) == 1


def test_large_mesh_builder_invariants():
    """bvh 0.7.1 restatement on 100 352 triangles: record count 3N - 2, every leaf reachable exactly once, navigator boxes
    contain their subtrees (checked on a sample), build time bounded."""
    import time
    from bevy_hikari_b200 import scenes
    t0 = time.time()
    w = scenes.terrain().populate(plugin.World())
    assert time.time() - t0 < 20.0
    bufs = w.buffers()
    inst = bufs["instances"][0]
    n_tris = 2 * 224 * 224
    nodes = bufs["asset_nodes"][inst["mesh"]["node_offset"]:inst["mesh"]["node_offset"] + inst["mesh"]["node_count"]]
    assert len(nodes) == 3 * n_tris - 2
    leaves = walk_all_leaves(nodes)
    assert sorted(leaves) == list(range(n_tris))
    prim = bufs["primitives"]["vertices"]["position"][inst["mesh"]["primitive"]:inst["mesh"]["primitive"] + n_tris]
    mn, mx = prim.min(axis=1), prim.max(axis=1)
    rng = np.random.default_rng(0)
    nav = np.flatnonzero(nodes["entry_index"] < 0x80000000)
    for i in rng.choice(nav, 200, replace=False):
        n = nodes[i]
        sub = nodes[i + 1:int(n["exit_index"])]
        ids = sub["entry_index"][sub["entry_index"] >= 0x80000000] - 0x80000000
        assert np.all(mn[ids] >= n["min"]) and np.all(mx[ids] <= n["max"])
