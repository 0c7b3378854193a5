"""The image-exact traversal mode (hk_set_tuning(HK_TUNE_WIDE_TRAVERSAL), csrc/hk_wide.cuh): 4-wide trees walked front to back with a
stack, the reference's box / triangle arithmetic and tie rule.  Contract stated in include/hikari_b200.h and held here against the CPU
oracle (= the reference's fixed-order walk):

  * closest-hit rays: instance, primitive, distance, u, v bit-identical except for rays whose two nearest hits tie within the
    rounding of a box test — fewer than 1e-4 of the rays (measured: none in these dumps);
  * shadow rays (early_distance > 0): occluded-or-not identical for every ray; WHICH occluder is reported may differ;
  * frames: G-buffer (ids, positions, normals, velocities) and albedo bit-identical; the three radiance planes, their variances and
    the tone-mapped image identical in all but < 1e-4 of the pixels per frame over a free-running sequence; the reservoir records
    differ only where a zero-radiance sample stores the position of an occluder (not compared).

On the device this runs on the exact flavour (tests/conftest.py), so everything that is not the walk is bit-exact with the oracle."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_gpu_parity import mismatch, random_rays

pytestmark = pytest.mark.gpu

GBUFFER = [L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_DEPTH_GRADIENT, L.OUT_GBUFFER_INSTANCE_MATERIAL,
           L.OUT_GBUFFER_VELOCITY_UV, L.OUT_ALBEDO]
IMAGES = [L.OUT_RENDER_DIRECT, L.OUT_RENDER_EMISSIVE, L.OUT_RENDER_INDIRECT, L.OUT_VARIANCE_DIRECT, L.OUT_VARIANCE_EMISSIVE,
          L.OUT_VARIANCE_INDIRECT, L.OUT_TONE_MAPPED]


def wide_device(b, **kw):
    dev = b.device(**kw)
    dev.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 3)          # every ray, whatever the scene's size
    st = dev.stats()
    assert st.wide_traversal == 3, f"the scene's 4-wide trees were not derived (stack need {st.wide_stack_need})"
    return dev


def scene_rays(b, n, seed):
    """rays that start on the scene's surfaces (the G-buffer points) in cosine-distributed directions, plus shadow-like rays"""
    orc = b.oracle()
    orc.prepass(b.inputs(1))
    pos = orc.readback(L.OUT_GBUFFER_POSITION).reshape(-1, 4)
    nrm = np.maximum(orc.readback(L.OUT_GBUFFER_NORMAL).astype(np.float32) / 127.0, -1.0).reshape(-1, 4)[:, :3]
    covered = np.nonzero(pos[:, 3] > 0)[0]
    rng = np.random.default_rng(seed)
    pick = rng.choice(covered, n, replace=True)
    N = nrm[pick] / np.maximum(np.linalg.norm(nrm[pick], axis=1, keepdims=True), 1e-6)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where((d * N).sum(1, keepdims=True) < 0, -d, d)
    rays = np.zeros(n, L.RAY)
    rays["origin"] = (pos[pick, :3] + N * 0.01).astype(np.float32)
    rays["direction"] = d.astype(np.float32)
    rays["max_distance"] = np.float32(3.402823466e38)
    shadow = rng.random(n) < 0.4
    far = rng.uniform(0.5, 30.0, n).astype(np.float32)
    rays["max_distance"] = np.where(shadow, far, rays["max_distance"])
    rays["early_distance"] = np.where(shadow, far - np.float32(0.1), np.float32(0.0))
    rays["exclude_instance"] = 0xFFFFFFFF
    return rays, shadow


@pytest.mark.parametrize("scene,config,n", [("cornell", "cornell_1080p", 60_000), ("city", "city_4k", 40_000), ("town", "scene_1080p", 20_000)])
def test_ray_dump_equals_the_fixed_order_walk(scene, config, n):
    b = Bench(scene, 96, 54, config=config)
    dev, orc = wide_device(b), b.oracle()
    for rays, shadow in (scene_rays(b, n, 3), (random_rays(n // 2, 5, any_hit_fraction=0.0), np.zeros(n // 2, bool))):
        hw, ho = dev.trace_rays(rays), orc.trace_rays(rays)
        hit_w, hit_o = hw["instance_index"] != 0xFFFFFFFF, ho["instance_index"] != 0xFFFFFFFF
        assert np.array_equal(hit_w[shadow], hit_o[shadow]), "a shadow ray's verdict differs"
        c = ~shadow
        same = np.ones(int(c.sum()), bool)
        for f in ("instance_index", "primitive_index"):
            same &= hw[f][c] == ho[f][c]
        for f in ("distance", "u", "v"):
            same &= hw[f][c].view(np.uint32) == ho[f][c].view(np.uint32)
        bad = int((~same).sum())
        assert bad <= 1e-4 * c.sum(), f"{bad} of {int(c.sum())} closest-hit rays differ from the fixed-order walk"


@pytest.mark.parametrize("scene,config,w,h,frames", [("cornell", "cornell_1080p", 96, 64, 8), ("city", "city_4k", 96, 54, 5),
                                                      ("simple", "cornell_1080p", 80, 48, 6)])
def test_frames_are_image_exact(scene, config, w, h, frames):
    b = Bench(scene, w, h, config=config)
    dev, orc = wide_device(b), b.oracle()
    worst = 0
    for f in range(1, frames + 1):
        inp = b.moving_inputs(f, step=(0.01, 0.004, -0.006)) if scene != "simple" else b.inputs(f)
        dev.render_frame(inp); orc.render_frame(inp)
        for k in GBUFFER:
            assert mismatch(dev.readback(k), orc.readback(k)) == 0, (f, k)
        for k in IMAGES:
            n_bad = mismatch(dev.readback(k), orc.readback(k))
            worst = max(worst, n_bad)
            # "< 1e-4 of the pixels" at benchmark resolutions; one pixel is the floor of what a small test frame can resolve
            assert n_bad <= max(1, int(1e-4 * w * h)), (f, k, n_bad)


def test_ray_counts_do_not_depend_on_the_walk():
    b = Bench("cornell", 80, 48, config="cornell_1080p")
    dev, orc = wide_device(b), b.oracle()
    dev.set_profiling(True, False)
    for f in range(1, 4):
        dev.render_frame(b.inputs(f)); orc.render_frame(b.inputs(f))
        sd, so = dev.stats(), orc.stats()
        assert (sd.primary_rays, sd.tlas_rays, sd.blas_rays) == (so.primary_rays, so.tlas_rays, so.blas_rays), f


def test_tie_rule_first_in_array_order():
    """two coincident planes in different instances (and materials): every ray that reaches them hits both at the same distance; the
    reference keeps the first in TLAS array order (strict '<', light.wgsl:416,470), and so must the ordered walk whichever child it
    enters first — the G-buffer's instance / material ids say which one won"""
    from bevy_hikari_b200 import scenes
    sd = scenes.minimal()
    xf = sd.inst_transform
    sd.inst_mesh, sd.inst_material, sd.inst_transform = [0, 0, 1, 0], [0, 1, 1, 1], [xf[0], xf[0], xf[1], xf[0]]
    scenes.SCENE_BUILDERS["coincident_planes"] = lambda: sd
    try:
        b = Bench("coincident_planes", 96, 64, config="cornell_1080p")
    finally:
        del scenes.SCENE_BUILDERS["coincident_planes"]
    dev, orc = wide_device(b), b.oracle()
    for f in range(1, 4):
        inp = b.inputs(f)
        dev.render_frame(inp); orc.render_frame(inp)
        for k in GBUFFER + IMAGES:
            assert mismatch(dev.readback(k), orc.readback(k)) == 0, (f, k)
    rays, shadow = scene_rays(b, 20_000, 9)
    hw, ho = dev.trace_rays(rays), orc.trace_rays(rays)
    c = ~shadow
    for name in ("instance_index", "primitive_index"):
        assert np.array_equal(hw[name][c], ho[name][c]), name


def test_switch_and_stats():
    b = Bench("cornell", 48, 32, config="cornell_256")
    dev = b.device()
    dev.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 3)
    st = dev.stats()
    assert st.wide_traversal == 3 and 0 < st.wide_stack_need <= 64
    dev.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 1)          # primary rays only, and only for scenes deep enough to gain: cornell is not
    assert dev.stats().wide_traversal == 0
    dev.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 0)
    assert dev.stats().wide_traversal == 0
    city = Bench("city", 48, 32, config="city_4k").device()
    city.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 1)
    assert city.stats().wide_traversal == 1
    with pytest.raises(Exception):
        city.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 2)


def test_unparseable_flat_array_keeps_the_reference_walk():
    """a TLAS that is not in bvh 0.7.1's flatten_custom layout (here: a navigator whose entry link skips its leaf) cannot be turned
    into a 4-wide tree: the launches keep the fixed-order walk and stay bit-exact with the oracle"""
    b = Bench("cornell", 48, 32, config="cornell_256")
    bufs = b.world.buffers()
    nodes = bufs["instance_nodes"]
    nav = np.nonzero(nodes["entry_index"] < 0x80000000)[0]
    last = nav[-1]                                   # the last navigator precedes the last leaf: point it at its own exit instead
    nodes["entry_index"][last] = nodes["exit_index"][last]
    desc = plugin.scene_desc_from_buffers(bufs)
    dev = plugin.HikariPlugin(48, 32)
    dev.upload_scene_desc(desc)
    dev.set_tuning(plugin.TUNE_WIDE_TRAVERSAL, 3)
    assert dev.stats().wide_traversal == 0
    from oracle import oracle
    orc = oracle.Oracle(48, 32, plugin.load_noise())
    orc.upload_scene_desc(desc)
    for f in range(1, 3):
        dev.render_frame(b.inputs(f)); orc.render_frame(b.inputs(f))
        for k in GBUFFER + IMAGES:
            assert mismatch(dev.readback(k), orc.readback(k)) == 0, (f, k)
