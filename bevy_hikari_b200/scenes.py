"""The reference's example scenes restated as data (examples/cornell.rs, examples/city.rs) + the benchmark configs of
BASELINE.json.  A scene = what the Bevy app spawns: meshes, materials, textures, instances, a camera, lights."""
import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import camera as cam
from . import layout as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "scenes")
F = np.float32


@dataclass
class SceneData:
    meshes: list                      # (positions, normals, uvs, indices)
    materials: np.ndarray             # L.MATERIAL records
    textures: list                    # dicts: rgba, address_mode_u/v, filter_linear, srgb
    inst_mesh: list = field(default_factory=list)
    inst_material: list = field(default_factory=list)
    inst_transform: list = field(default_factory=list)   # 16 floats, column-major
    eye: tuple = (0.0, 1.0, 4.0)
    target: tuple = (0.0, 1.0, 0.0)
    fov: float = math.pi / 4.0        # bevy PerspectiveProjection default
    near: float = 0.1
    sun_illuminance: float = None     # lux; None = no DirectionalLight
    sun_direction_to_light: tuple = (0.0, 1.0, 0.0)

    def view_inputs(self, width, height):
        proj = cam.perspective_infinite_reverse_rh(self.fov, width / height, self.near)
        world = cam.look_at(self.eye, self.target)
        view = cam.make_view(world, proj, width, height)
        return view, cam.make_previous_view(view), cam.make_lights(self.sun_illuminance, (1.0, 1.0, 1.0),
                                                                   self.sun_direction_to_light)

    def populate(self, world):
        """Spawn everything into a plugin.World (host mirror) and prepare the GPU buffers."""
        for t in self.textures:
            world.add_texture(t["rgba"], t["address_mode_u"], t["address_mode_v"], t["filter_linear"], t["srgb"])
        for m in self.meshes:
            world.add_mesh(*m)
        for m in self.materials:
            world.add_material(m)
        for me, ma, xf in zip(self.inst_mesh, self.inst_material, self.inst_transform):
            world.add_instance(int(me), int(ma), xf)
        world.prepare()
        return world


def _load_npz(name, texture_offset=0):
    z = np.load(os.path.join(SCENES, name + ".npz"))
    meshes = [(z[f"m{i}_pos"], z[f"m{i}_nrm"], z[f"m{i}_uv"], z[f"m{i}_idx"]) for i in range(int(z["mesh_count"]))]
    n = len(z["mat_base_color"])
    mats = np.zeros(n, L.MATERIAL)
    mats["base_color"] = z["mat_base_color"]
    mats["emissive"] = z["mat_emissive"]
    mats["perceptual_roughness"] = z["mat_perceptual_roughness"]
    mats["metallic"] = z["mat_metallic"]
    mats["reflectance"] = z["mat_reflectance"]
    tex = z["mat_textures"].astype(np.int64)
    tex = np.where(tex == 0xFFFFFFFF, 0xFFFFFFFF, tex + texture_offset).astype(np.uint32)
    for k, name_ in enumerate(("base_color_texture", "emissive_texture", "metallic_roughness_texture", "normal_map_texture",
                               "occlusion_texture")):
        mats[name_] = tex[:, k]
    textures = []
    for t in range(int(z["tex_count"])):
        info = z[f"t{t}_info"]
        textures.append({"rgba": z[f"t{t}_rgba"], "address_mode_u": int(info[0]), "address_mode_v": int(info[1]),
                         "filter_linear": int(info[2]), "srgb": int(info[3])})
    return meshes, mats, textures, z["inst_mesh"], z["inst_material"], z["inst_transform"]


def cornell():
    """examples/cornell.rs:37-61: cornell.glb, camera (0,1,4) -> (0,1,0), no directional light."""
    meshes, mats, textures, im, imat, ixf = _load_npz("cornell")
    return SceneData(meshes, mats, textures, list(im), list(imat), list(ixf), eye=(0.0, 1.0, 4.0), target=(0.0, 1.0, 0.0))


def _translation(x, y, z, s=(1.0, 1.0, 1.0)):
    m = np.zeros((4, 4), F)
    m[0, 0], m[1, 1], m[2, 2], m[3, 3] = s[0], s[1], s[2], 1.0
    m[3, :3] = (x, y, z)
    return m.reshape(16)


def _compose(parent16, child16):
    a, b = parent16.reshape(4, 4).astype(np.float64), child16.reshape(4, 4).astype(np.float64)
    return (b @ a).astype(F).reshape(16)


def _plane_mesh(size=1.0):
    """bevy shape::Plane { size }: 4 vertices, normal +Y."""
    e = size / 2.0
    pos = np.array([[e, 0, -e], [e, 0, e], [-e, 0, e], [-e, 0, -e]], F)
    nrm = np.tile(np.array([[0, 1, 0]], F), (4, 1))
    uv = np.array([[1, 0], [1, 1], [0, 1], [0, 0]], F)
    idx = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    return pos, nrm, uv, idx


def _uv_sphere_mesh(radius=0.5, sectors=36, stacks=18):
    """bevy shape::UVSphere."""
    pos, nrm, uv = [], [], []
    for i in range(stacks + 1):
        stack_angle = math.pi / 2 - i * math.pi / stacks
        xy, z = radius * math.cos(stack_angle), radius * math.sin(stack_angle)
        for j in range(sectors + 1):
            a = j * 2 * math.pi / sectors
            x, y = xy * math.cos(a), xy * math.sin(a)
            pos.append((x, y, z)); nrm.append((x / radius, y / radius, z / radius)); uv.append((j / sectors, i / stacks))
    idx = []
    for i in range(stacks):
        k1, k2 = i * (sectors + 1), (i + 1) * (sectors + 1)
        for j in range(sectors):
            if i != 0:
                idx += [k1, k2, k1 + 1]
            if i != stacks - 1:
                idx += [k1 + 1, k2, k2 + 1]
            k1 += 1; k2 += 1
    return np.array(pos, F), np.array(nrm, F), np.array(uv, F), np.array(idx, np.uint32)


def city():
    """examples/city.rs:56-202 after all loads: ground plane x100, emissive sphere at (0,1,0), 12 Low-Poly houses,
    sun 10 klx, camera (0,2.5,20) -> origin."""
    meshes, mat_list, textures = [], [], []
    inst_mesh, inst_mat, inst_xf = [], [], []

    def std_material(base=(1, 1, 1, 1), emissive=(0, 0, 0, 1), rough=0.089, metallic=0.01, refl=0.5):
        m = np.zeros((), L.MATERIAL)
        m["base_color"], m["emissive"] = base, emissive
        m["perceptual_roughness"], m["metallic"], m["reflectance"] = rough, metallic, refl
        for k in ("base_color_texture", "emissive_texture", "metallic_roughness_texture", "normal_map_texture", "occlusion_texture"):
            m[k] = 0xFFFFFFFF
        return m

    # ground: shape::Plane::default() scaled (100,1,100), base colour rgb(0.8,0.7,0.6), perceptual_roughness 0.9 (city.rs:62-77)
    meshes.append(_plane_mesh(1.0)); mat_list.append(std_material(base=(0.8, 0.7, 0.6, 1.0), rough=0.9))
    inst_mesh.append(0); inst_mat.append(0); inst_xf.append(_translation(0, 0, 0, (100.0, 1.0, 100.0)))
    # emissive UV sphere r=0.5 at (0,1,0) rotated -90 deg about X; earth_daymap.jpg is both base-colour and emissive
    # texture, emissive rgba(1,1,1,0.5) (city.rs:80-100).  The JPEG is shipped down-sampled to 512x256.
    earth = np.load(os.path.join(SCENES, "earth.npz"))["rgba"]
    textures.append({"rgba": earth, "address_mode_u": 1, "address_mode_v": 1, "filter_linear": 1, "srgb": 1})  # bevy default sampler: clamp
    sphere_mat = std_material(base=(1, 1, 1, 1), emissive=(1.0, 1.0, 1.0, 0.5))
    sphere_mat["base_color_texture"] = 0
    sphere_mat["emissive_texture"] = 0
    meshes.append(_uv_sphere_mesh(0.5, 36, 18)); mat_list.append(sphere_mat)
    rot = np.zeros((4, 4), F)   # Quat::from_rotation_x(-pi/2): columns (1,0,0), (0,0,-1), (0,1,0)
    rot[0, 0] = 1.0; rot[1, 2] = -1.0; rot[2, 1] = 1.0; rot[3, :] = (0.0, 1.0, 0.0, 1.0)
    inst_mesh.append(1); inst_mat.append(1); inst_xf.append(rot.reshape(16))

    def add_house(name, positions):
        hm, hmat, htex, him, himat, hixf = _load_npz(name, texture_offset=len(textures))
        mesh0, mat0 = len(meshes), len(mat_list)
        meshes.extend(hm); mat_list.extend(list(hmat)); textures.extend(htex)
        for p in positions:
            parent = _translation(*p)
            for me, ma, xf in zip(him, himat, hixf):
                inst_mesh.append(mesh0 + int(me)); inst_mat.append(mat0 + int(ma)); inst_xf.append(_compose(parent, xf))

    add_house("house2", [(-12.0, 0.0, 0.0), (-4.0, 0.0, 0.0), (4.0, 0.0, 0.0), (12.0, 0.0, 0.0)])
    add_house("house3", [(-12.0, 0.0, 8.0), (-4.0, 0.0, -8.0), (4.0, 0.0, 8.0), (12.0, 0.0, -8.0)])
    add_house("house", [(-12.0, 0.0, -8.0), (-4.0, 0.0, 8.0), (4.0, 0.0, -8.0), (12.0, 0.0, 8.0)])
    mats = np.array(mat_list, L.MATERIAL)
    return SceneData(meshes, mats, textures, inst_mesh, inst_mat, inst_xf, eye=(0.0, 2.5, 20.0), target=(0.0, 0.0, 0.0),
                     sun_illuminance=10000.0,
                     sun_direction_to_light=tuple(cam.euler_xyz_back(-math.pi / 4.0, math.pi / 4.0, 0.0)))


def town():
    """examples/scene.rs:54-143 (BASELINE configs[2]): ground plane scaled (10000,1,10000) at y = -3, assets/models/scene.gltf
    (84 meshes / instances, 120 440 triangles, 66 materials, 52 base-colour textures, shipped box-filtered to <= 256 px),
    emissive UV sphere r = 0.5 at (2,2,0) with the earth texture, sun 100 klx rotated XYZ(-pi/4, pi/4, 0), camera
    (-20,10,20) -> origin.  Instance order = entity order: the two PbrBundles spawned in `setup` precede the entities the
    glTF scene spawner creates once the asset has loaded."""
    tm, tmat, ttex, tim, timat, tixf = _load_npz("town", texture_offset=1)
    meshes, mat_list, textures = [], [], []
    inst_mesh, inst_mat, inst_xf = [], [], []
    # ground (scene.rs:60-76)
    meshes.append(_plane_mesh(1.0)); mat_list.append(_std_material(base=(0.8, 0.7, 0.6, 1.0), rough=0.9))
    inst_mesh.append(0); inst_mat.append(0); inst_xf.append(_translation(0.0, -3.0, 0.0, (10000.0, 1.0, 10000.0)))
    # emissive sphere (scene.rs:85-107); bevy's default image sampler clamps
    earth = np.load(os.path.join(SCENES, "earth.npz"))["rgba"]
    textures.append({"rgba": earth, "address_mode_u": 1, "address_mode_v": 1, "filter_linear": 1, "srgb": 1})
    sphere_mat = _std_material(base=(1, 1, 1, 1), emissive=(1.0, 1.0, 1.0, 0.5))
    sphere_mat["base_color_texture"] = 0
    sphere_mat["emissive_texture"] = 0
    meshes.append(_uv_sphere_mesh(0.5, 36, 18)); mat_list.append(sphere_mat)
    rot = np.zeros((4, 4), F)   # Quat::from_rotation_x(-pi/2)
    rot[0, 0] = 1.0; rot[1, 2] = -1.0; rot[2, 1] = 1.0; rot[3, :] = (2.0, 2.0, 0.0, 1.0)
    inst_mesh.append(1); inst_mat.append(1); inst_xf.append(rot.reshape(16))
    # the glTF scene, Transform::default() (scene.rs:79-83)
    mesh0, mat0 = len(meshes), len(mat_list)
    meshes.extend(tm); mat_list.extend(list(tmat)); textures.extend(ttex)
    for me, ma, xf in zip(tim, timat, tixf):
        inst_mesh.append(mesh0 + int(me)); inst_mat.append(mat0 + int(ma)); inst_xf.append(np.asarray(xf, F))
    mats = np.array(mat_list, L.MATERIAL)
    return SceneData(meshes, mats, textures, inst_mesh, inst_mat, inst_xf, eye=(-20.0, 10.0, 20.0), target=(0.0, 0.0, 0.0),
                     sun_illuminance=100000.0,
                     sun_direction_to_light=tuple(cam.euler_xyz_back(-math.pi / 4.0, math.pi / 4.0, 0.0)))


def _box_mesh(sx=1.0, sy=1.0, sz=1.0):
    """bevy shape::Box::new(sx, sy, sz) (shape::Cube { size } = Box::new(size, size, size)): 24 vertices, 6 faces in
    bevy's order (+z, -z, +x, -x, +y, -y), two triangles per face."""
    x0, x1, y0, y1, z0, z1 = -sx / 2, sx / 2, -sy / 2, sy / 2, -sz / 2, sz / 2
    v = [
        ((x0, y0, z1), (0, 0, 1), (0, 0)), ((x1, y0, z1), (0, 0, 1), (1, 0)), ((x1, y1, z1), (0, 0, 1), (1, 1)), ((x0, y1, z1), (0, 0, 1), (0, 1)),
        ((x0, y1, z0), (0, 0, -1), (1, 0)), ((x1, y1, z0), (0, 0, -1), (0, 0)), ((x1, y0, z0), (0, 0, -1), (0, 1)), ((x0, y0, z0), (0, 0, -1), (1, 1)),
        ((x1, y0, z0), (1, 0, 0), (0, 0)), ((x1, y1, z0), (1, 0, 0), (1, 0)), ((x1, y1, z1), (1, 0, 0), (1, 1)), ((x1, y0, z1), (1, 0, 0), (0, 1)),
        ((x0, y0, z1), (-1, 0, 0), (1, 0)), ((x0, y1, z1), (-1, 0, 0), (0, 0)), ((x0, y1, z0), (-1, 0, 0), (0, 1)), ((x0, y0, z0), (-1, 0, 0), (1, 1)),
        ((x1, y1, z0), (0, 1, 0), (1, 0)), ((x0, y1, z0), (0, 1, 0), (0, 0)), ((x0, y1, z1), (0, 1, 0), (0, 1)), ((x1, y1, z1), (0, 1, 0), (1, 1)),
        ((x1, y0, z1), (0, -1, 0), (0, 0)), ((x0, y0, z1), (0, -1, 0), (1, 0)), ((x0, y0, z0), (0, -1, 0), (1, 1)), ((x1, y0, z0), (0, -1, 0), (0, 1)),
    ]
    pos = np.array([a for a, _, _ in v], F); nrm = np.array([b for _, b, _ in v], F); uv = np.array([c for _, _, c in v], F)
    idx = np.array([k + o for k in range(0, 24, 4) for o in (0, 1, 2, 2, 3, 0)], np.uint32)
    return pos, nrm, uv, idx


def _std_material(base=(1, 1, 1, 1), emissive=(0, 0, 0, 1), rough=0.089, metallic=0.01, refl=0.5):
    """bevy 0.9 StandardMaterial::default() with overrides; Color::rgb values are passed through as_rgba_f32 (material.rs:168)."""
    m = np.zeros((), L.MATERIAL)
    m["base_color"], m["emissive"] = base, emissive
    m["perceptual_roughness"], m["metallic"], m["reflectance"] = rough, metallic, refl
    for k in ("base_color_texture", "emissive_texture", "metallic_roughness_texture", "normal_map_texture", "occlusion_texture"):
        m[k] = 0xFFFFFFFF
    return m


_ROT_X_NEG_90 = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], F)   # Quat::from_rotation_x(-pi/2), rows = columns


def minimal():
    """examples/minimal.rs:20-66: Plane { size: 5 } rgb(0.3,0.5,0.3), Cube { size: 1 } rgb(0.8,0.7,0.6) at (0,0.5,0), sun
    10 klx with rotation XYZ(-pi/4, pi/4, 0), camera (-2,2.5,5) -> origin, HikariSettings::default().  No emissive at all."""
    meshes = [_plane_mesh(5.0), _box_mesh(1.0, 1.0, 1.0)]
    mats = np.array([_std_material((0.3, 0.5, 0.3, 1.0)), _std_material((0.8, 0.7, 0.6, 1.0))], L.MATERIAL)
    return SceneData(meshes, mats, [], [0, 1], [0, 1], [_translation(0, 0, 0), _translation(0.0, 0.5, 0.0)],
                     eye=(-2.0, 2.5, 5.0), target=(0.0, 0.0, 0.0), sun_illuminance=10000.0,
                     sun_direction_to_light=tuple(cam.euler_xyz_back(-math.pi / 4.0, math.pi / 4.0, 0.0)))


def simple():
    """examples/simple.rs:54-262 without the two extinguisher.glb instances: a room of scaled unit cubes (ground, left,
    right, back, top), a 400 x 400 ground plane, TWO emissive earth spheres (emissive alpha 0.5 and 0.1) that the example
    rotates about their local z (sphere_rotate_system, :385-389), sun 10 klx, camera (-10,2.5,20) -> origin."""
    earth = np.load(os.path.join(SCENES, "earth.npz"))["rgba"]
    textures = [{"rgba": earth, "address_mode_u": 1, "address_mode_v": 1, "filter_linear": 1, "srgb": 1}]
    meshes = [_box_mesh(1.0, 1.0, 1.0), _plane_mesh(1.0), _uv_sphere_mesh(0.5, 36, 18), _uv_sphere_mesh(0.5, 36, 18)]
    mat = [_std_material((0.3, 0.5, 0.3, 1.0), rough=0.9), _std_material((1, 1, 1, 1), rough=0.9),
           _std_material((1.0, 0.08, 0.58, 1.0), rough=0.9),       # Color::PINK
           _std_material((1, 1, 1, 1), rough=0.9),
           _std_material((0.49, 1.0, 0.83, 1.0), rough=0.9),       # Color::AQUAMARINE
           _std_material((1, 1, 1, 1), rough=0.9)]
    for alpha in (0.5, 0.1):
        m = _std_material((1, 1, 1, 1), emissive=(1.0, 1.0, 1.0, alpha))
        m["base_color_texture"] = 0
        m["emissive_texture"] = 0
        mat.append(m)
    inst_mesh = [0, 1, 0, 0, 0, 0, 2, 3]
    inst_mat = [0, 1, 2, 3, 4, 5, 6, 7]
    xf = [_translation(0.0, -0.5, 0.0, (8.0, 1.0, 8.0)), _translation(0.0, -1.0, 0.0, (400.0, 1.0, 400.0)),
          _translation(-3.5, 3.0, 0.0, (1.0, 6.0, 8.0)), _translation(3.5, 3.0, 0.0, (1.0, 6.0, 8.0)),
          _translation(0.0, 3.0, -3.5, (6.0, 6.0, 1.0)), _translation(0.0, 6.5, 0.0, (8.0, 1.0, 8.0))]
    for x in (2.0, -2.0):
        m = _ROT_X_NEG_90.copy()
        m[3, :3] = (x, 1.0, 0.0)
        xf.append(m.reshape(16))
    return SceneData(meshes, np.array(mat, L.MATERIAL), textures, inst_mesh, inst_mat, xf,
                     eye=(-10.0, 2.5, 20.0), target=(0.0, 0.0, 0.0), sun_illuminance=10000.0,
                     sun_direction_to_light=tuple(cam.euler_xyz_back(-math.pi / 4.0, math.pi / 4.0, 0.0)))


def samplers():
    """Test scene for the texture path (mod.rs:760-799 binds every texture with its own sampler): one ground plane whose
    uvs run from -1 to 2 so that addressing matters, textured four times over with the earth image under different samplers
    (repeat / clamp / mirror, linear / nearest, sRGB / linear data), lit by a small emissive sphere with a mirrored emissive
    texture and the sun."""
    earth = np.load(os.path.join(SCENES, "earth.npz"))["rgba"][::4, ::4].copy()    # 128 x 64 keeps texels visible
    combos = [(0, 0, 1, 1), (1, 2, 0, 1), (2, 1, 1, 0), (2, 0, 0, 0)]             # (mode_u, mode_v, linear, srgb); 0 repeat, 1 clamp, 2 mirror
    textures = [{"rgba": earth, "address_mode_u": u, "address_mode_v": v, "filter_linear": lin, "srgb": srgb} for u, v, lin, srgb in combos]
    textures.append({"rgba": earth, "address_mode_u": 2, "address_mode_v": 2, "filter_linear": 1, "srgb": 1})
    pos, nrm, uv, idx = _plane_mesh(1.0)
    uv = (uv * 3.0 - 1.0).astype(F)
    meshes = [(pos, nrm, uv, idx), _uv_sphere_mesh(0.5, 24, 12)]
    mats = []
    for t in range(4):
        m = _std_material((1, 1, 1, 1), rough=0.7)
        m["base_color_texture"] = t
        mats.append(m)
    light = _std_material((1, 1, 1, 1), emissive=(1.0, 0.9, 0.8, 0.4))
    light["emissive_texture"] = 4
    light["base_color_texture"] = 4
    mats.append(light)
    inst_mesh = [0, 0, 0, 0, 1]
    inst_mat = [0, 1, 2, 3, 4]
    xf = [_translation(-1.5, 0.0, -1.5, (3.0, 1.0, 3.0)), _translation(1.5, 0.0, -1.5, (3.0, 1.0, 3.0)),
          _translation(-1.5, 0.0, 1.5, (3.0, 1.0, 3.0)), _translation(1.5, 0.0, 1.5, (3.0, 1.0, 3.0))]
    m = _ROT_X_NEG_90.copy()
    m[3, :3] = (0.0, 0.9, 0.0)
    xf.append(m.reshape(16))
    return SceneData(meshes, np.array(mats, L.MATERIAL), textures, inst_mesh, inst_mat, xf, eye=(0.0, 3.5, 5.5), target=(0.0, 0.0, 0.0),
                     sun_illuminance=6000.0, sun_direction_to_light=tuple(cam.euler_xyz_back(-math.pi / 3.0, math.pi / 6.0, 0.0)))


def soup(seed=0):
    """Random test scene (tools/fuzz_parity.py): a handful of random triangle meshes — some triangles degenerate, some
    meshes a single triangle — instanced with random rotations, non-uniform and mirrored scales, random materials, several of
    them emissive, a random sun.  Nothing here comes from the reference; it exists to take both implementations off the
    beaten path together."""
    rng = np.random.default_rng(seed)
    meshes, mats, inst_mesh, inst_mat, xf = [], [], [], [], []
    for _ in range(int(rng.integers(2, 6))):
        n_tris = int(rng.integers(1, 13))
        centre = rng.uniform(-0.5, 0.5, 3)
        pos = (centre + rng.uniform(-0.6, 0.6, (3 * n_tris, 3))).astype(F)
        if n_tris > 2 and rng.integers(0, 2):
            pos[3:6] = pos[3]                                        # a zero-area triangle
        if n_tris > 4 and rng.integers(0, 2):
            pos[7] = pos[6] + (pos[8] - pos[6]) * 0.5                # three collinear points
        e1, e2 = pos[1::3] - pos[0::3], pos[2::3] - pos[0::3]
        nrm = np.repeat(np.cross(e1, e2), 3, axis=0)
        ln = np.linalg.norm(nrm, axis=1, keepdims=True)
        nrm = np.where(ln > 1e-12, nrm / np.maximum(ln, 1e-12), np.array([[0.0, 1.0, 0.0]])).astype(F)
        uv = rng.uniform(0, 1, (3 * n_tris, 2)).astype(F)
        meshes.append((pos, nrm, uv, np.arange(3 * n_tris, dtype=np.uint32)))
    for _ in range(int(rng.integers(2, 6))):
        emissive = (0, 0, 0, 1)
        if rng.integers(0, 3) == 0:
            emissive = tuple(rng.uniform(0.2, 1.0, 3)) + (float(rng.uniform(0.01, 0.3)),)
        mats.append(_std_material(tuple(rng.uniform(0.05, 1.0, 3)) + (1.0,), emissive=emissive, rough=float(rng.uniform(0.0, 1.0)),
                                  metallic=float(rng.choice([0.0, 0.01, 0.5, 1.0])), refl=float(rng.uniform(0.0, 1.0))))
    for _ in range(int(rng.integers(2, 9))):
        inst_mesh.append(int(rng.integers(0, len(meshes)))); inst_mat.append(int(rng.integers(0, len(mats))))
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        scale = rng.uniform(0.3, 2.0, 3) * rng.choice([1.0, 1.0, -1.0], 3)
        m = np.zeros((4, 4), F)
        m[:3, :3] = (R * scale[None, :]).T          # rows of the array = columns of the matrix
        m[3, :3] = rng.uniform(-1.5, 1.5, 3)
        m[3, 3] = 1.0
        xf.append(m.reshape(16))
    sun = float(rng.choice([0.0, 3000.0, 20000.0]))
    d = rng.normal(size=3); d[1] = abs(d[1]) + 0.2
    return SceneData(meshes, np.array(mats, L.MATERIAL), [], inst_mesh, inst_mat, xf,
                     eye=tuple(rng.uniform(-1.0, 1.0, 3) + np.array([0.0, 0.5, 4.0])), target=tuple(rng.uniform(-0.5, 0.5, 3)),
                     sun_illuminance=sun if sun > 0 else None, sun_direction_to_light=tuple(d / np.linalg.norm(d)))


def terrain(n=224, seed=7):
    """Stress scene, not a reference example: one n x n-quad displaced grid (2 n^2 triangles — 100 352 at n = 224, the size
    class of the reference's scene.gltf) under a small emissive sphere and the sun; exercises a deep
    BLAS (3 * 100 352 - 2 records) in the host builder and in traversal."""
    rng = np.random.default_rng(seed)
    g = np.linspace(-6.0, 6.0, n + 1, dtype=np.float64)
    x, z = np.meshgrid(g, g, indexing="xy")
    y = 0.35 * np.sin(1.3 * x) * np.cos(0.9 * z) + 0.15 * np.sin(3.1 * x + 1.0) + 0.03 * rng.standard_normal(x.shape)
    pos = np.stack([x, y, z], axis=2).reshape(-1, 3).astype(F)
    gy, gx = np.gradient(y, g, g)                      # d/dz (rows), d/dx (cols)
    nrm = np.stack([-gx, np.ones_like(y), -gy], axis=2)
    nrm = (nrm / np.linalg.norm(nrm, axis=2, keepdims=True)).reshape(-1, 3).astype(F)
    uv = np.stack([(x + 6.0) / 12.0, (z + 6.0) / 12.0], axis=2).reshape(-1, 2).astype(F)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    a = (j * (n + 1) + i).reshape(-1); b = a + 1; c = a + (n + 1); d = c + 1
    idx = np.stack([a, c, b, b, c, d], axis=1).reshape(-1).astype(np.uint32)

    def std_material(base, emissive=(0, 0, 0, 1), rough=0.6):
        m = np.zeros((), L.MATERIAL)
        m["base_color"], m["emissive"] = base, emissive
        m["perceptual_roughness"], m["metallic"], m["reflectance"] = rough, 0.01, 0.5
        for k in ("base_color_texture", "emissive_texture", "metallic_roughness_texture", "normal_map_texture", "occlusion_texture"):
            m[k] = 0xFFFFFFFF
        return m

    meshes = [(pos, nrm, uv, idx), _uv_sphere_mesh(0.5, 24, 12)]
    mats = np.array([std_material((0.55, 0.7, 0.45, 1.0)), std_material((1, 1, 1, 1), emissive=(1.0, 0.8, 0.5, 0.6))], L.MATERIAL)
    return SceneData(meshes, mats, [], [0, 1], [0, 1], [_translation(0, 0, 0), _translation(0.5, 1.6, 0.0)],
                     eye=(0.0, 3.0, 9.0), target=(0.0, 0.0, 0.0), sun_illuminance=8000.0,
                     sun_direction_to_light=tuple(cam.euler_xyz_back(-math.pi / 3.0, math.pi / 5.0, 0.0)))


SCENE_BUILDERS = {"cornell": cornell, "city": city, "town": town, "terrain": terrain, "minimal": minimal, "simple": simple, "samplers": samplers}

# BASELINE.json configs.  All run with Upscale::SmaaTu4x{ratio 1.0}, Taa::None so that the render
# resolution equals the stated resolution.
CONFIGS = {
    # configs[0]: cornell 256x256, 1 indirect bounce, denoise off
    "cornell_256": dict(scene="cornell", width=256, height=256,
                        settings=dict(indirect_bounces=1, denoise=0, temporal_reuse=1, emissive_spatial_reuse=0,
                                      indirect_spatial_reuse=1)),
    # configs[1]: cornell 1920x1080, 2 bounces, ReSTIR temporal + spatial (emissive and indirect) + denoise  <- bench default
    "cornell_1080p": dict(scene="cornell", width=1920, height=1080,
                          settings=dict(indirect_bounces=2, denoise=1, temporal_reuse=1, emissive_spatial_reuse=1,
                                        indirect_spatial_reuse=1)),
    # configs[2]: examples/scene.rs 1920x1080, 3 bounces, emissive + indirect spatial reuse
    "scene_1080p": dict(scene="town", width=1920, height=1080,
                        settings=dict(indirect_bounces=3, denoise=1, temporal_reuse=1, emissive_spatial_reuse=1,
                                      indirect_spatial_reuse=1)),
    # configs[3]: city 3840x2160, 2 bounces, indirect spatial + denoise, row bands over 4 GPUs
    "city_4k": dict(scene="city", width=3840, height=2160,
                    settings=dict(indirect_bounces=2, denoise=1, temporal_reuse=1, emissive_spatial_reuse=0,
                                  indirect_spatial_reuse=1)),
    # configs[4]: city 7680x4320, 4 bounces, full ReSTIR + denoise, 8 GPUs
    "city_8k": dict(scene="city", width=7680, height=4320,
                    settings=dict(indirect_bounces=4, denoise=1, temporal_reuse=1, emissive_spatial_reuse=1,
                                  indirect_spatial_reuse=1)),
}


def config_settings(name, **extra):
    from . import plugin
    kw = dict(taa=plugin.TAA_NONE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=1.0)
    kw.update(CONFIGS[name]["settings"])
    kw.update(extra)
    return plugin.HikariSettings(**kw)
