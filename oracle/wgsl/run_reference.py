#!/usr/bin/env python
"""TEST INFRASTRUCTURE — drives the C++ translations of the reference's WGSL (oracle/wgsl/wgsl2cpp.py) pass by pass, wired as the
reference's render-graph nodes wire their bind groups:

  LightNode::run          src/light.rs:590-702     pass order, which pipeline variant runs when, workgroup counts
  queue_light_bind_groups src/light.rs:463-555     reservoir buffers (temporal, spatial) = (0,4) sun, (2,4) emissive, (6,8) indirect,
                                                   `current` = frame counter % 2; render / variance texture per pass
  PostProcessNode::run    src/post_process.rs:1140-1234 + bind groups :840-1000: demodulation + four denoise levels per signal
                                                   (firefly filtering for the emissive and indirect signals only), tone mapping;
                                                   :1236-1277 + bind groups :983-1036: smaa_tu4x, smaa_tu4x_extrapolate, taa_jasmine

Only in this container (it needs /root/reference and g++); the libraries it builds live in oracle/_ref/wgsl/ (git-ignored).  The
G-buffer is an INPUT here: the reference rasterises it (src/prepass.rs + prepass.wgsl, a render pipeline, not translated), so the
caller passes the five G-buffer planes — tools/make_wgsl_golden.py and the tests take them from the oracle's prepass."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "oracle", "_ref", "wgsl")
sys.path.insert(0, HERE)
import wgsl2cpp  # noqa: E402
import glsl2cpp  # noqa: E402

RGBA32F, RGBA16F, R32F, RG32F, RGBA8SNORM, RGBA8UNORM, RGBA8SRGB = range(7)
CXXFLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-w", "-I", HERE, "-I", os.path.join(ROOT, "include")]


def available():
    return os.path.isdir(wgsl2cpp.REF_SHADERS) and (os.path.exists("/usr/bin/g++"))


def build(shader, defs):
    """translate + compile src/shaders/<shader>.wgsl with `defs`; returns the loaded library"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(wgsl2cpp.REF_SHADERS, shader + ".wgsl")
    cpp_text = wgsl2cpp.translate(src, defs)
    rt = open(os.path.join(HERE, "wgsl_rt.h")).read()
    tag = hashlib.sha1((cpp_text + rt + " ".join(CXXFLAGS)).encode()).hexdigest()[:12]
    name = shader + ("_" + "_".join(sorted(defs)).lower() if defs else "")
    so = os.path.join(OUT, f"{name}_{tag}.so")
    if not os.path.exists(so):
        cpp = os.path.join(OUT, name + ".cpp")
        open(cpp, "w").write(cpp_text)
        r = subprocess.run(["g++"] + CXXFLAGS + [cpp, "-o", so], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on the translation of {shader}.wgsl {defs}:\n{r.stderr[-3000:]}")
    return C.CDLL(so, mode=C.RTLD_LOCAL)


def build_fsr(define):
    """translate + compile one FSR 1.0 pass (src/shaders/fsr/source.zip, oracle/wgsl/glsl2cpp.py); returns the loaded library"""
    os.makedirs(OUT, exist_ok=True)
    cpp_text = glsl2cpp.translate(define)
    rt = open(os.path.join(HERE, "wgsl_rt.h")).read() + open(os.path.join(HERE, "glsl_rt.h")).read()
    tag = hashlib.sha1((cpp_text + rt + " ".join(CXXFLAGS)).encode()).hexdigest()[:12]
    name = "fsr_" + define.lower()
    so = os.path.join(OUT, f"{name}_{tag}.so")
    if not os.path.exists(so):
        cpp = os.path.join(OUT, name + ".cpp")
        open(cpp, "w").write(cpp_text)
        r = subprocess.run(["g++"] + CXXFLAGS + [cpp, "-o", so], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on the translation of FSR_Pass.glsl ({define}):\n{r.stderr[-3000:]}")
    return C.CDLL(so, mode=C.RTLD_LOCAL)


class FsrPass:
    """one FSR 1.0 pass: InputTexture / InputSampler / OutputTexture / const_buffer of FSR_Pass.glsl"""

    def __init__(self, define):
        self.lib = build_fsr(define)
        self.lib.bind_input.argtypes = self.lib.bind_output.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.bind_sampler.argtypes = [C.c_int] * 3
        self.lib.set_constants.argtypes = [C.c_float] * 7 + [C.c_uint]
        self.lib.run.argtypes = [C.c_uint, C.c_uint]

    def run(self, src, dst, constants, groups):
        assert src.flags["C_CONTIGUOUS"] and dst.flags["C_CONTIGUOUS"]
        self.lib.bind_input(src.ctypes.data, src.shape[1], src.shape[0], RGBA16F)
        self.lib.bind_output(dst.ctypes.data, dst.shape[1], dst.shape[0], RGBA16F)
        self.lib.bind_sampler(1, 1, 1)                         # binding 1 of the sampler group: linear_sampler, clamp to edge
        self.lib.set_constants(*constants)
        self.lib.run(*groups)


class Module:
    """one translated shader module: typed access to bind_* / run_*"""

    def __init__(self, shader, defs):
        self.lib = build(shader, defs)
        self.keep = []          # arrays bound into the library must outlive the calls

    def buffer(self, name, array):
        a = np.ascontiguousarray(array)
        self.keep.append(a)
        fn = getattr(self.lib, "bind_" + name)
        fn.argtypes = [C.c_void_p, C.c_size_t]
        fn(a.ctypes.data, a.nbytes)
        return a

    def buffer_inplace(self, name, array):
        assert array.flags["C_CONTIGUOUS"]
        fn = getattr(self.lib, "bind_" + name)
        fn.argtypes = [C.c_void_p, C.c_size_t]
        fn(array.ctypes.data, array.nbytes)

    def texture(self, name, array, fmt, index=None):
        assert array.flags["C_CONTIGUOUS"]
        h, w = array.shape[0], array.shape[1]
        fn = getattr(self.lib, "bind_" + name)
        if index is None:
            fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
            fn(array.ctypes.data, w, h, fmt)
        else:
            fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
            fn(index, array.ctypes.data, w, h, fmt)

    def sampler(self, name, mode_u, mode_v, linear, index=None):
        fn = getattr(self.lib, "bind_" + name)
        if index is None:
            fn.argtypes = [C.c_int] * 3
            fn(mode_u, mode_v, linear)
        else:
            fn.argtypes = [C.c_int] * 4
            fn(index, mode_u, mode_v, linear)

    def run(self, entry, width, height, order=None):
        """`order`: (n, 2) uint32 global ids, every invocation of the dispatch once — or None for raster order"""
        so = self.lib.wgsl_set_order
        so.argtypes = [C.c_void_p, C.c_size_t]
        if order is not None:
            order = np.ascontiguousarray(order, np.uint32)
            so(order.ctypes.data, len(order))
        else:
            so(None, 0)
        fn = getattr(self.lib, "run_" + entry)
        fn.argtypes = [C.c_uint] * 3
        fn((width + 7) // 8, (height + 7) // 8, 1)       # WORKGROUP_SIZE = 8, src/lib.rs:53
        so(None, 0)


def nodes_buffer(nodes):
    """the reference's `Nodes` storage buffer: { count: u32, data: array<Node> } (mesh_material_types.wgsl:80-83): 16-byte header"""
    raw = np.zeros(16 + nodes.nbytes, np.uint8)
    raw[:4] = np.frombuffer(np.uint32(len(nodes)).tobytes(), np.uint8)
    raw[16:] = np.frombuffer(nodes.tobytes(), np.uint8)
    return raw


def lights_buffer(lights):
    """bevy_pbr 0.9 `Lights` (prelude/bevy_pbr_mesh_view_types.wgsl) from the slice the path reads (hk_lights)"""
    raw = np.zeros(1184, np.uint8)
    f = raw.view(np.float32)
    f[16:20] = np.array(lights.directional_color, np.float32)          # directional_lights[0].color           @ 64
    f[20:23] = np.array(lights.direction_to_light, np.float32)         # directional_lights[0].direction_to_light @ 80
    f[280:284] = np.array(lights.ambient_color, np.float32)            # ambient_color                          @ 1120
    raw.view(np.uint32)[292] = 1 if any(lights.directional_color) else 0   # n_directional_lights                @ 1168
    return raw


class WgslReference:
    """The reference's compute passes of one camera: state = what the render world keeps between frames (10 reservoir buffers,
    render / variance / albedo textures, denoise internals)."""

    LIGHT_VARIANTS = {"base": [], "sun": ["RENDER_EMISSIVE"], "emissive": ["EMISSIVE_LIT"], "multi": ["MULTIPLE_BOUNCES"]}

    def __init__(self, scene_buffers, textures, noise, width, height, upscale_ratio=1.0):
        """scene_buffers: dict of the nine storage buffers as numpy records (plugin.World.buffers()); textures: list of
        (rgba8 HxWx4, address_mode_u, address_mode_v, filter_linear, srgb); noise: 16 x 64 x 64 x 4 uint8"""
        self.w, self.h = width, height
        # scaled_size = (ratio.recip() * size).ceil(), light.rs:321-322 (f32 arithmetic)
        scale = np.float32(1.0) / np.float32(upscale_ratio)
        self.rw, self.rh = int(np.ceil(scale * np.float32(width))), int(np.ceil(scale * np.float32(height)))
        self.scene = {k: np.ascontiguousarray(v) for k, v in scene_buffers.items()}
        self.textures = textures
        self.noise = np.ascontiguousarray(noise, np.uint8).reshape(16, 64, 64, 4)
        no_tex = ["NO_TEXTURE"] if not textures else []                      # light.rs:141-143
        self.light = {k: Module("light", no_tex + d) for k, d in self.LIGHT_VARIANTS.items()}
        self.demodulation = Module("denoise", ["DENOISE_LEVEL_0"])          # post_process.rs:423: every denoise-shader pipeline gets a level
        self.denoise = {(lvl, ff): Module("denoise", [f"DENOISE_LEVEL_{lvl}"] + (["FIREFLY_FILTERING"] if ff else []))
                        for lvl in range(4) for ff in (False, True)}
        self.tone = Module("tone_mapping", [])
        self.smaa = self.taa = None                                          # built on first use (smaa.wgsl / taa.wgsl)
        self.fsr_easu = self.fsr_rcas = None                                 # built on first use (src/shaders/fsr/source.zip)
        self.fsr_output = None                                               # upscale_output[0] / [1] under Upscale::Fsr1, post_process.rs:723
        self.upscale_ratio = upscale_ratio
        n = width * height                                                   # reservoirs: size.x * size.y records, light.rs:343
        rw, rh = self.rw, self.rh
        self.reservoir = [np.zeros((n, 16), np.uint32) for _ in range(10)]   # GpuPackedReservoir::default(), light.rs:347-356
        self.render = [np.zeros((rh, rw, 4), np.uint16) for _ in range(3)]   # scaled_size, light.rs:366-367
        self.variance = [np.zeros((rh, rw), np.float32) for _ in range(3)]
        self.albedo = np.zeros((height, width, 4), np.uint16)                # full size, light.rs:368
        self.internal = [np.zeros((rh, rw, 4), np.uint16) for _ in range(4)]
        self.internal_variance = np.zeros((rh, rw), np.float32)
        self.denoise_render = [np.zeros((rh, rw, 4), np.uint16) for _ in range(3)]
        self.tone_mapping_output = [np.zeros((rh, rw, 4), np.uint16) for _ in range(2)]   # post_process.rs:713, [head] is written
        self.head = 0                                                        # PostProcessTextures::head = counter % 2 of the last frame run
        self.upscale_output = None                                           # post_process.rs:715-724 / :726-731, allocated by upscale_node
        self.taa_output = None
        self.gbuffer = None
        self.previous_gbuffer = None
        self.dummy = np.zeros((1, 1, 4), np.uint8)

    @property
    def tone_mapped(self):
        return self.tone_mapping_output[self.head]

    def set_gbuffer(self, position, normal, depth_gradient, instance_material, velocity_uv):
        """the five G-buffer targets of the prepass (src/prepass.rs:43-47 formats), deferred size; what was current becomes the
        previous_* textures of deferred_bindings.wgsl (prepass.rs:312-321 swaps the two sets every frame; cleared at creation)"""
        if self.gbuffer is not None:
            self.previous_gbuffer = self.gbuffer
        else:
            self.previous_gbuffer = [np.zeros_like(np.ascontiguousarray(position, np.float32)), None, None, None,
                                     np.zeros_like(np.ascontiguousarray(velocity_uv, np.float32))]
        self.gbuffer = [np.ascontiguousarray(position, np.float32), np.ascontiguousarray(normal), np.ascontiguousarray(depth_gradient, np.float32),
                        np.ascontiguousarray(instance_material, np.float32), np.ascontiguousarray(velocity_uv, np.float32)]

    # ------------------------------------------------------------------------------------------------ bind groups
    def _view_groups(self, m, inputs):
        m.buffer("frame", np.frombuffer(bytes(inputs.frame), np.uint8))
        m.buffer("view", np.frombuffer(bytes(inputs.view), np.uint8))
        m.buffer("previous_view", np.frombuffer(bytes(inputs.previous_view), np.uint8))
        m.buffer("lights", lights_buffer(inputs.lights))
        g = self.gbuffer
        m.texture("position_texture", g[0], RGBA32F)
        m.texture("normal_texture", g[1], RGBA8SNORM)
        m.texture("depth_gradient_texture", g[2], RG32F)
        m.texture("instance_material_texture", g[3], RG32F)
        m.texture("velocity_uv_texture", g[4], RGBA32F)
        p = self.previous_gbuffer
        m.texture("previous_position_texture", p[0], RGBA32F)                   # read by smaa.wgsl / taa.wgsl only
        m.texture("previous_velocity_uv_texture", p[4], RGBA32F)

    def _light_groups(self, m, inputs, signal):
        m.keep = []
        self._view_groups(m, inputs)
        s = self.scene
        m.buffer("vertex_buffer", s["vertices"])
        m.buffer("primitive_buffer", s["primitives"])
        m.buffer("asset_node_buffer", nodes_buffer(s["asset_nodes"]))
        m.buffer("alias_table_buffer", s["alias_table"])
        m.buffer("instance_buffer", s["instances"])
        m.buffer("instance_node_buffer", nodes_buffer(s["instance_nodes"]))
        m.buffer("material_buffer", s["materials"])
        m.buffer("emissive_node_buffer", nodes_buffer(s["emissive_nodes"]))
        m.buffer("emissive_buffer", s["emissives"])
        if self.textures:
            for i, (rgba, mu, mv, linear, srgb) in enumerate(self.textures):
                rgba = np.ascontiguousarray(rgba)
                m.keep.append(rgba)
                m.texture("textures", rgba, RGBA8SRGB if srgb else RGBA8UNORM, index=i)
                m.sampler("samplers", mu, mv, linear, index=i)
        else:
            m.texture("textures", self.dummy, RGBA8UNORM)
            m.sampler("samplers", 1, 1, 0)
        for i in range(16):
            m.texture("noise_texture", self.noise[i], RGBA8UNORM, index=i)      # lib.rs:189-219: linear Rgba8Unorm, nearest, repeat
        m.sampler("noise_sampler", 0, 0, 0)
        m.texture("albedo_texture", self.albedo, RGBA16F)
        m.texture("variance_texture", self.variance[signal], R32F)
        m.texture("render_texture", self.render[signal], RGBA16F)
        temporal, spatial = [(0, 4), (2, 4), (6, 8)][signal]
        current = inputs.frame.number % 2                                       # LightTextures::head = counter % 2, light.rs:376
        previous = 1 - current
        m.buffer_inplace("previous_reservoir_buffer", self.reservoir[current + temporal])          # binding 0 <- "current_temporal"
        m.buffer_inplace("reservoir_buffer", self.reservoir[previous + temporal])                  # binding 1 <- "previous_temporal"
        m.buffer_inplace("previous_spatial_reservoir_buffer", self.reservoir[current + spatial])   # binding 2 <- "current_spatial"
        m.buffer_inplace("spatial_reservoir_buffer", self.reservoir[previous + spatial])           # binding 3 <- "previous_spatial"

    # ------------------------------------------------------------------------------------------------ LightNode::run
    def full_screen_albedo(self, inputs):
        m = self.light["base"]
        self._light_groups(m, inputs, 0)                  # render[0] / reservoir[0] bound, light.rs:647-648
        m.run("full_screen_albedo", self.w, self.h)

    def direct_lit(self, inputs, emissive):
        m = self.light["emissive" if emissive else "sun"]
        self._light_groups(m, inputs, 1 if emissive else 0)
        m.run("direct_lit", self.rw, self.rh)

    def indirect_lit_ambient(self, inputs):
        m = self.light["multi" if inputs.frame.indirect_bounces >= 2 else "base"]      # light.rs:668-671
        self._light_groups(m, inputs, 2)
        m.run("indirect_lit_ambient", self.rw, self.rh)

    def spatial_reuse(self, inputs, emissive):
        m = self.light["emissive" if emissive else "base"]
        self._light_groups(m, inputs, 1 if emissive else 2)
        m.run("spatial_reuse", self.rw, self.rh)

    def light_node(self, inputs):
        self.full_screen_albedo(inputs)
        self.direct_lit(inputs, False)
        self.direct_lit(inputs, True)
        if inputs.frame.emissive_spatial_reuse:
            self.spatial_reuse(inputs, True)
        self.indirect_lit_ambient(inputs)
        if inputs.frame.indirect_spatial_reuse:
            self.spatial_reuse(inputs, False)

    # ------------------------------------------------------------------------------------------------ PostProcessNode::run
    def _post_groups(self, m, inputs):
        m.keep = []
        self._view_groups(m, inputs)
        m.sampler("nearest_sampler", 1, 1, 0)
        m.sampler("linear_sampler", 1, 1, 1)

    def denoise_signal(self, inputs, signal):
        mods = [self.demodulation] + [self.denoise[(lvl, signal != 0)] for lvl in range(4)]      # denoise_direct: no firefly filter
        for k, m in enumerate(mods):
            self._post_groups(m, inputs)
            for i in range(4):
                m.texture(f"internal_texture_{i}", self.internal[i], RGBA16F)
            m.texture("internal_variance", self.internal_variance, R32F)
            m.texture("albedo_texture", self.albedo, RGBA16F)
            m.texture("variance_texture", self.variance[signal], R32F)
            m.texture("render_texture", self.render[signal], RGBA16F)
            m.texture("output_texture", self.denoise_render[signal], RGBA16F)
            m.run("demodulation" if k == 0 else "denoise", self.rw, self.rh)

    def tone_mapping(self, inputs, denoise, signals=3):
        m = self.tone
        self._post_groups(m, inputs)
        src = self.denoise_render if denoise else self.render
        fallback = np.zeros((1, 1, 4), np.uint16)
        m.keep.append(fallback)
        m.texture("direct_render_texture", src[0], RGBA16F)
        m.texture("emissive_render_texture", src[1], RGBA16F)
        m.texture("indirect_render_texture", src[2] if signals == 3 else fallback, RGBA16F)       # post_process.rs:948-953
        self.head = inputs.frame.number % 2                                                       # post_process.rs:735
        m.texture("output_texture", self.tone_mapping_output[self.head], RGBA16F)                 # :975-981
        m.run("tone_mapping", self.rw, self.rh)

    # ------------------------------------------------------------------------------------------------ temporal upscalers
    def _scaled(self, scale):
        """create_texture(format, scale): (size as f32 * scale).ceil(), post_process.rs:663-667"""
        return int(np.ceil(np.float32(self.w) * np.float32(scale))), int(np.ceil(np.float32(self.h) * np.float32(scale)))

    def fsr_node(self, inputs, taa, sharpness):
        """post_process.rs:1279-1308: EASU from upscale_input_texture (:1037-1040) into upscale_output[0], RCAS from there into
        upscale_output[1]; constants as FsrConstantsUniform::extract_component builds them (:519-534); both dispatched over
        (size * 2 + 15) / 16 workgroups (the surplus ones store outside the image: nothing)"""
        if self.fsr_easu is None:
            self.fsr_easu, self.fsr_rcas = FsrPass("SAMPLE_EASU"), FsrPass("SAMPLE_RCAS")
            self.fsr_output = [np.zeros((self.h, self.w, 4), np.uint16) for _ in range(2)]        # create_texture(format, 1.0) x 2
        src = self.taa_output[self.head] if taa else self.tone_mapping_output[self.head]
        constants = (float(self.rw), float(self.rh), float(self.rw), float(self.rh), float(self.w), float(self.h), float(sharpness), 0)
        groups = ((self.w + 15) // 16, (self.h + 15) // 16)       # the groups that cover the image, of the (2 size + 15) / 16 dispatched
        self.fsr_easu.run(src, self.fsr_output[0], constants, groups)
        self.fsr_rcas.run(self.fsr_output[0], self.fsr_output[1], constants, groups)

    def upscale_node(self, inputs, smaa, taa):
        """post_process.rs:1236-1277: smaa_tu4x + smaa_tu4x_extrapolate over scaled_size, then taa_jasmine over the (doubled) size"""
        scale = np.float32(1.0) / np.float32(self.upscale_ratio)                                  # settings.upscale.ratio().recip(), :711
        current, previous = self.head, 1 - self.head
        if smaa:
            scale = np.float32(scale * np.float32(2.0))                                           # :717
            ow, oh = self._scaled(scale)
            if self.smaa is None:
                self.smaa = Module("smaa", [])
                self.upscale_output = np.zeros((oh, ow, 4), np.uint16)
            m = self.smaa
            self._post_groups(m, inputs)
            m.texture("previous_render_texture", self.tone_mapping_output[previous], RGBA16F)     # :983-999
            m.texture("render_texture", self.tone_mapping_output[current], RGBA16F)
            m.texture("output_texture", self.upscale_output, RGBA16F)                             # :1001-1008
            m.run("smaa_tu4x", self.rw, self.rh)                                                  # :1240-1245
            m.run("smaa_tu4x_extrapolate", self.rw, self.rh)                                      # :1247-1255
        if taa:
            tw, th = self._scaled(scale)                                                          # taa_output at the (doubled) scale, :726-729
            if self.taa is None:
                self.taa = Module("taa", [])
                self.taa_output = [np.zeros((th, tw, 4), np.uint16) for _ in range(2)]
            m = self.taa
            self._post_groups(m, inputs)
            m.texture("previous_render_texture", self.taa_output[previous], RGBA16F)              # :1014-1027
            m.texture("render_texture", self.upscale_output if smaa else self.tone_mapping_output[current], RGBA16F)   # :1010-1013
            m.texture("output_texture", self.taa_output[current], RGBA16F)                        # :1028-1035
            sw, sh = (2 * self.rw, 2 * self.rh) if smaa else (self.rw, self.rh)                   # scaled_size *= 2, :1258
            m.run("taa_jasmine", sw, sh)                                                          # :1271-1276

    def post_process_node(self, inputs, denoise):
        signals = 3 if inputs.frame.indirect_bounces else 2
        if denoise:
            for s in range(signals):
                self.denoise_signal(inputs, s)
        self.tone_mapping(inputs, denoise, signals)
